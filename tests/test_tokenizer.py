"""CPU: the native BERT WordPiece tokenizer (climb_b200/csrc/wordpiece.cu behind climb_wordpiece_*, wrapped by
climb_b200.text_processing.B200BertTokenizer) against
  * tests/golden/tokenizer_golden.json -- the ids / masks the REFERENCE tokenizers (vendored BertTokenizerFast and the slow
    BertTokenizer of adapter-transformers 4.17) return for the process_inputs call (src/modeling/vilt.py:93-95) on a corpus of
    VQA / NLVR2 / VCR-style sentences, Unicode stress cases and seeded fuzz, written by oracle/make_golden_tokenizer.py;
  * the installed `tokenizers` library, live, on random strings over ALL Unicode scalar values;
  * the stock transformers tokenizer through B200BertTokenizer.from_hf and B200ViltEncoderWrapper.tokenize.
Integer work: everything is compared for equality. Host code only -- no GPU needed."""
import json
import os
import random
import unicodedata

import pytest
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VOCAB = os.path.join(GOLDEN_DIR, "tokenizer_vocab.txt")


def _golden():
    with open(os.path.join(GOLDEN_DIR, "tokenizer_golden.json"), encoding="utf-8") as f:
        return json.load(f)


def _vocab_dict():
    with open(VOCAB, encoding="utf-8") as f:
        return {t: i for i, t in enumerate(f.read().split("\n")[:-1])}


@pytest.mark.parametrize("case_idx", [0, 1, 2, 3])
def test_native_tokenizer_matches_reference_golden(case_idx):
    from climb_b200.text_processing import B200BertTokenizer
    g = _golden()
    case, texts = g["cases"][case_idx], g["texts"]
    assert len(texts) >= 700 and all(all(blk) for blk in case["slow_agrees"]), "fast and slow reference tokenizers agree on the corpus"
    tok = B200BertTokenizer(VOCAB, do_lower_case=case["do_lower_case"])
    for bi, b in enumerate(range(0, len(texts), case["batch"])):
        enc = tok(texts[b:b + case["batch"]], max_length=case["max_length"], padding=True, truncation=True, return_tensors="pt")
        for k in ("input_ids", "attention_mask", "token_type_ids"):
            ref = torch.tensor(case[k][bi])
            assert enc[k].dtype == torch.int64 and enc[k].shape == ref.shape, (k, bi, enc[k].shape, ref.shape)
            assert torch.equal(enc[k], ref), (k, bi, texts[b:b + case["batch"]])
    # padding='max_length' keeps every column; a single string is a batch of one
    enc = tok(texts[60:64], max_length=case["max_length"], padding="max_length")
    assert enc["input_ids"].shape == (4, case["max_length"])
    one = tok("What color is the cat?", max_length=case["max_length"])
    assert one["input_ids"].shape[0] == 1 and one["attention_mask"].sum() == one["input_ids"].shape[1]


@pytest.mark.parametrize("lower", [True, False])
def test_native_tokenizer_matches_tokenizers_library_on_random_unicode(lower):
    tk = pytest.importorskip("tokenizers")
    from climb_b200.text_processing import B200BertTokenizer
    vocab = _vocab_dict()
    ref = tk.Tokenizer(tk.models.WordPiece(vocab, unk_token="[UNK]"))
    ref.normalizer = tk.normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=lower)
    ref.pre_tokenizer = tk.pre_tokenizers.BertPreTokenizer()
    ref.add_special_tokens(["[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]"])
    rng = random.Random(11 + lower)
    known = [t for t in vocab if not t.startswith("[")]
    bmp_sample = [c for c in range(0x3000, 0x10000) if not 0xD800 <= c <= 0xDFFF][::97]

    def rand_text():
        out = []
        for _ in range(rng.randint(0, 20)):
            r = rng.random()
            if r < 0.35:
                out.append(rng.choice(known).replace("##", ""))       # vocabulary material: real WordPiece matches
                if rng.random() < 0.6:
                    out.append(" ")
            elif r < 0.55:
                out.append(chr(rng.randint(0x20, 0x7E)))
            elif r < 0.8:
                out.append(chr(rng.randint(0x80, 0x2FFF)))
            elif r < 0.9:
                out.append(chr(rng.choice(bmp_sample)))
            else:
                out.append(chr(rng.randint(0x10000, 0x10FFFF)))
        return "".join(out)

    texts = [rand_text() for _ in range(4000)]
    mine = B200BertTokenizer(VOCAB, do_lower_case=lower)(texts, max_length=48, padding="max_length")["input_ids"].tolist()
    want = ref.encode_batch(texts, add_special_tokens=False)
    n_pieces = 0
    for t, row, w in zip(texts, mine, want):
        exp = [vocab["[CLS]"]] + w.ids[:46] + [vocab["[SEP]"]]
        exp += [vocab["[PAD]"]] * (48 - len(exp))
        assert row == exp, (t, [hex(ord(c)) for c in t])
        n_pieces += sum(1 for i in w.ids if i != vocab["[UNK]"])
    assert n_pieces > 10000, "the fuzz must exercise real WordPiece matches, not only [UNK]"


def test_canonical_reordering_of_surviving_combining_marks():
    """NFD sorts runs of combining marks by class before the nonspacing ones are dropped; what survives (Hangul tone marks,
    musical symbols, viramas, and marks the library's category table does not know as nonspacing) must come out in the
    library's order. A vocabulary with one piece per mark makes the order visible in the ids."""
    tk = pytest.importorskip("tokenizers")
    from climb_b200.text_processing import B200BertTokenizer
    norm = tk.normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    cands = [cp for cp in range(0x300, 0x1F000) if unicodedata.combining(chr(cp)) != 0]
    survivors = [cp for cp in cands if norm.normalize_str("x" + chr(cp)) == "x" + chr(cp)]
    dropped = [cp for cp in cands if norm.normalize_str("x" + chr(cp)) == "x"]
    assert len(survivors) >= 60 and len(dropped) >= 500
    starters_dropped = [0x0941, 0x0A41, 0x0F35]      # nonspacing marks: class 0 ones are dropped but still end a run
    starters_dropped = [cp for cp in starters_dropped if unicodedata.category(chr(cp)) == "Mn" and unicodedata.combining(chr(cp)) == 0]
    toks = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "x", "y", "##x", "##y"]
    toks += [chr(cp) for cp in survivors] + ["##" + chr(cp) for cp in survivors]
    vocab = {t: i for i, t in enumerate(toks)}
    ref = tk.Tokenizer(tk.models.WordPiece(vocab, unk_token="[UNK]"))
    ref.normalizer = norm
    ref.pre_tokenizer = tk.pre_tokenizers.BertPreTokenizer()
    ref.add_special_tokens(["[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]"])
    rng = random.Random(5)
    alphabet = ([chr(c) for c in survivors] * 4 + [chr(c) for c in rng.sample(dropped, 120)] + [chr(c) for c in starters_dropped] * 6 +
                ["x", "y", "X", " ", "\u200d", "\u00ad", "\x01", "é", "한", "中"] * 8)
    texts = ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 14))) for _ in range(6000)]
    texts += ["x" + "".join(chr(rng.choice(survivors)) for _ in range(40))]                # a run longer than the hold-back buffer
    mine = B200BertTokenizer(toks)(texts, max_length=64, padding="max_length")["input_ids"].tolist()
    want = ref.encode_batch(texts, add_special_tokens=False)
    n_multi = 0
    for t, row, w in zip(texts, mine, want):
        if len(t) <= 14 or len(set(unicodedata.combining(c) for c in t[1:])) == 1:       # (the long run only if it needs no split)
            exp = [vocab["[CLS]"]] + w.ids[:62] + [vocab["[SEP]"]]
            assert row[:len(exp)] == exp and all(v == 0 for v in row[len(exp):]), [hex(ord(c)) for c in t]
        n_multi += sum(1 for a, b in zip(t, t[1:]) if unicodedata.combining(a) > unicodedata.combining(b) > 0)
    assert n_multi > 2000, "the corpus must contain out-of-order mark pairs"


def test_from_hf_and_encoder_wrapper_use_the_native_tokenizer():
    tr = pytest.importorskip("transformers")
    from climb_b200.modeling import B200ViltConfig, B200ViltEncoderWrapper, B200ViltModel
    from climb_b200.text_processing import B200BertTokenizer
    hf = tr.BertTokenizerFast(vocab=_vocab_dict())
    texts = ["What color is the cat?", "Is the man holding a Frisbee near the train station, or is he riding a skateboard behind the bus?",
             "", "Café İstanbul 中国 [MASK] don't"]
    ref = hf(text=texts, max_length=20, padding=True, truncation=True, return_tensors="pt")
    mine = B200BertTokenizer.from_hf(hf)(texts, max_length=20, padding=True, truncation=True, return_tensors="pt")
    for k in ("input_ids", "attention_mask", "token_type_ids"):
        assert torch.equal(mine[k], ref[k]), k
    cfg = B200ViltConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=64,
                         patch_size=32, vocab_size=len(_vocab_dict()), max_position_embeddings=20)
    enc = B200ViltEncoderWrapper(None, B200ViltModel(cfg), torch.device("cpu"))
    out = enc.tokenize(hf, texts)
    assert enc._native_tok is not None and torch.equal(out["input_ids"], ref["input_ids"])
    enc.native_tokenizer = False
    assert torch.equal(enc.tokenize(hf, texts)["input_ids"], ref["input_ids"])
    calls = []

    def other(text, max_length, padding, truncation, return_tensors):        # not a BERT tokenizer: called as is
        calls.append(len(text))
        return {"input_ids": torch.zeros(len(text), 3, dtype=torch.long)}
    enc.native_tokenizer = True
    assert enc.tokenize(other, texts)["input_ids"].shape == (4, 3) and calls == [4]
    # text PAIRS (the language-only multiple-choice path, convert_mc_batch_to_vilt_input_dict vilt.py:561-567) keep the
    # processor's own tokenizer: second segment with token_type_ids 1, pair truncation
    pairs = [["What color is the cat?", "black"], ["Is the man holding a Frisbee near the train station?", "he is riding a skateboard"]]
    ref_p = hf(text=pairs, max_length=20, padding=True, truncation=True, return_tensors="pt")
    got_p = enc.tokenize(hf, pairs)
    for k in ("input_ids", "attention_mask", "token_type_ids"):
        assert torch.equal(got_p[k], ref_p[k]), k
    assert int(got_p["token_type_ids"].max()) == 1
    # so does a tokenizer built with non-default normalisation
    hf2 = tr.BertTokenizerFast(vocab=_vocab_dict(), strip_accents=False)
    enc._native_tok = None
    ref2 = hf2(text=["Café"], max_length=20, padding=True, truncation=True, return_tensors="pt")
    assert torch.equal(enc.tokenize(hf2, ["Café"])["input_ids"], ref2["input_ids"]) and enc._native_tok is None


def test_threads_edge_cases_and_errors():
    from climb_b200 import _lib
    from climb_b200.text_processing import B200BertTokenizer
    g = _golden()
    texts = g["texts"][:300]
    a = B200BertTokenizer(VOCAB, n_threads=1)(texts, max_length=40)
    b = B200BertTokenizer(VOCAB, n_threads=8)(texts, max_length=40)
    assert all(torch.equal(a[k], b[k]) for k in a)
    many = g["texts"] * 3                                         # enough rows for the thread pool to split them
    m1 = B200BertTokenizer(VOCAB, n_threads=1)(many, max_length=24)
    m8 = B200BertTokenizer(VOCAB, n_threads=8)(many, max_length=24)
    assert all(torch.equal(m1[k], m8[k]) for k in m1)
    tok = B200BertTokenizer(VOCAB)
    empty = tok([], max_length=40)
    assert empty["input_ids"].shape[0] == 0
    two = tok(["", " \t "], max_length=40)                       # nothing but [CLS] [SEP]
    v = _vocab_dict()
    assert two["input_ids"].tolist() == [[v["[CLS]"], v["[SEP]"]]] * 2
    long = tok(["the " * 500], max_length=40)                    # truncation keeps max_length - 2 tokens between the specials
    assert long["input_ids"].shape == (1, 40) and long["input_ids"][0, -1] == v["[SEP]"] and long["attention_mask"].sum() == 40
    # a dict vocabulary with a hole in the id range and the token list form give the same ids
    d = dict(v)
    hole = d.pop("[unused3]")
    assert hole not in d.values()
    c = B200BertTokenizer(d)(texts[:50], max_length=40)
    T = c["input_ids"].shape[1]
    assert torch.equal(c["input_ids"], a["input_ids"][:50, :T]) and bool((a["input_ids"][:50, T:] == v["[PAD]"]).all())
    lst = B200BertTokenizer(list(v))(texts[:50], max_length=40)
    assert torch.equal(lst["input_ids"], c["input_ids"])
    import copy
    import pickle
    again = pickle.loads(pickle.dumps(tok))                      # torch.save(model) pickles the wrapper: the native handle is rebuilt
    assert torch.equal(again(texts[:50], max_length=40)["input_ids"], c["input_ids"]) and copy.deepcopy(tok) is tok
    with pytest.raises(_lib.ClimbError):
        B200BertTokenizer(["a", "b", "[UNK]"])                   # no [CLS] / [SEP] / [PAD]
    with pytest.raises(NotImplementedError):
        tok(["a"], truncation=False)
    with pytest.raises(NotImplementedError):
        tok(["a"], return_tensors="np")
    with pytest.raises(TypeError):
        tok([["a", "b"]])
    with pytest.raises(_lib.ClimbError):
        tok(["a"], max_length=1)


@pytest.mark.parametrize("case_idx", [0, 1, 3])
def test_python_restatement_matches_reference_golden(case_idx):
    """oracle/wordpiece_oracle.py (the slow tokenizer restated with file:line citations) against the same golden vectors."""
    from oracle import wordpiece_oracle as wo
    g = _golden()
    case, texts, vocab = g["cases"][case_idx], g["texts"], _vocab_dict()
    for bi, b in enumerate(range(0, len(texts), case["batch"])):
        ids, mask, types = wo.encode_batch(texts[b:b + case["batch"]], vocab, case["max_length"], case["do_lower_case"])
        assert ids == case["input_ids"][bi], (bi, texts[b:b + case["batch"]])
        assert mask == case["attention_mask"][bi] and types == case["token_type_ids"][bi]
