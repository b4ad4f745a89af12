"""Generate tests/golden/tokenizer_*.{json,txt} by running the reference's tokenizers -- the vendored
`transformers.BertTokenizerFast` (adapter-transformers 4.17 over the installed Rust `tokenizers` backend: what
ViltProcessor.from_pretrained gives CLiMB, src/modeling/vilt.py:49,491) and the vendored slow `BertTokenizer`
(tokenization_bert.py) -- exactly as ViltEncoderWrapper.process_inputs calls them (src/modeling/vilt.py:93-95:
padding=True, truncation=True, max_length=..., return_tensors='pt') on a synthetic vocabulary and corpus.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_tokenizer

TEST INFRASTRUCTURE (build container only). The real bert-base-uncased vocabulary is not available offline, so the
vocabulary is synthetic: whole words, pieces and single characters drawn from the corpus itself (some left out on purpose
so that [UNK] paths are exercised), in the vocab.txt format. Both tokenizers are given the from_pretrained treatment
(`sanitize_special_tokens`), i.e. literal "[SEP]" etc. in a text are kept whole.
"""
from __future__ import annotations

import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

SUBJECTS = ["the man", "a woman", "two dogs", "the cat", "a child", "three people", "the bus", "a red car", "the giraffe",
            "an old couple", "the pizza", "a skateboarder", "the tennis player", "a laptop", "the umbrella"]
VERBS = ["is sitting on", "is holding", "are standing near", "is looking at", "is riding", "are eating", "is wearing",
         "is next to", "are playing with", "is behind", "is jumping over", "is parked beside"]
OBJECTS = ["a bench", "the table", "a frisbee", "the street", "a surfboard", "the kitchen counter", "a blue umbrella",
           "the snowy mountain", "a stop sign", "the refrigerator", "a baseball bat", "the train station"]
QUESTIONS = ["What color is {o}?", "How many {s2} are there?", "Is {s} {v} {o}?", "What is {s} doing?", "Where is {o}?",
             "Why is {s} {v} {o}?", "What's written on {o}?", "Are there any {s2} in the picture?", "Who is {v} {o}?",
             "What time of day is it?", "Does {s} look happy?", "What kind of animal is that?"]
STRESS = [
    "", " ", "\t\n", "a", "A", "Hello, WORLD!!", "don't can't it's o'clock", "e-mail: someone@example.com (urgent)", "3.14159 vs 2,718",
    "Café crème brûlée à la façon de Noël", "İstanbul'da ÇAĞ ığdır", "Straße und Fuß ẞ", "ΑΣ ΣΊΣΥΦΟΣ Όσο", "ПРИВЕТ мир Ёлка",
    "ﬁnance ﬂuﬀy ǅ ǆ", "Ｆｕｌｌｗｉｄｔｈ　ｔｅｘｔ！", "中国人民 日本語のテキスト 漢字カタカナ", "한국어 텍스트 입니다", "豈 更 車 賈 滑",
    "العربية مَرْحَبًا", "हिन्दी में पाठ", "Tiếng Việt có dấu", "emoji 😀 👨‍👩‍👧 ✈️ done", "zero​width‍join﻿bom",
    "nbsp here em thin　ideographic", "line sep para\x85nel", "soft­hyphen ctrl\x01\x02\x7f chars \x00 nul � repl",
    "“quoted” ‘single’ — dash… «guillemets» ¿qué? ¡sí! § ¶ † ‡ • ‰", "$100 ^caret `tick` ~tilde |pipe| <lt> =eq+ plus",
    "x" * 100, "y" * 101, "z" * 250 + " tail", "supercalifragilisticexpialidocious antidisestablishmentarianism",
    "a [SEP] b [CLS] c [MASK] d [PAD] e [UNK] f", "[SEP][SEP] [sep] [ SEP ] [MASK]ed un[MASK]", "[", "]", "[[CLS]]",
    "ǖ ṩ ệ ȭ ǟ", "á̧ ë̄ ộ", "́ lone mark", "한 jamo", "𝒜𝓁𝓅𝒽𝒶 𝔹𝕠𝕝𝕕 𝟘𝟙𝟚", "\U0002b820 \U0002b920 \U00020000",
    " leading and trailing   spaces  ", "MiXeD CaSe WoRdS", "word." * 30, "the " * 80,
]
FUZZ_ALPHABET = ("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789" + " " * 12 + ".,!?'\"-()[]:;/" +
                 "éèêëàâäçñöüßøåæœÉÈÀÇÑÖÜİıΣσςαβγδεабвгдеж中国日本語한글カタ" + "́̈  \t\n​­\x00\x85“”—…€£")


def build_corpus(seed=1234):
    rng = random.Random(seed)
    texts = list(STRESS)
    for _ in range(260):
        s, v, o = rng.choice(SUBJECTS), rng.choice(VERBS), rng.choice(OBJECTS)
        kind = rng.random()
        if kind < 0.45:
            t = rng.choice(QUESTIONS).format(s=s, v=v, o=o, s2=s.split()[-1])
        elif kind < 0.8:
            t = f"{s} {v} {o}.".capitalize()
        else:       # VCR-style long answers
            t = f"{s} {v} {o} because {rng.choice(SUBJECTS)} {rng.choice(VERBS)} {rng.choice(OBJECTS)}, and then {rng.choice(SUBJECTS)} left."
        if rng.random() < 0.15:
            t = t.upper()
        texts.append(t)
    for _ in range(400):
        n = rng.randint(0, 60)
        texts.append("".join(rng.choice(FUZZ_ALPHABET) for _ in range(n)))
    return texts


def build_vocab(texts, seed=99):
    """vocab.txt lines: specials first (as bert-base-uncased: [PAD] = 0, [UNK] = 100-ish does not matter), then pieces."""
    from tokenizers.normalizers import BertNormalizer
    from tokenizers.pre_tokenizers import BertPreTokenizer
    rng = random.Random(seed)
    norm = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    pre = BertPreTokenizer()
    words = {}
    for t in texts:
        for w, _ in pre.pre_tokenize_str(norm.normalize_str(t)):
            words[w] = words.get(w, 0) + 1
    vocab = ["[PAD]"] + [f"[unused{i}]" for i in range(5)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    seen = set(vocab)

    def add(tok):
        if tok and tok not in seen and "\n" not in tok:
            seen.add(tok)
            vocab.append(tok)

    chars = sorted({c for w in words for c in w})
    for c in chars:                               # ~85 % of the characters, bare and as continuation
        if rng.random() < 0.85:
            add(c)
        if rng.random() < 0.85:
            add("##" + c)
    for w, n in sorted(words.items(), key=lambda kv: (-kv[1], kv[0])):
        if len(w) <= 30 and (n >= 3 or rng.random() < 0.35):
            add(w)
        if len(w) >= 4 and rng.random() < 0.5:    # pieces: a prefix and the matching continuation, plus a random inner piece
            k = rng.randint(1, len(w) - 1)
            add(w[:k])
            add("##" + w[k:])
            a = rng.randint(0, len(w) - 2)
            b = rng.randint(a + 1, len(w))
            add(("##" if a else "") + w[a:b])
    return vocab


def main():
    from oracle import ref_shim
    ref_shim.install()
    from transformers import BertTokenizer, BertTokenizerFast
    import tokenizers
    import transformers
    texts = build_corpus()
    vocab = build_vocab(texts)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    vpath = os.path.join(GOLDEN_DIR, "tokenizer_vocab.txt")
    with open(vpath, "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    out = {"transformers": transformers.__version__, "tokenizers": tokenizers.__version__, "texts": texts, "cases": []}
    for do_lower in (True, False):
        fast = BertTokenizerFast(vpath, do_lower_case=do_lower)
        slow = BertTokenizer(vpath, do_lower_case=do_lower)
        fast.sanitize_special_tokens()
        slow.sanitize_special_tokens()
        for max_length in (40, 12, 160):
            if not do_lower and max_length != 40:
                continue
            # the process_inputs call (src/modeling/vilt.py:93-95), in batches of 16 as a data loader would hand them over
            ids, mask, types, slow_ids = [], [], [], []
            for b in range(0, len(texts), 16):
                chunk = texts[b:b + 16]
                enc = fast(text=chunk, max_length=max_length, padding=True, truncation=True, return_tensors="pt")
                ids.append(enc["input_ids"].tolist())
                mask.append(enc["attention_mask"].tolist())
                types.append(enc["token_type_ids"].tolist())
                senc = slow(text=chunk, max_length=max_length, padding=True, truncation=True, return_tensors="pt")
                slow_ids.append(senc["input_ids"].tolist())
            slow_agrees = [[a == b for a, b in zip(x, y)] if len(x[0]) == len(y[0]) else [False] * len(x) for x, y in zip(ids, slow_ids)]
            n_dis = sum(1 for blk in slow_agrees for ok in blk if not ok)
            out["cases"].append({"do_lower_case": do_lower, "max_length": max_length, "batch": 16, "input_ids": ids,
                                 "attention_mask": mask, "token_type_ids": types, "slow_agrees": slow_agrees})
            print(f"lower={do_lower} max_length={max_length}: {len(texts)} texts, slow tokenizer differs on {n_dis}")
            if do_lower and max_length == 40:
                flat_fast = [r for blk in ids for r in blk]
                flat_slow = [r for blk in slow_ids for r in blk]
                for t, a, b in zip(texts, flat_fast, flat_slow):
                    if a != b:
                        print("   differs:", repr(t)[:70])
    with open(os.path.join(GOLDEN_DIR, "tokenizer_golden.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=True)
    unk = vocab.index("[UNK]")
    n_unk = sum(r.count(unk) for blk in out["cases"][0]["input_ids"] for r in blk)
    print(f"vocab {len(vocab)} tokens; [UNK] occurrences in case 0: {n_unk}")


if __name__ == "__main__":
    main()
