"""CPU restatement of the reference's BERT tokenization for the text half of process_inputs (src/modeling/vilt.py:93-95),
following the slow tokenizer of the vendored library line by line (adapter-transformers/src/transformers/models/bert/
tokenization_bert.py; character classes from tokenization_utils.py:268-304). TEST INFRASTRUCTURE: pure Python, small cases;
pinned by tests/golden/tokenizer_golden.json (written by the vendored BertTokenizerFast AND BertTokenizer, which agree on
that corpus). The product path is climb_b200/csrc/wordpiece.cu; nothing under climb_b200/ imports this file.

Unicode data here is Python's `unicodedata`; the fast tokenizer CLiMB runs carries its own (older) tables, which is why the
native tokenizer's tables are probed from that library instead (tools/gen_bert_unicode_tables.py) -- on the golden corpus the
two agree.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, List

SPECIALS = ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")


def _is_whitespace(ch: str) -> bool:            # tokenization_utils.py:268-277
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:               # tokenization_utils.py:280-289
    return ch not in "\t\n\r" and unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:           # tokenization_utils.py:292-304
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_chinese_char(cp: int) -> bool:          # tokenization_bert.py:462-484
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


def basic_tokenize(text: str, do_lower_case: bool = True) -> List[str]:
    """BasicTokenizer.tokenize, tokenization_bert.py:379-414 (strip_accents unset: follows do_lower_case)."""
    cleaned = []
    for ch in text:                             # _clean_text, :486-497
        cp = ord(ch)
        if cp == 0 or cp == 0xFFFD or _is_control(ch):
            continue
        cleaned.append(" " if _is_whitespace(ch) else ch)
    spaced = []
    for ch in cleaned:                          # _tokenize_chinese_chars, :449-460
        spaced.extend([" ", ch, " "] if _is_chinese_char(ord(ch)) else [ch])
    out: List[str] = []
    for token in "".join(spaced).split():       # whitespace_tokenize, :108-114
        if do_lower_case:
            token = token.lower()
            token = "".join(c for c in unicodedata.normalize("NFD", token) if unicodedata.category(c) != "Mn")   # :416-425
        word: List[str] = []
        for ch in token:                        # _run_split_on_punc, :427-447
            if _is_punctuation(ch):
                if word:
                    out.append("".join(word))
                    word = []
                out.append(ch)
            else:
                word.append(ch)
        if word:
            out.append("".join(word))
    return " ".join(out).split()


def wordpiece(token: str, vocab: Dict[str, int], unk: str = "[UNK]", max_chars: int = 100) -> List[str]:
    """WordpieceTokenizer.tokenize for one token, tokenization_bert.py:508-555: greedy longest match first."""
    if len(token) > max_chars:
        return [unk]
    pieces, start = [], 0
    while start < len(token):
        end, cur = len(token), None
        while start < end:
            sub = ("##" if start > 0 else "") + token[start:end]
            if sub in vocab:
                cur = sub
                break
            end -= 1
        if cur is None:
            return [unk]
        pieces.append(cur)
        start = end
    return pieces


def encode_batch(texts: List[str], vocab: Dict[str, int], max_length: int, do_lower_case: bool = True):
    """tokenizer(text=texts, padding=True, truncation=True, max_length=max_length) -> (input_ids, attention_mask, token_type_ids)
    as lists of rows. Literal special tokens are kept whole (tokenization_utils.py tokenize(): the text is split on them
    first and, when lowercasing, everything else is lowercased character by character before BasicTokenizer sees it)."""
    rows = []
    for text in texts:
        ids: List[int] = []
        i = seg = 0
        segments = []
        while i < len(text):                    # leftmost-longest special-token split
            hit = None
            if text[i] == "[":
                for sp in SPECIALS:
                    if sp in vocab and text.startswith(sp, i) and (hit is None or len(sp) > len(hit)):
                        hit = sp
            if hit:
                segments.append((text[seg:i], None))
                segments.append((None, hit))
                i += len(hit)
                seg = i
            else:
                i += 1
        segments.append((text[seg:], None))
        for plain, special in segments:
            if special is not None:
                ids.append(vocab[special])
            elif plain:
                if do_lower_case:
                    # PreTrainedTokenizer.tokenize (tokenization_utils.py) lowercases everything but the special tokens through
                    # re.sub(r"(special)|(.+?)", ...): ONE CHARACTER per match, so str.lower()'s final-sigma rule never sees a word
                    # (capital sigma always becomes the medial form, as in the fast tokenizer's per-character lowercase)
                    plain = "".join(c.lower() for c in plain)
                for tok in basic_tokenize(plain, do_lower_case):
                    ids.extend(vocab[p] for p in wordpiece(tok, vocab))
        rows.append([vocab["[CLS]"]] + ids[:max_length - 2] + [vocab["[SEP]"]])
    longest = max((len(r) for r in rows), default=0)
    input_ids = [r + [vocab["[PAD]"]] * (longest - len(r)) for r in rows]
    mask = [[1] * len(r) + [0] * (longest - len(r)) for r in rows]
    types = [[0] * longest for _ in rows]
    return input_ids, mask, types
