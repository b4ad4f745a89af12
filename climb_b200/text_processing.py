"""Text half of ViltEncoderWrapper.process_inputs (src/modeling/vilt.py:83-96) without the HuggingFace tokenizer on the
path: `B200BertTokenizer` runs the BERT WordPiece pipeline of `BertTokenizerFast` (special-token split, BertNormalizer,
BertPreTokenizer, WordPiece, [CLS] .. [SEP], truncation, padding to the longest row) in native host code behind the C ABI
(`climb_wordpiece_*`, climb_b200/csrc/wordpiece.cu) and writes the int64 rows straight into (optionally pinned) staging
memory for the batch's host-to-device copy. Token ids are identical to the reference tokenizers' on the golden corpus
(tests/golden/tokenizer_golden.json, written by the vendored BertTokenizerFast / BertTokenizer) and on seeded fuzz against
the installed `tokenizers` library (tests/test_tokenizer.py).

Only what process_inputs uses is implemented: a list of single texts, truncation=True, padding=True ('longest') or
'max_length', return_tensors='pt'.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterable, List, Optional, Sequence, Union

import torch

from . import _lib


class B200BertTokenizer:
    def __init__(self, vocab: Union[str, Dict[str, int], Sequence[str]], do_lower_case: bool = True,
                 tokenize_chinese_chars: bool = True, n_threads: Optional[int] = None):
        """vocab: path of a vocab.txt, a token -> id dict (e.g. tokenizer.get_vocab()) or the token list in id order."""
        if isinstance(vocab, str):
            with open(vocab, "rb") as f:
                data = f.read()
            if data.endswith(b"\n"):
                data = data[:-1]
        else:
            if isinstance(vocab, dict):
                size = max(vocab.values()) + 1
                toks: List[Optional[str]] = [None] * size
                for t, i in vocab.items():
                    toks[i] = t
                # ids no token maps to get a line that can never match (clean_text drops NUL before the lookup)
                toks = [t if t is not None else f"\x00gap{i}" for i, t in enumerate(toks)]
            else:
                toks = list(vocab)
            if any("\n" in t for t in toks):
                raise ValueError("vocabulary tokens must not contain newlines")
            data = "\n".join(toks).encode("utf-8")
        self.do_lower_case, self.tokenize_chinese_chars = bool(do_lower_case), bool(tokenize_chinese_chars)
        self.n_threads = n_threads if n_threads is not None else min(8, os.cpu_count() or 1)
        self._vocab_bytes = data
        self._create()

    def _create(self):
        self._handle = _lib.climb_wordpiece_create(self._vocab_bytes, len(self._vocab_bytes), int(self.do_lower_case),
                                                   int(self.tokenize_chinese_chars))
        if not self._handle:
            msg = _lib.climb_last_error()
            raise _lib.ClimbError(msg.decode() if msg else "climb_wordpiece_create failed")

    def __getstate__(self):                 # the native handle does not survive pickling (torch.save(model)): rebuilt on load
        state = dict(self.__dict__)
        state["_handle"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._create()

    @classmethod
    def from_hf(cls, tokenizer, **kw) -> "B200BertTokenizer":
        """From a transformers BertTokenizer / BertTokenizerFast: same vocabulary and casing."""
        lower = getattr(tokenizer, "do_lower_case", None)
        if lower is None:
            lower = getattr(tokenizer, "init_kwargs", {}).get("do_lower_case", True)
        cjk = getattr(tokenizer, "init_kwargs", {}).get("tokenize_chinese_chars", True)
        return cls(tokenizer.get_vocab(), do_lower_case=lower, tokenize_chinese_chars=cjk, **kw)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        destroy = getattr(_lib, "climb_wordpiece_destroy", None) if _lib is not None else None
        if h and destroy is not None:       # (module globals may already be gone at interpreter shutdown)
            destroy(h)

    def __deepcopy__(self, memo):           # copy.deepcopy(model) shares the immutable native tokenizer
        return self

    def __call__(self, text: Union[str, Iterable[str]], max_length: int = 40, padding=True, truncation=True,
                 return_tensors: str = "pt", pin_memory: bool = False, **unused) -> Dict[str, torch.Tensor]:
        if return_tensors != "pt":
            raise NotImplementedError("B200BertTokenizer returns torch tensors (return_tensors='pt')")
        if truncation is not True and truncation != "longest_first":
            raise NotImplementedError("B200BertTokenizer implements truncation=True (what process_inputs uses)")
        if padding not in (True, "longest", "max_length"):
            raise NotImplementedError("B200BertTokenizer pads to the longest row (padding=True) or to max_length")
        texts = [text] if isinstance(text, str) else list(text)
        if any(not isinstance(t, str) for t in texts):
            raise TypeError("B200BertTokenizer takes a string or a list of strings (one text per sample)")
        n = len(texts)
        blobs = [t.encode("utf-8", "replace") for t in texts]
        offsets = (ctypes.c_int64 * (n + 1))()
        total = 0
        for i, b in enumerate(blobs):
            offsets[i] = total
            total += len(b)
        offsets[n] = total
        out = torch.empty((3, max(n, 1), max_length), dtype=torch.int64, pin_memory=bool(pin_memory))
        longest = ctypes.c_int(0)
        _lib.check(_lib.climb_wordpiece_encode(self._handle, b"".join(blobs), ctypes.addressof(offsets), n, int(max_length),
                                               out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), ctypes.byref(longest),
                                               int(self.n_threads)))
        T = max_length if padding == "max_length" else max(longest.value, 2 if n else 0)
        out = out[:, :n, :T]
        if T != max_length:
            out = out.contiguous()
            if pin_memory:
                out = out.pin_memory()
        return {"input_ids": out[0], "token_type_ids": out[2], "attention_mask": out[1]}
