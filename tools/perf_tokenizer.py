"""Throughput of the text half of process_inputs on the host: native WordPiece (climb_wordpiece_encode) against the
`tokenizers` library behind transformers' BertTokenizerFast, on batches of 64 VQA / VCR-style texts, synthetic vocabulary
(tests/golden/tokenizer_vocab.txt). CPU only.   python tools/perf_tokenizer.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from climb_b200.text_processing import B200BertTokenizer  # noqa: E402


def bench(fn, n=200):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n


def main():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json"), encoding="utf-8"))
    texts = [t for t in g["texts"] if 20 <= len(t) <= 160 and t.isascii()][:64]
    assert len(texts) == 64
    vp = os.path.join(ROOT, "tests", "golden", "tokenizer_vocab.txt")
    vocab = {t: i for i, t in enumerate(open(vp, encoding="utf-8").read().split("\n")[:-1])}
    out = {"batch": 64, "mean_chars": sum(map(len, texts)) / 64, "host_cores": os.cpu_count()}
    tok = B200BertTokenizer(vp)
    s = bench(lambda: tok(texts, max_length=40))
    out["native"] = {"us_per_batch": round(s * 1e6, 1), "texts_per_s": round(64 / s)}
    big = texts * 64                                             # 4096 texts: the thread pool splits them
    for threads in (1, 8):
        tk = B200BertTokenizer(vp, n_threads=threads)
        s = bench(lambda: tk(big, max_length=40), n=20)
        out[f"native_4096_texts_{threads}_threads"] = {"ms": round(s * 1e3, 2), "texts_per_s": round(len(big) / s)}
    try:
        from transformers import BertTokenizerFast
        hf = BertTokenizerFast(vocab=vocab)
        s = bench(lambda: hf(text=texts, max_length=40, padding=True, truncation=True, return_tensors="pt"))
        out["transformers_BertTokenizerFast"] = {"us_per_batch": round(s * 1e6, 1), "texts_per_s": round(64 / s)}
    except Exception as e:  # pragma: no cover
        out["transformers_BertTokenizerFast"] = {"unavailable": repr(e)[:100]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
