"""Generate climb_b200/csrc/bert_unicode_tables.inc: the per-code-point behaviour of the BERT text normaliser and
pre-tokeniser that CLiMB's `BertTokenizerFast` runs (the `tokenizers` library: BertNormalizer(clean_text, handle_chinese_chars,
strip_accents = lowercase, lowercase) + BertPreTokenizer), obtained by PROBING that library for every Unicode scalar value --
so the native tokenizer (climb_b200/csrc/wordpiece.cu) agrees with the library's own Unicode tables, whatever their version.

    python tools/gen_bert_unicode_tables.py            # needs the `tokenizers` package (build container); output is committed

Per code point c, uncased mode, normalize_str(c) is one of
    c itself                 -> no entry
    ""                       -> REMOVED  (NUL, U+FFFD, category C* except \\t \\n \\r; lone nonspacing marks after NFD)
    " "                      -> SPACE    (\\t \\n \\r and White_Space characters)
    " " + c + " "            -> CJK      (the CJK ideograph ranges of BertNormalizer::handle_chinese_chars)
    anything else            -> MAP      (NFD with nonspacing marks dropped, then lowercase), listed explicitly; Hangul
                                          syllables decompose algorithmically and are not listed
cased mode (do_lower_case = False: clean_text + handle_chinese_chars only) has its own REMOVED / SPACE / CJK ranges.
PUNCT = code points BertPreTokenizer isolates ("a" + c + "a" splits into three pieces).

NFD also puts runs of combining marks into canonical order. Nonspacing marks are dropped afterwards, so the order only shows
between the few marks that SURVIVE (the library's category table is older than its normalisation table: 83 code points are
non-starters to NFD but not nonspacing to the filter). Probed as well:
    NONSTARTER_KEEP = survivors c for which "x" + U+302E + c + U+1D165 comes back with U+1D165 before U+302E (c did not break
                      the run), with their combining class (pairwise probing of all survivors agrees with these classes);
    TRANSPARENT     = code points that are dropped AND do not break a run (controls removed before NFD, nonspacing marks with
                      a non-zero class); every other dropped code point (nonspacing marks of class 0) ends the run.
"""
import os
import sys

from tokenizers.normalizers import BertNormalizer
from tokenizers.pre_tokenizers import BertPreTokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "climb_b200", "csrc", "bert_unicode_tables.inc")

S_BASE, L_COUNT, V_COUNT, T_COUNT = 0xAC00, 19, 21, 28


def ranges(cps):
    out, start, prev = [], None, None
    for cp in cps:
        if start is None:
            start = prev = cp
        elif cp == prev + 1:
            prev = cp
        else:
            out.append((start, prev))
            start = prev = cp
    if start is not None:
        out.append((start, prev))
    return out


def hangul_nfd(cp):
    s = cp - S_BASE
    l, v, t = 0x1100 + s // (V_COUNT * T_COUNT), 0x1161 + (s % (V_COUNT * T_COUNT)) // T_COUNT, 0x11A7 + s % T_COUNT
    return [l, v] + ([t] if t != 0x11A7 else [])


def classify(norm):
    removed, space, cjk, mapping = [], [], [], []
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        c = chr(cp)
        o = norm.normalize_str(c)
        if o == c:
            continue
        if o == "":
            removed.append(cp)
        elif o == " ":
            space.append(cp)
        elif o == " " + c + " ":
            cjk.append(cp)
        else:
            cps = [ord(x) for x in o]
            if S_BASE <= cp < S_BASE + L_COUNT * V_COUNT * T_COUNT:
                assert cps == hangul_nfd(cp), hex(cp)
                continue
            mapping.append((cp, cps))
    return removed, space, cjk, mapping


def probe_reordering(norm):
    import itertools
    import unicodedata
    hi, lo = chr(0x302E), chr(0x1D165)
    assert norm.normalize_str("x" + hi + lo) == "x" + lo + hi, "the probe marks are not reordered by this library version"
    keep, transparent = [], []
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        c = chr(cp)
        iso = norm.normalize_str("x" + c)
        out = norm.normalize_str("x" + hi + c + lo)
        if iso == "x":
            if out == "x" + lo + hi:
                transparent.append(cp)
        elif hi in out and lo in out and out.index(lo) < out.index(hi):
            assert iso == "x" + c, hex(cp)           # no surviving non-starter is also mapped
            keep.append(cp)
    cls = {cp: unicodedata.combining(chr(cp)) for cp in keep}
    assert all(v > 0 for v in cls.values())
    for a, b in itertools.permutations(keep, 2):    # the classes explain every pairwise swap the library makes
        swapped = norm.normalize_str("x" + chr(a) + chr(b)) == "x" + chr(b) + chr(a)
        assert swapped == (cls[a] > cls[b]), (hex(a), hex(b))
    return [(cp, cls[cp]) for cp in keep], transparent


def emit_ranges(f, name, rs):
    f.write(f"static const uint32_t {name}[][2] = {{\n")
    for i in range(0, len(rs), 6):
        f.write("    " + " ".join(f"{{0x{a:X}, 0x{b:X}}}," for a, b in rs[i:i + 6]) + "\n")
    f.write("};\n")
    f.write(f"static const int {name}_count = {len(rs)};\n\n")


def main():
    import tokenizers
    uncased = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    cased = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=False)
    u_removed, u_space, u_cjk, u_map = classify(uncased)
    c_removed, c_space, c_cjk, c_map = classify(cased)
    assert not c_map, c_map[:3]
    pre = BertPreTokenizer()
    punct = [cp for cp in range(0x110000) if not (0xD800 <= cp <= 0xDFFF) and len(pre.pre_tokenize_str("a" + chr(cp) + "a")) == 3]
    keep, transparent = probe_reordering(uncased)
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_bert_unicode_tables.py (probing tokenizers %s: BertNormalizer / BertPreTokenizer over every\n"
                "// Unicode scalar value). Do not edit. Inclusive code-point ranges, sorted; MAP entries sorted by code point.\n\n" % tokenizers.__version__)
        emit_ranges(f, "kUncasedRemoved", ranges(u_removed))
        emit_ranges(f, "kUncasedSpace", ranges(u_space))
        emit_ranges(f, "kUncasedCjk", ranges(u_cjk))
        emit_ranges(f, "kCasedRemoved", ranges(c_removed))
        emit_ranges(f, "kCasedSpace", ranges(c_space))
        emit_ranges(f, "kCasedCjk", ranges(c_cjk))
        emit_ranges(f, "kPunct", ranges(punct))
        emit_ranges(f, "kUncasedTransparent", ranges(transparent))
        f.write("struct BertNonStarter { uint32_t cp; uint32_t cls; };\n")
        f.write("static const BertNonStarter kUncasedNonStarterKeep[] = {\n")
        for i in range(0, len(keep), 6):
            f.write("    " + " ".join(f"{{0x{a:X}, {c}}}," for a, c in keep[i:i + 6]) + "\n")
        f.write("};\n")
        f.write(f"static const int kUncasedNonStarterKeep_count = {len(keep)};\n\n")
        # MAP: key table (cp, offset, length) + flat output array
        flat, keys = [], []
        for cp, cps in u_map:
            keys.append((cp, len(flat), len(cps)))
            flat.extend(cps)
        f.write("struct BertMapKey { uint32_t cp; uint32_t offset; uint32_t length; };\n")
        f.write("static const BertMapKey kUncasedMapKeys[] = {\n")
        for i in range(0, len(keys), 4):
            f.write("    " + " ".join(f"{{0x{a:X}, {o}, {l}}}," for a, o, l in keys[i:i + 4]) + "\n")
        f.write("};\n")
        f.write(f"static const int kUncasedMapKeys_count = {len(keys)};\n")
        f.write("static const uint32_t kUncasedMapOut[] = {\n")
        for i in range(0, len(flat), 12):
            f.write("    " + " ".join(f"0x{a:X}," for a in flat[i:i + 12]) + "\n")
        f.write("};\n")
    print(f"{OUT}: uncased removed {len(u_removed)} space {len(u_space)} cjk {len(u_cjk)} map {len(u_map)} ({len(flat)} cps); "
          f"cased removed {len(c_removed)} space {len(c_space)} cjk {len(c_cjk)}; punct {len(punct)}; surviving non-starters "
          f"{len(keep)}, transparent {len(transparent)}", file=sys.stderr)


if __name__ == "__main__":
    main()
