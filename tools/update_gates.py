#!/usr/bin/env python
"""Regenerate tests/parity_gates.json from the `MEASURED <key> <error> gate <limit>` lines that the -m gpu parity
tests print (run them with `-s` on a B200 and pass the log files here):

    python tools/update_gates.py gpurun_out/parity.log gpurun_out/viltbert.log [...]

gate = max(FACTOR x the measured error, floor of the metric's class), rounded UP to two significant digits (FACTOR = 2: a
kernel change that doubles the error of a fixture fails the suite). The floors are 2 x the TYPICAL error of the arithmetic
on these fixtures -- bf16 mode: 5e-3 on pooled / logits, 1e-2 on gradients; bf16x3 mode: 1e-5 / 1e-4 -- because a fixture
with two logits per sample moves by a factor of two when nothing but the summation order of a kernel changes (measured:
base_nlvr2 logits 3.7e-3 with the round-1 attention forward, 8.0e-3 with the TMEM-P one), and a relative loss error of 1e-5
is fp32 noise. Keys measured more than once keep their largest measurement; keys absent from the logs are kept.
"""
import json
import math
import os
import re
import sys

FACTOR = 2.0
FLOOR = 1e-7          # fp32 noise level: errors that are exactly zero in one run still get a gate


def class_floor(key: str) -> float:
    precise = key.startswith("precise/")
    if key.endswith("/loss"):
        return 2e-5 if precise else 1e-3
    if "/grad" in key:
        return 2e-4 if precise else 2e-2
    return 2e-5 if precise else 1e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "parity_gates.json")
LINE = re.compile(r"MEASURED (\S+) ([0-9.eE+-]+) gate")


def round_up(x: float) -> float:
    if x <= 0:
        return FLOOR
    e = math.floor(math.log10(x)) - 1
    return math.ceil(x / 10 ** e) * 10 ** e


def main(paths):
    measured = {}
    for p in paths:
        for line in open(p, errors="replace"):
            m = LINE.search(line)
            if m:
                measured[m.group(1)] = max(measured.get(m.group(1), 0.0), float(m.group(2)))
    table = json.load(open(PATH)) if os.path.exists(PATH) else {}
    for k, v in measured.items():
        table[k] = float(f"{round_up(max(FACTOR * v, class_floor(k), FLOOR)):.3g}")
    json.dump(dict(sorted(table.items())), open(PATH, "w"), indent=1)
    print(f"{len(measured)} measurements -> {PATH} ({len(table)} gates)")
    mp = os.path.join(ROOT, "profiles", "r2_parity_measured.json")
    merged = json.load(open(mp)) if os.path.exists(mp) else {}           # keys absent from these logs keep their last measurement
    merged.update(measured)
    json.dump(dict(sorted(merged.items())), open(mp, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
