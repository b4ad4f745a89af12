"""Stage the UNMODIFIED reference for the benchmark's reference arm.

    python -m oracle.stage_ref          # (also run by __graft_entry__.build() when /root/reference is mounted)

TEST / BASELINE INFRASTRUCTURE, not product code. /root/reference exists in the build container only; the GPU box gets a
snapshot of this repo. The recipe packs the Python packages of the reference that its ViLT path imports -- the vendored
adapter-transformers fork and CLiMB's own modeling / configs / cl_algorithms / utils / data packages, byte for byte, with
their directory layout -- into ONE archive, oracle/_ref/climb_reference_src.zip. oracle/_ref/ is git-ignored (no reference
source enters the history) but travels to the GPU box with the snapshot, like the built .so files. oracle/ref_shim.py
unpacks the archive into a temporary directory when /root/reference is absent and imports the reference from there, so that
`bench.py --impl reference` and the `gpu_eager_baseline` leg time the reference's own classes (ViltContinualLearner.forward,
create_optimizer, the vendored ViltModel) instead of the oracle port.
"""
from __future__ import annotations

import os
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHIVE = os.path.join(ROOT, "oracle", "_ref", "climb_reference_src.zip")
SOURCE = os.environ.get("CLIMB_REFERENCE_ROOT", "/root/reference")
# (directory under the reference root, what it is)
PACKAGES = [
    ("src/adapter-transformers/src/transformers", "vendored adapter-transformers fork (transformers 4.17 + adapters)"),
    ("src/modeling", "CLiMB encoder wrappers / continual learners"),
    ("src/configs", "model / task configs"),
    ("src/cl_algorithms", "EWC / experience replay / adapters"),
    ("src/utils", "helpers imported by the packages above"),
    ("src/data", "dataset modules imported by configs.task_configs"),
    ("src/train", "task trainers imported by configs.task_configs (and driven by tests/golden/trainer_*.npz)"),
]


def stage(force: bool = False) -> str:
    """Build the archive from SOURCE. Returns its path ('' when the reference is not mounted and nothing was staged before)."""
    have_src = os.path.isdir(os.path.join(SOURCE, PACKAGES[0][0]))
    if not have_src:
        return ARCHIVE if os.path.exists(ARCHIVE) else ""
    newest = 0.0
    files = []
    for rel, _ in PACKAGES:
        base = os.path.join(SOURCE, rel)
        for dp, dn, fn in os.walk(base):
            dn[:] = sorted(d for d in dn if d != "__pycache__")
            for f in sorted(fn):
                if f.endswith((".pyc", ".pyo")):
                    continue
                full = os.path.join(dp, f)
                files.append((full, os.path.relpath(full, SOURCE)))
                newest = max(newest, os.path.getmtime(full))
    if not force and os.path.exists(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest:
        return ARCHIVE
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for full, arc in files:
            z.write(full, arc)
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


if __name__ == "__main__":
    path = stage(force="--force" in sys.argv)
    print(f"{path or 'reference not mounted and no archive staged'}"
          + (f" ({os.path.getsize(path) / 1e6:.1f} MB)" if path else ""))
