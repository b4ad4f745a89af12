"""Import shim for the UNMODIFIED reference (GLAMOR-USC/CLiMB + its vendored adapter-transformers
fork). TEST INFRASTRUCTURE ONLY: used by oracle/make_golden.py and tests that pin the oracle while
/root/reference is mounted (this container); nothing under climb_b200/ imports it, and the GPU box
has no /root/reference at all.

The four shims are the ones SURVEY.md section 8c lists; none of them edits the reference tree.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CLIMB_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ARCHIVE = os.path.join(_HERE, "_ref", "climb_reference_src.zip")      # written by oracle/stage_ref.py at build() time


def _mounted(root: str) -> bool:
    return os.path.isdir(os.path.join(root, "src", "adapter-transformers", "src", "transformers"))


def _unpack_staged() -> str:
    """The GPU box has no /root/reference: unpack the archive that oracle/stage_ref.py staged (unmodified reference
    packages, same directory layout) into a temporary directory, once per archive version."""
    import tempfile
    import zipfile
    st = os.stat(STAGED_ARCHIVE)
    dest = os.path.join(tempfile.gettempdir(), f"climb_b200_reference_{int(st.st_mtime)}_{st.st_size}")
    marker = os.path.join(dest, ".complete")
    if not os.path.exists(marker):
        tmp = dest + f".{os.getpid()}"
        with zipfile.ZipFile(STAGED_ARCHIVE) as z:
            z.extractall(tmp)
        try:
            os.rename(tmp, dest)
        except OSError:                      # another process won the race
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
        open(marker, "w").close()
    return dest


def reference_available() -> bool:
    return _mounted(REFERENCE_ROOT) or os.path.exists(STAGED_ARCHIVE)


def install():
    """Make `import transformers` resolve to the vendored 4.17 fork and `modeling.*`,
    `cl_algorithms.*`, `configs.*` to CLiMB's own packages. Idempotent."""
    global REFERENCE_ROOT
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT} and no staged archive at {STAGED_ARCHIVE}")
    if getattr(install, "_done", False):
        return
    if not _mounted(REFERENCE_ROOT):
        REFERENCE_ROOT = _unpack_staged()
    sys.dont_write_bytecode = True           # the reference tree is read-only
    for name in list(sys.modules):
        if name == "transformers" or name.startswith("transformers."):
            del sys.modules[name]            # drop the stock transformers if something imported it
    stub = types.ModuleType("transformers.dependency_versions_check")
    stub.dep_version_check = lambda *a, **k: None
    sys.modules["transformers.dependency_versions_check"] = stub     # dependency_versions_check.py:41
    import huggingface_hub as hh
    for n in ("HfFolder", "Repository", "create_repo", "list_repo_files", "whoami"):   # file_utils.py:51
        if not hasattr(hh, n):
            setattr(hh, n, type(n, (), {}))
    if "jsonlines" not in sys.modules:
        sys.modules["jsonlines"] = types.ModuleType("jsonlines")     # nlvr2_dataset.py:5
    sys.path[:0] = [os.path.join(REFERENCE_ROOT, "src", "adapter-transformers", "src"),
                    os.path.join(REFERENCE_ROOT, "src")]
    from transformers import BertTokenizerFast
    BertTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: None)   # vilt.py:49 (network)
    install._done = True


class StubProcessor:
    """Stands in for ViltProcessor (needs the bert-base-uncased vocab, not available offline)."""
    tokenizer = None
    feature_extractor = types.SimpleNamespace(size=384)
