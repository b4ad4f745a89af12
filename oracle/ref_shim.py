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


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "adapter-transformers", "src", "transformers"))


def install():
    """Make `import transformers` resolve to the vendored 4.17 fork and `modeling.*`,
    `cl_algorithms.*`, `configs.*` to CLiMB's own packages. Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if getattr(install, "_done", False):
        return
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
