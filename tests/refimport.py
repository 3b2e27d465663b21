"""Locate and import the UNMODIFIED reference (ColdRec) from the git-ignored copy baseline/_ref/ (made by
baseline/install_ref.py in the build container; it travels to the GPU box with the snapshot).  Never /root/reference:
that path does not exist on the GPU box."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(ROOT, "baseline", "stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model", "BaseRecommender.py"))


def import_reference():
    """Put baseline/_ref (and the faiss import stub) on sys.path; returns the reference root."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_ref.py` where /root/reference exists")
    if "faiss" not in sys.modules:
        spec = importlib.util.spec_from_file_location("faiss", os.path.join(STUBS, "faiss.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["faiss"] = mod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return REF
