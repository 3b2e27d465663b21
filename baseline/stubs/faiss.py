"""Stand-in for the `faiss` package so that the UNMODIFIED reference tree under baseline/_ref imports.

`model/__init__.py` of ColdRec imports every trainer, and `model/KNN.py:4` / `model/NCL.py:7` import faiss at module
level; faiss is not in this image.  Nothing here computes anything: touching an attribute raises, so a code path that
really needs faiss fails loudly instead of producing numbers.  Test / baseline infrastructure only — the product
(`coldrec_b200/`) never imports this."""


def __getattr__(name):
    raise ImportError(f"faiss.{name}: faiss is not installed in this image (baseline/stubs/faiss.py is an import stub)")
