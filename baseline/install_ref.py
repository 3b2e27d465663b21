#!/usr/bin/env python
"""Copy the UNMODIFIED reference (ColdRec, /root/reference) to the git-ignored baseline/_ref/ so that it travels to the
GPU box with the gpurun snapshot (the GPU box has no /root/reference).

ColdRec has no setup.py / pyproject.toml — there is nothing to `pip install`; it is a source tree run in place
(`python main.py ...`), so "installing" it is copying its Python sources.  Used by: `bench.py --impl reference` (the
reference's own `_evaluate` / `ranking_evaluation` / `LGCN_Encoder` timed on the host cores) and the drop-in tests
(`tests/test_gpu_dropin.py`: the reference's own `MF` / `LightGCN` trainers with `FusedEvalMixin` grafted on).
Never imported by the product path.  `python baseline/install_ref.py [src]`; `__graft_entry__.build()` calls it."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP_DIRS = ("model", "util", "config", "data")       # data/: only split.py / convert.py / README (no datasets are shipped)


def install(src: str = "/root/reference") -> bool:
    if not os.path.isdir(os.path.join(src, "model")):
        return os.path.isdir(os.path.join(DST, "model"))      # nothing to copy from (GPU box): use what travelled
    tmp = DST + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    for d in KEEP_DIRS:
        shutil.copytree(os.path.join(src, d), os.path.join(tmp, d),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.npy", "*.csv", "*.pkl", "*.pt", "*.zip"))
    for f in ("main.py", "LICENSE", "README.md"):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy2(os.path.join(src, f), os.path.join(tmp, f))
    shutil.rmtree(DST, ignore_errors=True)
    os.replace(tmp, DST)
    return True


if __name__ == "__main__":
    ok = install(*sys.argv[1:2])
    print("baseline/_ref:", "installed" if ok else "reference source not available")
