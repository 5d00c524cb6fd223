import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from scripts.bench_corr import run
shape = (32, 256, 36, 44, 4)
for cl in ("1", "2"):
    for bn, bk in (("256", "32"), ("128", "64")):
        os.environ.update(EEM_TF32_CLUSTER=cl, EEM_TF32_BN=bn, EEM_TF32_BK=bk)
        for dbg, label in [(0, "full"), (2, "no stores")]:
            os.environ["EEM_TF32_DEBUG"] = str(dbg)
            run(*shape, f"CL={cl} BN={bn} BK={bk} {label}")
        os.environ["EEM_TF32_DEBUG"] = "0"
        run(4, 256, 92, 160, 4, f"CL={cl} BN={bn} BK={bk} HREM B=4", iters=3)
