"""Print per-tensor relative errors of hav_render_backward against the reference-minted goldens (debugging aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

from test_render_bwd_gpu import cuda_grads, rel_errors  # noqa: E402
from test_render_bwd_oracle import BWD_CASES, bwd_case  # noqa: E402

gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in BWD_CASES:
    z, case, sc, rnd, cot = bwd_case(gold, name)
    for prec in ("fp16", "bf16"):
        got = cuda_grads(sc, case, rnd, cot, prec)
        ref = {k[2:]: z[k] for k in z.files if k.startswith("g_")}
        err = rel_errors(got, ref)
        print(name, prec, "swap" if os.environ.get("HAV_BWD_SWAP_MN") else "noswap", {k: "%.2e" % e for k, e in err.items()}, flush=True)
        for k in ("layers_xyz.1.bias", "fc_alpha.weight"):
            print("   ", k, got[k].reshape(-1)[:4], ref[k].reshape(-1)[:4])
