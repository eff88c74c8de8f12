"""Phase timing of render_tc2 (needs a build with HAV_NVCC_DEFS=-DHAV_TC_TIMING).  Prints cycles per tile of
thread 0 / pair 0 / CTA 0 in each phase."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from havatar_b200 import render, synth, _lib
L = C.CDLL(_lib.LIB_PATH)
sc = synth.scene(batch=1, height=512, width=512, seed=0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
args = [dev(sc[k]) for k in ("ray_batch", "background_prior", "inv_head_T", "planes", "wvol")]
w = {k: dev(v) for k, v in sc["weights"].items()}
out = render.render_rays(*args, w, 64, 0, precision="fp16"); torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
L.hav_debug_phase_cycles(buf, 1)
out = render.render_rays(*args, w, 64, 0, precision="fp16"); torch.cuda.synchronize()
L.hav_debug_phase_cycles(buf, 0)
tiles = 64 * ((2048 - 0 + 295) // 296)   # ray blocks of pair 0 of CTA 0 (7) x 64 samples
names = {0: "P row phase", 1: "P bar(stage)", 2: "P wait x_free", 3: "P PE+gather", 4: "P fence+arrive",
         8: "C(t0) wait x_full", 9: "C issue L0 + z bookkeeping", 10: "C wait L0", 11: "C epilogue0", 12: "C bar",
         13: "C issue L1", 14: "C wait L1", 15: "C epilogue1", 16: "C bar", 17: "C issue head", 18: "C wait head",
         19: "C composite", 20: "C bar(end)"}
tp = sum(buf[i] for i in range(0, 5)); tcn = sum(buf[i] for i in range(8, 21))
for i, nme in names.items():
    print("%-28s %8.0f cyc/tile" % (nme, buf[i] / tiles))
print("producer total %.0f  consumer total %.0f cyc/tile (tiles=%d)" % (tp / tiles, tcn / tiles, tiles))
