"""wgrad split-K / tile-width sweep (env NBEST_WGRAD_SPLITS, NBEST_GEMM_BN are read per call by the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from profiles.gemm_micro import bench
from nbest_b200 import ops
T2 = 12160
for (no, ki) in [(768, 3072), (3072, 768), (2304, 768), (768, 768)]:
    res = []
    for bn in (256, 128):
        for s in (1, 2, 3, 4, 6, 8, 12, 16):
            os.environ["NBEST_GEMM_BN"] = str(bn); os.environ["NBEST_WGRAD_SPLITS"] = str(s)
            us, tf = bench(no, ki, T2, ops.EPI_ACCUM_F32, a_mn=True, b_mn=True, reps=10)
            res.append((us, bn, s))
    res.sort()
    print("wgrad %4dx%4d best:" % (no, ki), ["%.1fus bn%d s%d" % r for r in res[:5]], "worst %.1f" % res[-1][0])
