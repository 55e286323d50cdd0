"""Run device-resident BLS12-381 MSMs (for ncu captures): run_msm_once.py LG [REPS]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as o
from sppark_b200 import msm
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << lg
base = o.gen_points("bls12_381", 1 << 12)
pts = np.tile(base, (max(1, n // base.shape[0]), 1))[:n].copy()
rng = np.random.default_rng(42)
sc = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
sc[:, 3] >>= np.uint64(2)
dp = torch.from_numpy(pts.view(np.int64)).cuda()
ds = torch.from_numpy(sc.view(np.int64)).cuda()
for _ in range(reps):
    t = time.perf_counter()
    msm.msm_dev(0, dp, ds)
    print("msm 2^%d: %.1f ms" % (lg, (time.perf_counter() - t) * 1e3))
