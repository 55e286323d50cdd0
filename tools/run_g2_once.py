"""Run device-resident BLS12-381 G2 MSMs (for ncu captures): run_g2_once.py LG [REPS]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sppark_b200 import msm
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << lg
base = msm.generate_points_dev(msm.BLS12_381_G2, 1 << 12)
pts = base[torch.arange(n, device="cuda") % (1 << 12)].contiguous()
sc = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda")
for _ in range(reps):
    t = time.perf_counter()
    msm.msm_dev(msm.BLS12_381_G2, pts, sc)
    print("g2 msm 2^%d: %.1f ms" % (lg, (time.perf_counter() - t) * 1e3))
