"""Phase timings (CUDA events inside the library) of a device-resident MSM: LG [CURVE_ID]  (0 = BLS12-381 G1)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sppark_b200 import _lib, msm
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
curve = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << lg
m = 1 << min(14, lg)
dp = msm.generate_points_dev(curve, m).repeat(n // m, 1).contiguous()
rng = np.random.default_rng(42)
sc = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
sc[:, 3] >>= np.uint64(2)
ds = torch.from_numpy(sc.view(np.int64)).cuda()
msm.msm_dev(curve, dp, ds)
_lib.profile_enable(True)
for _ in range(2):
    msm.msm_dev(curve, dp, ds)
    ph = _lib.profile_read()
print(os.environ.get("SPPARK_B200_LIB", "default"), "curve", curve, "2^%d" % lg, " ".join("%s=%.1f" % kv for kv in ph), "total=%.1f ms" % sum(v for _, v in ph))
