"""Run a few device-resident gl64 NTTs (for ncu captures)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sppark_b200 import ntt
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
order = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(0)
h = rng.integers(0, 2**64 - 2**32 + 1, size=1 << lg, dtype=np.uint64)
d = torch.from_numpy(h.view(np.int64)).cuda()
for _ in range(reps):
    ntt.ntt_dev(d, order)
torch.cuda.synchronize()
print("done")
