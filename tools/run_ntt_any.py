"""Run a few device-resident NTTs of any field (targets for ncu captures).
usage: run_ntt_any.py FIELD LG [ORDER [REPS [lde]]]   FIELD: gl64 | bb31 | bls12_381_fr"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sppark_b200 import ntt
field, lg = sys.argv[1], int(sys.argv[2])
order = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rng = np.random.default_rng(0)
if field == "gl64":
    h, fid = rng.integers(0, 2**64 - 2**32 + 1, size=1 << lg, dtype=np.uint64), ntt.GL64
    d = torch.from_numpy(h.view(np.int64)).cuda()
elif field == "bb31":
    h, fid = rng.integers(0, 0x78000001, size=1 << lg, dtype=np.uint32), ntt.BB31
    d = torch.from_numpy(h.view(np.int32)).cuda()
else:
    h, fid = rng.integers(0, 2**62, size=(1 << lg, 4), dtype=np.uint64), ntt.BLS12_381_FR
    d = torch.from_numpy(h.view(np.int64)).cuda()
if len(sys.argv) > 5 and sys.argv[5] == "lde":
    x = h[: 1 << (lg - 1)].copy()
    for _ in range(reps):
        ntt.LDE(0, x, 1, field=fid)
    ntt.coset_NTT(0, h.copy(), ntt.NN, field=fid)
else:
    for _ in range(reps):
        ntt.ntt_dev(d, order, field=fid)
torch.cuda.synchronize()
print("done")
