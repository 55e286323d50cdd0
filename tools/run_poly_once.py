"""Run every polynomial helper twice on device-resident data (target for ncu captures).
usage: run_poly_once.py FIELD LG      FIELD: gl64 | bb31 | bls12_381_fr"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sppark_b200 import _lib  # noqa: E402

FIELDS = {"gl64": (0, 8), "bb31": (1, 4), "bls12_381_fr": (2, 32)}
field, lg = sys.argv[1], int(sys.argv[2])
fid, ebytes = FIELDS[field]
n = 1 << lg
rng = np.random.default_rng(3)
if ebytes == 8:
    host = rng.integers(1, 2**64 - 2**32 + 1, size=n, dtype=np.uint64)
elif ebytes == 4:
    host = rng.integers(1, 0x78000001, size=n, dtype=np.uint32)
else:
    host = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    host[:, 7] &= 0x3fffffff
d = torch.from_numpy(host.reshape(-1).view(np.uint8)).cuda()
z = host.reshape(n, -1)[5].copy()
xs = d[: 4 * ebytes].clone()
ret = torch.zeros_like(xs)
l, s = _lib.lib(), torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(l.sppark_b200_prefix_op_dev(fid, 0, d.data_ptr(), d.data_ptr(), n, s))
    _lib.check(l.sppark_b200_prefix_op_dev(fid, 1, d.data_ptr(), d.data_ptr(), n, s))
    _lib.check(l.sppark_b200_div_by_x_minus_z_dev(fid, d.data_ptr(), n, z.ctypes.data, 1, s))
    _lib.check(l.sppark_b200_evaluate_dev(fid, ret.data_ptr(), xs.data_ptr(), 1, d.data_ptr(), n, s))
    _lib.check(l.sppark_b200_evaluate_dev(fid, ret.data_ptr(), xs.data_ptr(), 4, d.data_ptr(), n, s))
    _lib.check(l.sppark_b200_batch_inverse_dev(fid, d.data_ptr(), d.data_ptr(), n, s))
torch.cuda.synchronize()
print("done")
