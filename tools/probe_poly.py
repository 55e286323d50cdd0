"""Device-resident timing of the polynomial helpers, ours next to the reference's own kernels
(oracle/_ref/libref_poly_*_gpu.so).  One reference library per process:
    python tools/probe_poly.py gl64|bb31|bls12_381_fr [lg ...]
Each figure is the mean of `reps` back-to-back launches between two CUDA events on the stream the
kernels run on; inputs are restored between repetitions only where the operation is not
idempotent in cost (none is data dependent), so the timed region holds kernels only."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import poly as op  # noqa: E402
from oracle import pyoracle as o  # noqa: E402
from sppark_b200 import _lib  # noqa: E402

field = sys.argv[1]
lgs = [int(v) for v in sys.argv[2:]] or [16, 20, 24]
f = op.FIELDS[field]
fid = f["id"]
ebytes = f["words"] * np.dtype(f["dtype"]).itemsize
ref = C.CDLL(o.ref_path(f"libref_poly_{field if field != 'bls12_381_fr' else 'bls12_381'}_gpu.so"))
ref.ref_poly_stream.restype = C.c_void_p
ref.ref_prefix_op_dev.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
ref.ref_div_by_x_minus_z_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
ref.ref_evaluate_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
l = _lib.lib()
rstream = torch.cuda.ExternalStream(ref.ref_poly_stream())
ours = torch.cuda.current_stream()


def timed(stream, fn, reps):
    for _ in range(3):
        fn()
    stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / reps      # microseconds


print(f"# {field}: microseconds per call, device resident; GB/s = algorithmic bytes (read + write of the array) / time")
for lg in lgs:
    n = 1 << lg
    host = op.seeded_input(field, n, 7)
    d = torch.from_numpy(host.reshape(-1).view(np.uint8)).cuda()
    z = op.encode(field, [123456789])
    xs = torch.from_numpy(op.encode(field, [3, 5, 7, 11]).reshape(-1).view(np.uint8)).cuda()
    ret = torch.zeros_like(xs)
    reps = 20 if lg <= 20 else 5
    s = ours.cuda_stream
    rows = [
        ("prefix add", lambda: l.sppark_b200_prefix_op_dev(fid, 0, d.data_ptr(), d.data_ptr(), n, s),
         lambda: ref.ref_prefix_op_dev(0, d.data_ptr(), n), 2),
        ("prefix mul", lambda: l.sppark_b200_prefix_op_dev(fid, 1, d.data_ptr(), d.data_ptr(), n, s),
         lambda: ref.ref_prefix_op_dev(1, d.data_ptr(), n), 2),
        ("div_by_x_minus_z", lambda: l.sppark_b200_div_by_x_minus_z_dev(fid, d.data_ptr(), n, z.ctypes.data, 0, s),
         lambda: ref.ref_div_by_x_minus_z_dev(d.data_ptr(), n, z.ctypes.data, 0), 2),
        ("div rotate", lambda: l.sppark_b200_div_by_x_minus_z_dev(fid, d.data_ptr(), n, z.ctypes.data, 1, s),
         lambda: ref.ref_div_by_x_minus_z_dev(d.data_ptr(), n, z.ctypes.data, 1), 2),
        ("evaluate 1 pt", lambda: l.sppark_b200_evaluate_dev(fid, ret.data_ptr(), xs.data_ptr(), 1, d.data_ptr(), n, s),
         lambda: ref.ref_evaluate_dev(ret.data_ptr(), xs.data_ptr(), 1, d.data_ptr(), n), 1),
        ("evaluate 4 pts", lambda: l.sppark_b200_evaluate_dev(fid, ret.data_ptr(), xs.data_ptr(), 4, d.data_ptr(), n, s),
         lambda: ref.ref_evaluate_dev(ret.data_ptr(), xs.data_ptr(), 4, d.data_ptr(), n), 1),
        ("batch inverse", lambda: l.sppark_b200_batch_inverse_dev(fid, d.data_ptr(), d.data_ptr(), n, s), None, 2),
    ]
    for name, mine, theirs, passes in rows:
        t = timed(ours, mine, reps)
        gbs = passes * n * ebytes / t / 1e3
        line = f"lg={lg:2d} {name:18s} ours {t:10.1f} us ({gbs:7.1f} GB/s)"
        if theirs is not None:
            tr = timed(rstream, theirs, reps)
            line += f"   reference {tr:10.1f} us   ratio {tr / t:5.2f}x"
        print(line, flush=True)
