"""Which (operation, length) pairs the REFERENCE's polynomial templates accept on this GPU
(oracle/ref_poly.cu).  Prints the CUDA error code per call; one library per process."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import poly as op  # noqa: E402
from oracle import pyoracle as o  # noqa: E402

field, group, lens = sys.argv[1], sys.argv[2], [int(v) for v in sys.argv[3:]]
lib = C.CDLL(o.ref_path(f"libref_poly_{field if field != 'bls12_381_fr' else 'bls12_381'}_gpu.so"))
lib.ref_prefix_op.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
lib.ref_div_by_x_minus_z.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
lib.ref_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
f = op.FIELDS[field]
for n in lens:
    x = op.encode(field, list(range(1, min(n, 3000) + 1)) * (n // min(n, 3000) + 1))[:n].copy()
    z = op.encode(field, [5])
    xs = op.encode(field, [3, 7])
    r = []
    vals = op.decode(field, x)
    p = f["p"]
    if group == "prefix":
        for opc, name in ((0, "add"), (1, "mul")):
            y = x.copy()
            e = lib.ref_prefix_op(opc, y.ctypes.data, n)
            r.append((e, e == 0 and np.array_equal(y, op.encode(field, op.prefix_op(p, name, vals)))))
    elif group == "div":
        for rot in (0, 1):
            y = x.copy()
            e = lib.ref_div_by_x_minus_z(y.ctypes.data, n, z.ctypes.data, rot)
            r.append((e, e == 0 and np.array_equal(y, op.encode(field, op.div_by_x_minus_z(p, vals, 5, bool(rot))))))
    else:
        for npts in (1, 2):
            ret = np.zeros_like(xs)
            e = lib.ref_evaluate(ret.ctypes.data, xs.ctypes.data, npts, x.ctypes.data, n)
            r.append((e, e == 0 and np.array_equal(ret[:npts], op.encode(field, op.evaluate(p, vals, [3, 7][:npts])))))
    print(field, group, n, "-> (cuda error, equals oracle)", r, flush=True)
