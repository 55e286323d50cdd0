"""GPU probe: BLS12-381 G1 MSM, ours (device-resident and host-pointer) vs the reference's own
GPU build (oracle/_ref/libref_msm_gpu.so, host-pointer API only).  Development tool."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402
from sppark_b200 import _lib, msm  # noqa: E402


class RE(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_void_p)]


def main():
    lgs = [int(x) for x in sys.argv[1:]] or [16, 20, 22, 24]
    print(torch.cuda.get_device_name(0), flush=True)
    refp = os.path.join(ROOT, "oracle", "_ref", "libref_msm_gpu.so")
    ref = None
    if os.path.exists(refp):
        ref = C.CDLL(refp)
        ref.mult_pippenger.restype = RE
        ref.mult_pippenger.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    base = o.gen_points("bls12_381", 1 << 14)
    rng = np.random.default_rng(42)
    for lg in lgs:
        n = 1 << lg
        pts = np.tile(base, (max(1, n // base.shape[0]), 1))[:n].copy()
        sc = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
        sc[:, 3] >>= np.uint64(2)
        dp = torch.from_numpy(pts.view(np.int64)).cuda()
        ds = torch.from_numpy(sc.view(np.int64)).cuda()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            got = msm.msm_dev(0, dp, ds)
            ts.append(time.perf_counter() - t)
        print(f"ours dev  2^{lg}: {['%.1f ms' % (x*1e3) for x in ts]}", flush=True)
        del dp, ds
        ts = []
        for _ in range(2):
            t = time.perf_counter()
            got_h = msm.multi_scalar_mult(pts, sc)
            ts.append(time.perf_counter() - t)
        print(f"ours host 2^{lg}: {['%.1f ms' % (x*1e3) for x in ts]}  same={np.array_equal(o.jac_to_affine('bls12_381', got), o.jac_to_affine('bls12_381', got_h))}", flush=True)
        if ref is not None:
            out = np.zeros(18, dtype=np.uint64)
            ts = []
            for _ in range(2):
                t = time.perf_counter()
                e = ref.mult_pippenger(out.ctypes.data, pts.ctypes.data, n, sc.ctypes.data)
                ts.append(time.perf_counter() - t)
                assert e.code == 0, e.code
            same = np.array_equal(o.jac_to_affine("bls12_381", out), o.jac_to_affine("bls12_381", got))
            print(f"REF  host 2^{lg}: {['%.1f ms' % (x*1e3) for x in ts]}  ours==ref: {same}", flush=True)
    print("launches:", _lib.launch_count())


if __name__ == "__main__":
    main()
