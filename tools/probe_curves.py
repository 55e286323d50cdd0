"""GPU probe: host-pointer MSM on the other curves (BLS12-381 G2, BN254, BLS12-377), this library
against the reference's own CUDA build of the same templates (oracle/_ref).  Development tool."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sppark_b200 import msm  # noqa: E402


class RE(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_void_p)]


def best(fn, reps=4):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts[1:])


def main():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    cases = [("bls12_381_g2", msm.BLS12_381_G2, 12, "libref_msm_g2_packed_gpu.so", "ref_mult_pippenger_fp2", False),
             ("bn254_g2", msm.BN254_G2, 8, "libref_msm_g2_packed_bn254_gpu.so", "ref_mult_pippenger_fp2", False),
             ("bls12_377_g2", msm.BLS12_377_G2, 12, "libref_msm_g2_packed_bls12_377_gpu.so", "ref_mult_pippenger_fp2", False),
             ("bn254", msm.BN254_G1, 4, "libref_msm_bn254_gpu.so", "mult_pippenger_inf", True),
             ("bls12_377", msm.BLS12_377_G1, 6, "libref_msm_bls12_377_gpu.so", "mult_pippenger_inf", True),
             ("bls12_381_g1", msm.BLS12_381_G1, 6, "libref_msm_gpu.so", "mult_pippenger", False)]
    only = sys.argv[1:]
    for name, cid, nl, lib, sym, inf in cases:
        if only and name not in only:
            continue
        for lg in (16, 20, 22):
            n = 1 << lg
            base = msm.generate_points_dev(cid, 1 << 12)
            pts = base[torch.arange(n, device="cuda") % (1 << 12)].cpu().numpy().view(np.uint64)
            sc = torch.randint(0, 2**61, (n, 4), dtype=torch.int64).numpy().view(np.uint64)
            if inf:
                ark = np.zeros((n, 2 * nl + 1), dtype=np.uint64)
                ark[:, :2 * nl] = pts
                pts = ark
            ours = best(lambda: msm.msm(cid, pts, sc))
            line = f"{name:13s} 2^{lg}: ours {ours:8.1f} ms"
            p = os.path.join(ref_dir, lib)
            if os.path.exists(p):
                ref = C.CDLL(p)
                f = getattr(ref, sym)
                f.restype = RE
                out = np.zeros(3 * nl, dtype=np.uint64)
                if inf:
                    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
                    t = best(lambda: f(out.ctypes.data, pts.ctypes.data, n, sc.ctypes.data, pts.strides[0]))
                else:
                    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
                    t = best(lambda: f(out.ctypes.data, pts.ctypes.data, n, sc.ctypes.data))
                line += f"   reference {t:8.1f} ms   ratio {t / ours:5.2f}x"
            print(line, flush=True)


if __name__ == "__main__":
    main()
