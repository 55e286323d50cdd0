"""GPU probe: host-pointer MSM (mult_pippenger, pinned buffers) under different slice schedules
(SPPARK_B200_MSM_SCHED = relative slice sizes).  Development tool, not the bench."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sppark_b200 import msm  # noqa: E402
from oracle import pyoracle  # noqa: E402  (development tool: result comparison only)


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    scheds = sys.argv[2:] or ["", "1,1,2,4", "1,1,2,4,8", "1,2,4,9", "1,1,2,4,8,16", "1,3,4,8", "2,2,4,8,16"]
    n = 1 << lg
    base = msm.generate_points_dev(msm.BLS12_381_G1, 1 << 16)
    idx = torch.arange(n, device="cuda") % (1 << 16)
    pts_t = torch.empty((n, 12), dtype=torch.int64, pin_memory=True)
    pts_t.copy_(base[idx])
    del idx
    sc_t = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    sc_t.copy_(torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda"))
    torch.cuda.synchronize()
    pts, sc = pts_t.numpy().view(np.uint64), sc_t.numpy().view(np.uint64)
    ref = None
    for s in scheds:
        if s:
            os.environ["SPPARK_B200_MSM_SCHED"] = s
        else:
            os.environ.pop("SPPARK_B200_MSM_SCHED", None)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            out = msm.multi_scalar_mult(pts, sc)
            ts.append((time.perf_counter() - t0) * 1e3)
        out = pyoracle.jac_to_affine('bls12_381', out)
        if ref is None:
            ref = out
        print(f"sched {s or 'default':>14}: min {min(ts[1:]):7.1f} ms  median {sorted(ts[1:])[1]:7.1f} ms  same={np.array_equal(out, ref)}", flush=True)


def resident(lg=26):
    """preloaded points (sppark_b200_msm_ctx_*): only the scalars cross PCIe"""
    n = 1 << lg
    base = msm.generate_points_dev(msm.BLS12_381_G1, 1 << 16)
    idx = torch.arange(n, device="cuda") % (1 << 16)
    pts_t = torch.empty((n, 12), dtype=torch.int64, pin_memory=True)
    pts_t.copy_(base[idx])
    del idx, base
    sc_t = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    sc_t.copy_(torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda"))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx = msm.MsmContext(msm.BLS12_381_G1, pts_t.numpy().view(np.uint64))
    print(f"preload 2^{lg} points: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    sc = sc_t.numpy().view(np.uint64)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        out = ctx.invoke(sc)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"invoke with preloaded points: min {min(ts[1:]):.1f} ms  median {sorted(ts[1:])[1]:.1f} ms", flush=True)
    full = msm.multi_scalar_mult(pts_t.numpy().view(np.uint64), sc)
    print("same point as mult_pippenger:", np.array_equal(pyoracle.jac_to_affine("bls12_381", out), pyoracle.jac_to_affine("bls12_381", full)))
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "resident":
        resident(int(sys.argv[2]) if len(sys.argv) > 2 else 26)
        sys.exit(0)
    main()
