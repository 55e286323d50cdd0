"""Single-process multi-GPU entry points of the C ABI (sppark_b200_ntt_sharded / _msm_sharded) on
every visible device: timings, and a target for ncu (peer stores of the fused exchange)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sppark_b200 import msm, ntt, parallel
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = 1 << (min(torch.cuda.device_count(), 8).bit_length() - 1)
ids = list(range(g))
rng = np.random.default_rng(0)
x = rng.integers(0, 2**64 - 2**32 + 1, size=1 << lg, dtype=np.uint64)
want = x.copy()
ntt.NTT(0, want, ntt.NN)
for route in ("fused", "copy"):
    if route == "copy":
        os.environ["SPPARK_B200_NTT_EXCHANGE_COPY"] = "1"
    y = x.copy()
    parallel.ntt_sharded_c(y, ids)
    ok = np.array_equal(y, want)
    ts = []
    for _ in range(reps):
        y = x.copy()
        t = time.perf_counter()
        parallel.ntt_sharded_c(y, ids)
        ts.append(time.perf_counter() - t)
    print(f"ntt_sharded gl64 2^{lg} on {g} GPUs, {route} exchange: {min(ts)*1e3:.2f} ms host-to-host (1 GPU compute_ntt: see bench), equal to single-GPU result: {ok}")
os.environ.pop("SPPARK_B200_NTT_EXCHANGE_COPY", None)
if len(sys.argv) > 3:
    n = 1 << int(sys.argv[3])
    base = msm.generate_points_dev(0, 1 << 12).cpu().numpy().view(np.uint64)
    pts = np.tile(base, (n >> 12, 1))
    sc = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); sc[:, 3] >>= np.uint64(2)
    one = msm.multi_scalar_mult(pts, sc)
    for k in (1, g):
        t = time.perf_counter(); r = parallel.msm_sharded_c(0, pts, sc, list(range(k))); dt = time.perf_counter() - t
        t = time.perf_counter(); r = parallel.msm_sharded_c(0, pts, sc, list(range(k))); dt = time.perf_counter() - t
        print(f"msm_sharded 2^{int(sys.argv[3])} on {k} GPU(s): {dt*1e3:.1f} ms (pageable host arrays)")
