"""GPU probe: time our NTT (device-resident) against the reference's own kernels built for
sm_100a (oracle/_ref), over the kernel variants selected by environment knobs.  Development
tool, not the bench."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sppark_b200 import ntt, _lib  # noqa: E402

GL_P = 2**64 - 2**32 + 1


def time_fn(fn, stream, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    stream.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


VARIANTS = [("warp cpt1", {"SPPARK_B200_NTT_CPT": "1"}), ("warp cpt2", {"SPPARK_B200_NTT_CPT": "2"}),
            ("block", {"SPPARK_B200_NTT_BLOCK": "1"})]
KNOBS = ["SPPARK_B200_NTT_CPT", "SPPARK_B200_NTT_BLOCK", "SPPARK_B200_NTT_SPLIT"]


def main():
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    cur = torch.cuda.current_stream()
    flush = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")      # 256 MiB > L2
    sizes = [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24]
    for lg in sizes:
        n = 1 << lg
        rng = np.random.default_rng(lg)
        host = rng.integers(0, GL_P, size=n, dtype=np.uint64)
        bytes_alg = 2 * n * 8
        extra = []
        if lg == 20:
            extra = [("warp 7,7,6", {"SPPARK_B200_NTT_SPLIT": "7,7,6"}), ("warp 8,8,4", {"SPPARK_B200_NTT_SPLIT": "8,8,4"}),
                     ("warp 8,6,6", {"SPPARK_B200_NTT_SPLIT": "8,6,6"}), ("warp 5,5,5,5", {"SPPARK_B200_NTT_SPLIT": "5,5,5,5"})]
        for vname, env in VARIANTS + extra:
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(env)
            for order, name in ((ntt.NN, "NN"), (ntt.NR, "NR"), (ntt.RN, "RN")):
                d = torch.from_numpy(host.view(np.int64)).cuda()
                med, best = time_fn(lambda: ntt.ntt_dev(d, order), cur, flush=flush if lg >= 22 else None)
                print(f"ours[{vname}] gl64 2^{lg} {name}: median {med*1e3:.1f} us  min {best*1e3:.1f} us  "
                      f"-> {bytes_alg/ (med*1e-3) / 1e9:.0f} GB/s algorithmic", flush=True)
        for k in KNOBS:
            os.environ.pop(k, None)
        # reference kernels, same box, data resident
        p = os.path.join(ROOT, "oracle", "_ref", "libref_ntt_gl64_gpu.so")
        if os.path.exists(p):
            ref = C.CDLL(p)
            ref.ref_ntt_stream.restype = C.c_void_p
            ref.ref_ntt_dev_async.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
            rs = torch.cuda.ExternalStream(ref.ref_ntt_stream())
            for order, name in ((0, "NN"), (1, "NR"), (2, "RN")):
                d = torch.from_numpy(host.view(np.int64)).cuda()
                torch.cuda.synchronize()
                med, best = time_fn(lambda: ref.ref_ntt_dev_async(d.data_ptr(), lg, order, 0, 0), rs,
                                    flush=flush if lg >= 22 else None)
                print(f"REF   gl64 2^{lg} {name}: median {med*1e3:.1f} us  min {best*1e3:.1f} us  "
                      f"-> {bytes_alg/ (med*1e-3) / 1e9:.0f} GB/s algorithmic", flush=True)
            # parity of ours vs the reference GPU implementation at this size
            d1 = torch.from_numpy(host.view(np.int64)).cuda()
            d2 = d1.clone()
            ntt.ntt_dev(d1, ntt.NN)
            torch.cuda.synchronize()
            ref.ref_ntt_dev_async(d2.data_ptr(), lg, 0, 0, 0)
            rs.synchronize()
            print(f"ours == reference-GPU at 2^{lg} NN:", bool(torch.equal(d1, d2)), flush=True)
    print("launches:", _lib.launch_count())


if __name__ == "__main__":
    main()
