"""GPU probe: time our NTT (device-resident) against the reference's own kernels built for
sm_100a (oracle/_ref), over a few pass splits.  Development tool, not the bench."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sppark_b200 import ntt, _lib  # noqa: E402

GL_P = 2**64 - 2**32 + 1


def time_fn(fn, stream, iters=20, warm=3):
    for _ in range(warm):
        fn()
    stream.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    cur = torch.cuda.current_stream()
    for lg in (20, 24):
        n = 1 << lg
        rng = np.random.default_rng(lg)
        host = rng.integers(0, GL_P, size=n, dtype=np.uint64)
        bytes_alg = 2 * n * 8
        for split in (None, "13", "12"):
            if split:
                os.environ["SPPARK_B200_NTT_LG_TILE"] = split
            else:
                os.environ.pop("SPPARK_B200_NTT_LG_TILE", None)
            for order, name in ((ntt.NN, "NN"), (ntt.NR, "NR"), (ntt.RN, "RN")):
                d = torch.from_numpy(host.view(np.int64)).cuda()
                med, best = time_fn(lambda: ntt.ntt_dev(d, order), cur)
                print(f"ours  gl64 2^{lg} {name} split={split}: median {med*1e3:.1f} us  min {best*1e3:.1f} us  "
                      f"-> {bytes_alg/ (med*1e-3) / 1e9:.0f} GB/s algorithmic")
        os.environ.pop("SPPARK_B200_NTT_LG_TILE", None)
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
                med, best = time_fn(lambda: ref.ref_ntt_dev_async(d.data_ptr(), lg, order, 0, 0), rs)
                print(f"REF   gl64 2^{lg} {name}: median {med*1e3:.1f} us  min {best*1e3:.1f} us  "
                      f"-> {bytes_alg/ (med*1e-3) / 1e9:.0f} GB/s algorithmic")
            # parity of ours vs the reference GPU implementation at this size
            d1 = torch.from_numpy(host.view(np.int64)).cuda()
            d2 = d1.clone()
            ntt.ntt_dev(d1, ntt.NN)
            torch.cuda.synchronize()
            ref.ref_ntt_dev_async(d2.data_ptr(), lg, 0, 0, 0)
            rs.synchronize()
            print(f"ours == reference-GPU at 2^{lg} NN:", bool(torch.equal(d1, d2)))
    print("launches:", _lib.launch_count())


if __name__ == "__main__":
    main()
