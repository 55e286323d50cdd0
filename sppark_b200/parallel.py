"""Multi-GPU sharding of the two hot paths: one process per GPU, `torch.distributed` for the
plumbing (NCCL on GPUs, gloo in the CPU tests).

MSM is linear in its points: rank g takes the contiguous chunk [g*N/G, (g+1)*N/G), reduces it to
ONE Jacobian point with the single-GPU pipeline, the 144-byte partials are all-gathered and added
(`sppark_b200_msm_combine`; NCCL's reductions cannot add curve points).  The reference has no
multi-GPU path ("Parallelism inventory"); this is new functionality behind the same
per-GPU entry points.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous chunk of [0, n) owned by `rank`; chunks differ by at most one element and are
    multiples of nothing in particular (the MSM kernels accept any length, including 0)."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def all_gather_partials(partial, device=None):
    """partial: (k,) uint64 host array -> (world, k) uint64, same order on every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(partial, dtype=np.uint64)[None, :].copy()
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy().view(np.uint64)


def msm_sharded(local_msm, combine, partial_words, device=None):
    """Run `local_msm()` (this rank's chunk -> Jacobian words), exchange, and fold with
    `combine((world, k) array) -> (k,)`.  Every rank returns the full result."""
    part = np.asarray(local_msm(), dtype=np.uint64)
    assert part.shape == (partial_words,)
    return combine(all_gather_partials(part, device))


# ---------------------------------------------------------------------------------- NTT
_MAX_LG_R = {0: 12, 1: 12, 2: 11, 3: 11, 4: 11, 5: 11, 6: 11}       # F::NTT_MAX_LG_R per field id (csrc/ff/*.cuh)


def slab_first_digit(lg_n, field=0, s1=None):
    """log2(N1): the first digit of the planner's split (ntt_plan.hpp: slab_first_digit)."""
    if s1 is not None:
        return s1
    max_r = _MAX_LG_R[field]
    p = max(1, -(-lg_n // max_r))
    return (lg_n + 1) // 2 if p < 2 else lg_n // p + (1 if lg_n % p else 0)


def slab_shapes(lg_n, lg_g, field=0, s1=None):
    """(N1, N2, local input shape [N1][N2/G], local output shape [N2][N1/G]) of the slab-sharded
    transform: x[j1*N2 + j2] lives on the rank owning column j2, X[k1 + N1*k2] on the rank owning k1.
    N1 = 2^(first digit): half of lg_n while both halves fit one tile (<= 2^12 rows), otherwise
    the first of three (or more) digits, e.g. BabyBear 2^27 = 2^9 x 2^18."""
    s1 = slab_first_digit(lg_n, field, s1)
    s2 = lg_n - s1
    n1, n2, g = 1 << s1, 1 << s2, 1 << lg_g
    return n1, n2, (n1, n2 // g), (n2, n1 // g)


def scatter_columns(x, lg_n, lg_g, rank, field=0, s1=None):
    """This rank's slab of a full natural-order array (test/bench helper)."""
    n1, n2, (_, c), _ = slab_shapes(lg_n, lg_g, field, s1)
    return np.ascontiguousarray(x.reshape(n1, n2)[:, rank * c:(rank + 1) * c])


def gather_columns(parts, lg_n, lg_g, field=0, s1=None):
    """Inverse of the output distribution: parts[r] = [N2][N1/G] -> natural-order array."""
    n1, n2, _, (_, d) = slab_shapes(lg_n, lg_g, field, s1)
    full = np.concatenate([np.asarray(p).reshape(n2, d) for p in parts], axis=1)      # [k2][k1]
    return np.ascontiguousarray(full.reshape(-1))


def exchange(staging, world, all_to_all):
    """The one collective of the sharded NTT: block q of `staging` ([G][...]) goes to rank q.
    `all_to_all(recv, send)` is torch.distributed.all_to_all_single on NCCL; gloo (CPU tests) has
    no all-to-all, so there the blocks travel as point-to-point messages."""
    import torch
    import torch.distributed as dist
    recv = torch.empty_like(staging)
    if world == 1:
        recv.copy_(staging)
    elif all_to_all:
        dist.all_to_all_single(recv, staging)
    else:
        rank = dist.get_rank()
        sb, rb = staging.view(world, -1), recv.view(world, -1)
        rb[rank].copy_(sb[rank])
        reqs = []
        for q in range(world):
            if q != rank:
                reqs.append(dist.isend(sb[q].contiguous(), q))
                reqs.append(dist.irecv(rb[q], q))
        for r in reqs:
            r.wait()
    return recv


def ntt_slab(local, lg_n, field, pass_fn, inverse=False, all_to_all=True):
    """Slab-sharded NTT of 2^lg_n points over the default process group.  `local` is this rank's
    [N1][N2/G] slab (torch tensor, flat); returns its [N2][N1/G] slab of the result.
    pass_fn(which, src, dst) runs one local stage (sppark_b200_ntt_slab_pass on a GPU): which = 1
    writes dst, which = 2 transforms src with dst as scratch."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    staging = torch.empty_like(local)
    pass_fn(1, local, staging)
    recv = exchange(staging, world, all_to_all)
    pass_fn(2, recv, staging)            # staging doubles as scratch when N2 takes several passes
    return recv


def gpu_slab_pass(field, lg_n, lg_g, rank, inverse=False):
    """pass_fn for ntt_slab() on CUDA tensors, on torch's current stream."""
    import torch
    from . import _lib

    def run(which, src, dst):
        with torch.cuda.device(src.device):
            err = _lib.lib().sppark_b200_ntt_slab_pass(field, which, src.data_ptr(), dst.data_ptr(), lg_n, lg_g,
                                                       rank, int(inverse), torch.cuda.current_stream().cuda_stream)
        _lib.check(err)
    return run


# ------------------------------------------------------------ fused exchange over peer memory
class _DevMem:
    """Exposes a raw device allocation to torch (torch.as_tensor aliases it, no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class SlabPeers:
    """Receive buffers of the slab-sharded NTT, mapped into every rank of the node.

    Each rank cudaMallocs `nbuf` receive buffers of `nbytes` (its [N2][N1/G] slab), the CUDA IPC
    handles are all-gathered over the process group and opened by the other ranks, so stage 1 can
    store its output rows directly into the rank that owns them (sppark_b200_ntt_slab_pass_p2p):
    the all-to-all disappears into the store phase of the kernel and overlaps its arithmetic.
    Buffers alternate between calls, so a rank may start the next transform while a peer still
    reads the previous result (see ntt_slab_p2p)."""

    def __init__(self, nbytes, nbuf=2, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib, self.nbytes, self.nbuf = _lib, nbytes, nbuf
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local, self.opened, handles, err = [], [], [], None
        # every rank takes part in every collective below, whatever happened to it locally: a
        # failure (no IPC support, no peer access) is agreed on and raised by ALL ranks together
        try:
            for _ in range(nbuf):
                ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
                _lib.check(_lib.lib().sppark_b200_peer_alloc(nbytes, C.byref(ptr), h))
                self.local.append(ptr.value)
                handles.append(bytes(h))
        except Exception as e:
            err, handles = f"rank {self.rank}: {e}", None
        everyone = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(everyone, handles, group=group)
        else:
            everyone[0] = handles
        self.ptrs = []                                   # ptrs[k] = ctypes array of `world` pointers
        if all(h is not None for h in everyone):
            try:
                for k in range(nbuf):
                    arr = (C.c_void_p * max(self.world, 1))()
                    for q in range(self.world):
                        if q == self.rank:
                            arr[q] = self.local[k]
                        else:
                            p = C.c_void_p()
                            _lib.check(_lib.lib().sppark_b200_peer_open(everyone[q][k], C.byref(p)))
                            self.opened.append(p.value)
                            arr[q] = p.value
                    self.ptrs.append(arr)
            except Exception as e:
                err = f"rank {self.rank}: {e}"
        elif err is None:
            err = "a peer could not allocate its receive buffer"
        status = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(status, err, group=group)     # doubles as the barrier: every mapping exists
        else:
            status[0] = err
        self._views = []
        if any(st is not None for st in status):
            self.close()
            raise RuntimeError("SlabPeers: " + "; ".join(st for st in status if st is not None))
        self.flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.turn = 0
        self._views = [torch.as_tensor(_DevMem(p, nbytes), device="cuda") for p in self.local]

    def tensor(self, k, dtype):
        return self._views[k].view(dtype)

    def close(self):
        import torch
        torch.cuda.synchronize()
        for p in self.opened:
            self._lib.lib().sppark_b200_peer_close(p)
        self.opened = []
        self._views = []
        for p in self.local:
            self._lib.lib().sppark_b200_peer_free(p)
        self.local = []


def ntt_slab_p2p(local, lg_n, field, rank, peers, scratch, inverse=False):
    """Slab-sharded NTT with the exchange fused into stage 1 (NVLink peer stores).  `local`: this
    rank's [N1][N2/G] slab (CUDA tensor); returns its [N2][N1/G] slab of the result, a view of one
    of `peers`' buffers, valid until the call after next.  `scratch`: a tensor like `local`, used
    only when N2 takes several passes.

    Ordering: the tiny all-reduce after stage 1 completes on a rank only when every rank's stage 1
    (stream-ordered before its contribution) has finished, i.e. when all rows have landed; the
    receive buffers alternate, and reaching call k+2 implies every peer has passed the all-reduce
    of call k+1, which it enqueued after its stage 2 of call k -- so buffer k%2 is free again."""
    import torch
    import torch.distributed as dist
    from . import _lib
    lg_g = max(peers.world.bit_length() - 1, 0)
    k = peers.turn
    peers.turn = (k + 1) % peers.nbuf
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().sppark_b200_ntt_slab_pass_p2p(field, local.data_ptr(), peers.ptrs[k], lg_n, lg_g, rank,
                                                         int(inverse), stream))
    if peers.world > 1:
        dist.all_reduce(peers.flag)
    recv = peers.tensor(k, local.dtype)
    _lib.check(_lib.lib().sppark_b200_ntt_slab_pass(field, 2, recv.data_ptr(), scratch.data_ptr(), lg_n, lg_g, rank,
                                                     int(inverse), stream))
    return recv


# ------------------------------------------------------- single-process entry points of the C ABI
def msm_sharded_c(curve, points, scalars, device_ids, mont=False):
    """sppark_b200_msm_sharded: one MSM cut into len(device_ids) point-chunks, chunk i on GPU
    device_ids[i] of THIS process (host arrays in, Jacobian words out)."""
    import ctypes as C
    from . import _lib, msm
    points = np.ascontiguousarray(points)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    ids = (C.c_int * len(device_ids))(*device_ids)
    out = np.zeros(3 * msm._LIMBS[curve], dtype=np.uint64)
    stride = points.strides[0]
    _lib.check(_lib.lib().sppark_b200_msm_sharded(curve, out.ctypes.data, points.ctypes.data, points.shape[0],
                                                  scalars.ctypes.data, stride, int(mont), ids, len(device_ids)))
    return out


def ntt_sharded_c(inout, device_ids, inverse=False, field=None):
    """sppark_b200_ntt_sharded: in-place NN transform of a host array, slab-sharded over the GPUs
    device_ids of THIS process (1, 2, 4 or 8 of them)."""
    import ctypes as C
    from . import _lib, ntt
    if field is None:
        field = ntt._field_of(inout)
    n = inout.size if field in (ntt.GL64, ntt.BB31) else inout.shape[0]
    ids = (C.c_int * len(device_ids))(*device_ids)
    _lib.check(_lib.lib().sppark_b200_ntt_sharded(field, inout.ctypes.data, n.bit_length() - 1, int(inverse), ids, len(device_ids)))
