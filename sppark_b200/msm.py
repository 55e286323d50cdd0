"""Host-side mirror of the reference's msm-cuda crate (poc/msm-cuda/src/lib.rs:18-81):
multi_scalar_mult (blst layout, 96-byte affine points) and multi_scalar_mult_arkworks
(arkworks layout, 104-byte G1Affine with an infinity flag), through the C-ABI.

Arrays: points (n, 12) uint64 [or (n, 13) for the arkworks layout: X, Y, flag word],
scalars (n, 4) uint64 little-endian, result (18,) uint64 = Jacobian {X, Y, Z} in Montgomery form.
"""
import numpy as np

from . import _lib

BLS12_381_G1, PALLAS, VESTA, BLS12_381_G2, BN254_G1, BLS12_377_G1, BN254_G2, BLS12_377_G2 = 0, 1, 2, 3, 4, 5, 6, 7
_LIMBS = {BLS12_381_G1: 6, PALLAS: 4, VESTA: 4, BLS12_381_G2: 12, BN254_G1: 4, BLS12_377_G1: 6, BN254_G2: 8, BLS12_377_G2: 12}    # 64-bit limbs per coordinate


def _check(points, scalars):
    if points.shape[0] != scalars.shape[0]:
        raise ValueError("length mismatch")                  # same panic as the crate (lib.rs:33-35)
    if not (points.flags["C_CONTIGUOUS"] and scalars.flags["C_CONTIGUOUS"]):
        raise TypeError("points and scalars must be C-contiguous")
    if points.dtype != np.uint64 or scalars.dtype != np.uint64 or scalars.shape[1] != 4:
        raise TypeError("points/scalars must be uint64 limb arrays")


def multi_scalar_mult(points, scalars):
    """blst_p1_affine[] x blst_scalar[] -> blst_p1  via mult_pippenger."""
    _check(points, scalars)
    assert points.shape[1] == 12
    out = np.zeros(18, dtype=np.uint64)
    err = _lib.lib().mult_pippenger(out.ctypes.data, points.ctypes.data, points.shape[0], scalars.ctypes.data)
    _lib.check(err)
    return out


def multi_scalar_mult_arkworks(points, scalars):
    """ark G1Affine[] (X, Y, infinity flag; size_of::<G1Affine>() == 104) x BigInteger256[]
    via mult_pippenger_inf."""
    _check(points, scalars)
    assert points.shape[1] == 13
    out = np.zeros(18, dtype=np.uint64)
    err = _lib.lib().mult_pippenger_inf(out.ctypes.data, points.ctypes.data, points.shape[0],
                                        scalars.ctypes.data, points.strides[0])
    _lib.check(err)
    return out


def multi_scalar_mult_fp2_arkworks(points, scalars):
    """ark G2Affine[] (X, Y in Fq2, infinity flag; size_of::<G2Affine>() == 200) x BigInteger256[]
    -> G2Projective (36 limbs) via mult_pippenger_fp2_inf (poc/msm-cuda/src/lib.rs:84-119)."""
    _check(points, scalars)
    assert points.shape[1] == 25
    out = np.zeros(36, dtype=np.uint64)
    err = _lib.lib().mult_pippenger_fp2_inf(out.ctypes.data, points.ctypes.data, points.shape[0],
                                            scalars.ctypes.data, points.strides[0])
    _lib.check(err)
    return out


def msm(curve, points, scalars, mont=False):
    """Any supported curve; host arrays of packed affine points (n, 2*limbs) or rows with an
    infinity-flag word appended (n, 2*limbs + 1).  mont=True: the scalars are Montgomery residues
    (the `mont` flag of the reference's C++ template, msm/pippenger.cuh:730-733)."""
    _check(points, scalars)
    nl = _LIMBS[curve]
    assert points.shape[1] in (2 * nl, 2 * nl + 1)
    out = np.zeros(3 * nl, dtype=np.uint64)
    err = _lib.lib().sppark_b200_msm_ex(curve, out.ctypes.data, points.ctypes.data, points.shape[0],
                                        scalars.ctypes.data, points.strides[0], int(mont))
    _lib.check(err)
    return out


class MsmContext:
    """Points preloaded on the current GPU (the reference's msm_t{points, npoints}): every
    `invoke(scalars)` moves only the scalars.  Host arrays as for msm()."""

    def __init__(self, curve, points):
        import ctypes as C
        if points.dtype != np.uint64 or not points.flags["C_CONTIGUOUS"]:
            raise TypeError("points must be a C-contiguous uint64 limb array")
        nl = _LIMBS[curve]
        assert points.shape[1] in (2 * nl, 2 * nl + 1)
        self.curve, self.npoints, self._h = curve, points.shape[0], C.c_void_p()
        _lib.check(_lib.lib().sppark_b200_msm_ctx_create(curve, points.ctypes.data, points.shape[0],
                                                         points.strides[0], C.byref(self._h)))

    def invoke(self, scalars, mont=False):
        if scalars.dtype != np.uint64 or scalars.ndim != 2 or scalars.shape[1] != 4 or not scalars.flags["C_CONTIGUOUS"]:
            raise TypeError("scalars must be a C-contiguous (n, 4) uint64 array")
        out = np.zeros(3 * _LIMBS[self.curve], dtype=np.uint64)
        _lib.check(_lib.lib().sppark_b200_msm_ctx_invoke(self._h, out.ctypes.data, scalars.ctypes.data,
                                                         scalars.shape[0], int(mont)))
        return out

    def close(self):
        if self._h:
            _lib.lib().sppark_b200_msm_ctx_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def msm_dev(curve, d_points, d_scalars, npoints=None, stream=None):
    """msm_t::invoke with device-resident inputs (msm/pippenger.cuh:582-601): torch CUDA tensors
    of packed affine points / 32-byte scalars; synchronises the stream and returns the Jacobian
    result as a host array."""
    import torch
    nl = _LIMBS[curve]
    assert d_points.is_cuda and d_scalars.is_cuda and d_points.is_contiguous() and d_scalars.is_contiguous()
    n = npoints if npoints is not None else d_scalars.numel() * d_scalars.element_size() // 32
    out = np.zeros(3 * nl, dtype=np.uint64)
    with torch.cuda.device(d_points.device):
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        err = _lib.lib().sppark_b200_msm_dev(curve, out.ctypes.data, d_points.data_ptr(), n,
                                             d_scalars.data_ptr(), s)
    _lib.check(err)
    return out


def generate_points_dev(curve, n, device=None):
    """(n, 2*limbs) int64 CUDA tensor holding (i+1)*G, i < n, as packed Montgomery affine points."""
    import torch
    nl = _LIMBS[curve]
    out = torch.empty((n, 2 * nl), dtype=torch.int64, device=device or "cuda")
    with torch.cuda.device(out.device):
        err = _lib.lib().sppark_b200_generate_points_dev(curve, out.data_ptr(), n,
                                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(err)
    return out


def combine(curve, partials):
    """Sum of Jacobian points: partials (count, 3*limbs) uint64 host array."""
    partials = np.ascontiguousarray(partials, dtype=np.uint64)
    out = np.zeros(3 * _LIMBS[curve], dtype=np.uint64)
    _lib.check(_lib.lib().sppark_b200_msm_combine(curve, out.ctypes.data, partials.ctypes.data, partials.shape[0]))
    return out


def selftest_field(field, op, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    r = np.zeros_like(a)
    err = _lib.lib().sppark_b200_selftest_field(field, {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "mul_shared": 4, "sqr_shared": 5, "msub_shared": 6}[op],
                                                a.shape[0], r.ctypes.data, a.ctypes.data, b.ctypes.data)
    _lib.check(err)
    return r
