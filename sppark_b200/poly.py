"""Host-side mirror of the reference's polynomial helpers (polynomial/prefix_op.cuh:322,
polynomial/div_by_x_minus_z.cuh:445, polynomial/evaluate.cuh:308) and of ff/batch_inversion.hpp,
through the C-ABI `sppark_b200_*_dev` entry points (include/sppark_b200.h, polynomial block).

The reference's templates take DEVICE arrays and a stream; so do the C entries.  The functions
here accept either torch CUDA tensors (used in place, enqueued on torch's current stream, not
synchronised) or numpy arrays in the field's memory format (copied to the current device and
back -- a convenience for tests and small jobs).  Element type picks the field as in ntt.py:
uint64 = Goldilocks, uint32 = BabyBear Montgomery words; pass `field=` for the 256-bit fields
(arrays of shape (n, 4) uint64 or (n, 8) uint32).
"""
import numpy as np

from . import _lib
from .ntt import GL64, BB31, _field_of

ADD, MULTIPLY = 0, 1
_ELEM_BYTES = {GL64: 8, BB31: 4}


def _elem_bytes(field):
    return _ELEM_BYTES.get(field, 32)


class _Buf:
    """Device view of an argument: a torch CUDA tensor as is, a numpy array through a copy."""

    def __init__(self, a, field, writable=False):
        import torch
        self.np = None
        if isinstance(a, np.ndarray):
            if not a.flags["C_CONTIGUOUS"]:
                raise TypeError("arrays must be C-contiguous")
            self.np = a
            self.t = torch.from_numpy(a.reshape(-1).view(np.uint8)).cuda()
        else:
            if not (a.is_cuda and a.is_contiguous()):
                raise TypeError("tensors must be contiguous CUDA tensors")
            self.t = a
        self.nbytes = self.t.numel() * self.t.element_size()
        if self.nbytes % _elem_bytes(field):
            raise ValueError("array size is not a whole number of field elements")
        self.len = self.nbytes // _elem_bytes(field)
        self.ptr = self.t.data_ptr()
        self.writable = writable

    def back(self):
        if self.np is not None and self.writable:
            self.np.reshape(-1).view(np.uint8)[:] = self.t.cpu().numpy().reshape(-1).view(np.uint8)


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def _pick(a, field):
    if field is not None:
        return field
    if isinstance(a, np.ndarray):
        return _field_of(a)
    raise TypeError("pass field= with tensor arguments")


def prefix_op(op, inout, field=None):
    """inout[i] = inout[0] (op) ... (op) inout[i], op = ADD or MULTIPLY, in place."""
    field = _pick(inout, field)
    b = _Buf(inout, field, writable=True)
    _lib.check(_lib.lib().sppark_b200_prefix_op_dev(field, op, b.ptr, b.ptr, b.len, _stream()))
    b.back()


def div_by_x_minus_z(inout, z, rotate=False, field=None):
    """Divide the polynomial inout[0] + inout[1] x + ... by (x - z) in place; z is one element
    (numpy, memory format).  rotate=False: [remainder, quotient...]; True: [quotient..., remainder]."""
    field = _pick(inout, field)
    z = np.ascontiguousarray(z)
    if z.nbytes != _elem_bytes(field):
        raise ValueError("z must be exactly one field element")
    b = _Buf(inout, field, writable=True)
    _lib.check(_lib.lib().sppark_b200_div_by_x_minus_z_dev(field, b.ptr, b.len, z.ctypes.data, int(rotate), _stream()))
    b.back()


def evaluate(coeffs, x, field=None):
    """ret[k] = sum_i coeffs[i] * x[k]^i; returns an array/tensor shaped like x."""
    import torch
    field = _pick(coeffs, field)
    c, xs = _Buf(coeffs, field), _Buf(x, field)
    ret = torch.empty_like(xs.t)
    _lib.check(_lib.lib().sppark_b200_evaluate_dev(field, ret.data_ptr(), xs.ptr, xs.len, c.ptr, c.len, _stream()))
    if isinstance(x, np.ndarray):
        return ret.cpu().numpy().view(x.dtype).reshape(x.shape)
    return ret


def batch_inverse(inout, field=None):
    """inout[i] = 1 / inout[i] (zero stays zero), in place."""
    field = _pick(inout, field)
    b = _Buf(inout, field, writable=True)
    _lib.check(_lib.lib().sppark_b200_batch_inverse_dev(field, b.ptr, b.ptr, b.len, _stream()))
    b.back()
