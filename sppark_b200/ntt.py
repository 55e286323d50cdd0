"""Host-side mirror of the reference's ntt-cuda crate (poc/ntt-cuda/src/lib.rs:20-118):
NTT / iNTT / coset_NTT / coset_iNTT, in place on a host array, through the C-ABI
`compute_ntt`-style entry points.  Element type picks the field: uint64 = Goldilocks,
uint32 = BabyBear (Montgomery words), as the reference's per-FEATURE builds do."""
import numpy as np

from . import _lib

NN, NR, RN, RR = 0, 1, 2, 3            # sppark::NTTInputOutputOrder (rust/src/lib.rs:99-105)
BB = 4                                 # extension: bit-reversed in and out (RR == NN in the reference)
FORWARD, INVERSE = 0, 1                # sppark::NTTDirection
STANDARD, COSET = 0, 1                 # sppark::NTTType
GL64, BB31 = 0, 1
BLS12_381_FR, PALLAS_FR, VESTA_FR = 2, 3, 4      # (n, 4) uint64 arrays of Montgomery residues
BN254_FR, BLS12_377_FR = 5, 6                    # likewise (domains up to 2^28 / 2^30)


def _field_of(a):
    if a.dtype == np.uint64:
        return GL64
    if a.dtype == np.uint32:
        return BB31
    raise TypeError("inout must be uint64 (Goldilocks) or uint32 (BabyBear)")


def _run(device_id, inout, order, direction, typ, field=None):
    if not isinstance(inout, np.ndarray) or not inout.flags["C_CONTIGUOUS"] or not inout.flags["WRITEABLE"]:
        raise TypeError("inout must be a writable C-contiguous numpy array")
    n = inout.size if field in (None, GL64, BB31) else inout.shape[0]
    if n & (n - 1):
        raise ValueError("inout.len() is not power of 2")     # same panic text as the crate
    lg = n.bit_length() - 1 if n else 0
    if n == 0:
        return
    field = _field_of(inout) if field is None else field
    l = _lib.lib()
    if field == GL64:
        err = l.compute_ntt(device_id, inout.ctypes.data, lg, order, direction, typ)
    else:
        err = l.sppark_b200_ntt(field, device_id, inout.ctypes.data, lg, order, direction, typ)
    _lib.check(err)


def NTT(device_id, inout, order=NN, field=None):
    _run(device_id, inout, order, FORWARD, STANDARD, field)


def iNTT(device_id, inout, order=NN, field=None):
    _run(device_id, inout, order, INVERSE, STANDARD, field)


def coset_NTT(device_id, inout, order=NN, field=None):
    _run(device_id, inout, order, FORWARD, COSET, field)


def coset_iNTT(device_id, inout, order=NN, field=None):
    _run(device_id, inout, order, INVERSE, COSET, field)


def LDE(device_id, evals, lg_blowup, field=None, want_coefficients=False):
    """NTT::LDE (ntt/ntt.cuh:336-338): evaluations on the 2^lg domain -> evaluations on the coset
    of the 2^(lg+lg_blowup) domain (natural order).  Returns the extended array (and the natural
    -order coefficients if asked, as LDE_aux does)."""
    field = _field_of(evals) if field is None else field
    n = evals.size if field in (GL64, BB31) else evals.shape[0]
    if n & (n - 1):
        raise ValueError("inout.len() is not power of 2")
    lg = n.bit_length() - 1
    shape = (n << lg_blowup,) + tuple(evals.shape[1:])
    ext = np.zeros(shape, dtype=evals.dtype)
    ext[:n] = evals
    aux = np.zeros_like(evals) if want_coefficients else None
    err = _lib.lib().sppark_b200_lde(field, device_id, ext.ctypes.data, lg, lg_blowup,
                                     aux.ctypes.data if aux is not None else None)
    _lib.check(err)
    return (ext, aux) if want_coefficients else ext


def ntt_dev(tensor, order=NN, direction=FORWARD, typ=STANDARD, field=None, stream=None):
    """NTT::Base_dev_ptr (ntt/ntt.cuh:344-350): in place on a CUDA torch tensor, enqueued on
    torch's current stream (or `stream`), not synchronised."""
    import torch
    assert tensor.is_cuda and tensor.is_contiguous()
    n = tensor.numel()
    if n & (n - 1):
        raise ValueError("inout.len() is not power of 2")
    if field is None:
        field = {8: GL64, 4: BB31}[tensor.element_size()]
    with torch.cuda.device(tensor.device):
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        err = _lib.lib().sppark_b200_ntt_dev(field, tensor.data_ptr(), n.bit_length() - 1,
                                             order, direction, typ, s)
    _lib.check(err)
