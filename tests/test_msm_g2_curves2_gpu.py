"""G2 MSM of the reference's bn254 and bls12_377 builds (SPPARK_CURVE_BN254_G2 / _BLS12_377_G2),
through the C ABI, against the recordings of the reference's own CUDA templates and against the
Python restatement (oracle/g2py.py)."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CURVES = ["bn254", "bls12_377"]


def _curve(name):
    from oracle import g2py
    return g2py.curve(name + "_g2")


@pytest.mark.parametrize("name", CURVES)
def test_matches_the_reference_recordings(name):
    """packed rows and arkworks rows (infinity flag after Y), plain and Montgomery scalars"""
    from sppark_b200 import msm
    c = _curve(name)
    g = np.load(os.path.join(GOLDEN, f"msm_g2_{name}_ref_gpu.npz"))
    for k in range(int(g["ncases"])):
        pts, sc = g[f"points{k}"], g[f"scalars{k}"]
        want = c.jacobian_to_affine(g[f"out{k}"])
        got = msm.msm(c.id, pts, sc)
        assert c.jacobian_to_affine(got) == want, (name, k)
        ark = np.zeros((pts.shape[0], pts.shape[1] + 1), dtype=np.uint64)
        ark[:, :-1] = pts
        inf = ~pts.any(axis=1)
        ark[inf, -1] = 1
        ark[inf, :3] = 7                                    # a flagged row's coordinates are ignored
        assert c.jacobian_to_affine(msm.msm(c.id, ark, sc)) == want, (name, k, "arkworks rows")


@pytest.mark.parametrize("name", CURVES)
def test_matches_oracle_and_linearity(name):
    from oracle import g2py
    from sppark_b200 import msm
    c = _curve(name)
    rnd = random.Random(17)
    base = g2py.multiples(c, 32)
    for n in (1, 3, 50):
        pts = [base[rnd.randrange(32)] for _ in range(n)]
        sc = [rnd.randrange(c.r) for _ in range(n)]
        got = c.jacobian_to_affine(msm.msm(c.id, c.encode_affine(pts), g2py.scalars_to_rows(sc)))
        assert got == c.msm(pts, sc), (name, n)
    # 2^14 terms over 32 distinct points: fold the scalars per point on the CPU
    n = 1 << 14
    idx = [rnd.randrange(32) for _ in range(n)]
    sc = [rnd.randrange(c.r) for _ in range(n)]
    folded = [0] * 32
    for i, k in zip(idx, sc):
        folded[i] = (folded[i] + k) % c.r
    rows = c.encode_affine(base)[np.array(idx)]
    got = c.jacobian_to_affine(msm.msm(c.id, np.ascontiguousarray(rows), g2py.scalars_to_rows(sc)))
    assert got == c.msm(base, folded), name


@pytest.mark.parametrize("name", CURVES)
def test_generated_points_and_device_entry(name):
    """sppark_b200_generate_points_dev writes (i+1) G; the device-pointer MSM agrees with the host one"""
    import torch
    from oracle import g2py
    from sppark_b200 import _lib, msm
    c = _curve(name)
    n = 300
    l, s = _lib.lib(), torch.cuda.current_stream().cuda_stream
    d = torch.zeros(n * 4 * c.nl, dtype=torch.int64, device="cuda")
    _lib.check(l.sppark_b200_generate_points_dev(c.id, d.data_ptr(), n, s))
    torch.cuda.synchronize()
    rows = d.cpu().numpy().view(np.uint64).reshape(n, 4 * c.nl)
    assert c.decode_affine(rows[:40]) == g2py.multiples(c, 40)
    rnd = random.Random(5)
    sc = g2py.scalars_to_rows([rnd.randrange(c.r) for _ in range(n)])
    dsc = torch.from_numpy(sc.view(np.int64)).cuda()
    out = np.zeros(6 * c.nl, dtype=np.uint64)
    _lib.check(l.sppark_b200_msm_dev(c.id, out.ctypes.data, d.data_ptr(), n, dsc.data_ptr(), s))
    assert c.jacobian_to_affine(out) == c.jacobian_to_affine(msm.msm(c.id, rows, sc))
