"""CPU: oracle/g2py.py (G2 of BN254 and BLS12-377 on Python integers) against the recordings of the
reference's own CUDA templates (tests/golden/msm_g2_<curve>_ref_gpu.npz) and its own algebra."""
import os

import numpy as np
import pytest

from oracle import g2py

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["bn254", "bls12_377"])
def test_oracle_matches_reference_recordings(name):
    c = g2py.curve(name + "_g2")
    g = np.load(os.path.join(GOLDEN, f"msm_g2_{name}_ref_gpu.npz"))
    assert int(g["ncases"]) == 5
    for k in range(4):                                       # n = 1, 2, 33, 200 (1000 terms: the GPU test's job)
        pts = c.decode_affine(g[f"points{k}"])
        sc = [int(sum(int(v) << (64 * j) for j, v in enumerate(row))) for row in g[f"scalars{k}"]]
        assert c.msm(pts, sc) == c.jacobian_to_affine(g[f"out{k}"]), (name, k)


@pytest.mark.parametrize("name", ["bn254_g2", "bls12_377_g2"])
def test_generator_has_order_r_and_the_formats_round_trip(name):
    c = g2py.curve(name)
    assert c.G is not None and c.smul(c.r, c.G) is None
    pts = g2py.multiples(c, 5) + [None]
    assert c.decode_affine(c.encode_affine(pts)) == pts
    assert c.add(pts[0], c.neg(pts[0])) is None
    assert c.msm(pts, [3, 0, 1, c.r - 1, 2, 9]) == c.smul((3 + 3 - 4 + 10) % c.r, c.G)
    # Fp2: u^2 = -beta
    assert c.mul((0, 1), (0, 1)) == (-c.beta % c.p, 0)
    a = (12345, 67890)
    assert c.mul(a, c.inv(a)) == (1, 0)


def test_generator_constants_match_the_device_tables():
    """csrc/ff/fields.cuh (tools/gen_fields.py) and the oracle derive the same base points"""
    import re
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "sppark_b200", "csrc", "ff", "fields.cuh")).read()
    for name in ("bn254_g2", "bls12_377_g2"):
        c = g2py.curve(name)
        body = src[src.index(f"struct {name}_gen"):]
        body = body[:body.index("\n};")]
        words = [int(v, 16) for v in re.findall(r"0x([0-9a-f]{8})u", body)]
        assert words == list(c.encode_affine([c.G])[0].view(np.uint32)), name
