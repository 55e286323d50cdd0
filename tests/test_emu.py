"""CPU single-steppers (tests/emu): the exact HD kernel bodies of the CUDA path, executed
phase by phase on the host and compared with the oracle.  Validates the NTT planner / butterfly
schedule and the whole MSM pipeline logic without a GPU.  (The emulators are test
infrastructure; they are not part of libsppark_b200.so.)"""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
R_BLS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _build(name):
    so, src = os.path.join(EMU, f"lib{name}.so"), os.path.join(EMU, f"{name}.cpp")
    csrc = os.path.join(os.path.dirname(EMU), "..", "sppark_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(csrc) for f in fs)
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def ntt_emu():
    l = _build("ntt_emu")
    l.emu_ntt_gl64.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_uint]
    l.emu_ntt_bb31.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_uint]
    l.emu_ntt_256.argtypes = [C.c_int, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_uint]
    l.emu_ntt_slab_gl64.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_uint]
    l.emu_ntt_slab_bb31.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_uint]
    l.emu_slab_first_digit.argtypes = [C.c_uint, C.c_uint]
    l.emu_ntt_slab_p2p_gl64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_uint]
    l.emu_ntt_slab_p2p_bb31.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_uint]
    return l


@pytest.fixture(scope="module")
def msm_emu():
    l = _build("msm_emu")
    l.emu_msm_bls12_381.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_uint]
    l.emu_msm_bls12_381_sliced.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
    l.emu_msm_pallas.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_uint]
    l.emu_fp_op.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    l.emu_msm_bls12_381_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
    l.emu_msm_pallas_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
    return l


CASES = [(lg, None) for lg in list(range(1, 15)) + [16]] + [
    (6, "2,2,2"), (6, "1,5"), (6, "5,1"), (9, "3,3,3"), (9, "4,2,3"), (9, "1,1,7"), (9, "7,1,1"),
    (9, "2,2,2,3"), (13, "5,4,4"), (13, "6,7"), (10, "4,3,3"), (10, "3,7")]


@pytest.mark.parametrize("lg,split", CASES)
def test_ntt_plan_and_schedule(oracle, ntt_emu, lg, split, monkeypatch):
    if split:
        monkeypatch.setenv("SPPARK_B200_NTT_SPLIT", split)
    else:
        monkeypatch.delenv("SPPARK_B200_NTT_SPLIT", raising=False)
    rng = np.random.default_rng(lg)
    x = rng.integers(0, 2**64 - 2**32 + 1, size=1 << lg, dtype=np.uint64)
    xb = rng.integers(0, 0x78000001, size=1 << lg, dtype=np.uint32)
    for order in range(5):
        for inv in (0, 1):
            for lg_tile in (14, 7):
                y = x.copy()
                ntt_emu.emu_ntt_gl64(y.ctypes.data, lg, order, inv, lg_tile)
                assert np.array_equal(y, oracle.ntt_gl64(x, order, bool(inv))), (order, inv, lg_tile)
            yb = xb.copy()
            ntt_emu.emu_ntt_bb31(yb.ctypes.data, lg, order, inv, 14)
            assert np.array_equal(yb, oracle.ntt_bb31(xb, order, bool(inv))), (order, inv)


@pytest.mark.parametrize("fid,name", [(2, "bls12_381_fr"), (3, "vesta_fp"), (4, "pallas_fp")])
def test_ntt_256bit_fields(oracle, ntt_emu, fid, name):
    """the same pass kernel with 4 elements per thread over the 256-bit Montgomery fields
    (field ids as in include/sppark_b200.h: the Pallas feature's fr is Vesta's base field)"""
    rnd = random.Random(fid)
    p = oracle.ff_consts(name)["p"]
    for lg in (1, 2, 5, 8, 11, 12):
        x = np.array([oracle.int_to_limbs(rnd.randrange(p), 4) for _ in range(1 << lg)], dtype=np.uint64)
        for order in range(5):
            for inv in (0, 1):
                y = x.copy()
                ntt_emu.emu_ntt_256(fid, y.ctypes.data, lg, order, inv, 11 if lg != 8 else 5)
                assert np.array_equal(y, oracle.ntt_ff(name, x, order, bool(inv))), (lg, order, inv)


@pytest.mark.parametrize("lg,lg_g,lg_tile", [(4, 1, 14), (6, 2, 14), (8, 3, 6), (10, 3, 14), (12, 2, 8),
                                             (14, 3, 14), (7, 1, 5), (6, 3, 14), (2, 1, 14)])
def test_ntt_slab_sharded(oracle, ntt_emu, lg, lg_g, lg_tile):
    """G ranks simulated in one process: local pass 1 -> all-to-all (array slicing) -> local pass 2
    must equal the single-array NN transform."""
    from sppark_b200 import parallel
    G = 1 << lg_g
    rng = np.random.default_rng(lg * 8 + lg_g)
    for field, fn, ofn, dt, p in (("gl64", ntt_emu.emu_ntt_slab_gl64, oracle.ntt_gl64, np.uint64, 2**64 - 2**32 + 1),
                                  ("bb31", ntt_emu.emu_ntt_slab_bb31, oracle.ntt_bb31, np.uint32, 0x78000001)):
        x = rng.integers(0, p, size=1 << lg, dtype=dt)
        for inv in (0, 1):
            stag = []
            for r in range(G):
                loc = parallel.scatter_columns(x, lg, lg_g, r).reshape(-1).copy()
                st = np.zeros_like(loc)
                assert fn(1, loc.ctypes.data, st.ctypes.data, lg, lg_g, r, inv, lg_tile) == 0
                stag.append(st.reshape(G, -1))
            outs = []
            for r in range(G):
                recv = np.concatenate([stag[q][r] for q in range(G)]).copy()
                fn(2, recv.ctypes.data, recv.ctypes.data, lg, lg_g, r, inv, lg_tile)
                outs.append(recv)
            assert np.array_equal(parallel.gather_columns(outs, lg, lg_g), ofn(x, 0, bool(inv))), (field, inv)


@pytest.mark.parametrize("lg,lg_g,split,lg_tile", [
    (8, 1, "3,3,2", 14), (8, 3, "3,3,2", 5), (9, 2, "3,3,3", 14), (9, 3, "4,2,3", 6), (10, 2, "4,3,3", 14),
    (8, 2, "2,2,2,2", 14), (11, 3, "3,4,4", 7), (12, 3, "3,1,8", 14), (14, 3, "5,5,4", 14), (9, 3, "3,3,3", 14)])
def test_ntt_slab_sharded_multipass(oracle, ntt_emu, lg, lg_g, split, lg_tile, monkeypatch):
    """Sizes whose second factor N2 does not fit one tile (BabyBear 2^27 = 2^9 x 2^18 on the GPU):
    after the all-to-all the N2-point column NTTs are the remaining digits of the NN schedule,
    ping-ponging between the received buffer and a scratch buffer.  Small digits stand in for
    the 2^12-row tiles here."""
    from sppark_b200 import parallel
    monkeypatch.setenv("SPPARK_B200_NTT_SPLIT", split)
    s1 = int(split.split(",")[0])
    assert ntt_emu.emu_slab_first_digit(lg, 12) == s1
    G = 1 << lg_g
    rng = np.random.default_rng(lg * 16 + lg_g)
    for field, fn, ofn, dt, p in (("gl64", ntt_emu.emu_ntt_slab_gl64, oracle.ntt_gl64, np.uint64, 2**64 - 2**32 + 1),
                                  ("bb31", ntt_emu.emu_ntt_slab_bb31, oracle.ntt_bb31, np.uint32, 0x78000001)):
        x = rng.integers(0, p, size=1 << lg, dtype=dt)
        for inv in (0, 1):
            stag = []
            for r in range(G):
                loc = parallel.scatter_columns(x, lg, lg_g, r, s1=s1).reshape(-1).copy()
                st = np.zeros_like(loc)
                assert fn(1, loc.ctypes.data, st.ctypes.data, lg, lg_g, r, inv, lg_tile) == 0
                stag.append(st.reshape(G, -1))
            outs = []
            for r in range(G):
                recv = np.concatenate([stag[q][r] for q in range(G)]).copy()
                scratch = np.zeros_like(recv)
                assert fn(2, recv.ctypes.data, recv.ctypes.data, lg, lg_g, r, inv, lg_tile) == -2   # needs scratch
                assert fn(2, recv.ctypes.data, scratch.ctypes.data, lg, lg_g, r, inv, lg_tile) == 0
                outs.append(recv)
            assert np.array_equal(parallel.gather_columns(outs, lg, lg_g, s1=s1), ofn(x, 0, bool(inv))), (field, inv)


@pytest.mark.parametrize("lg,lg_g,split,lg_tile", [(4, 1, None, 14), (6, 2, None, 14), (8, 3, None, 6), (10, 3, None, 14),
                                                   (6, 3, None, 14), (2, 1, None, 14), (12, 0, None, 14),
                                                   (9, 3, "3,3,3", 14), (11, 2, "3,4,4", 7)])
def test_ntt_slab_fused_exchange(oracle, ntt_emu, lg, lg_g, split, lg_tile, monkeypatch):
    """Stage 1 storing its rows directly into the receivers' buffers (the NVLink peer-memory path
    of sppark_b200_ntt_slab_pass_p2p; here the "peers" are G host arrays) must leave in every
    receive buffer exactly what the staging + all-to-all route delivers."""
    from sppark_b200 import parallel
    if split:
        monkeypatch.setenv("SPPARK_B200_NTT_SPLIT", split)
    else:
        monkeypatch.delenv("SPPARK_B200_NTT_SPLIT", raising=False)
    s1 = int(split.split(",")[0]) if split else None
    G = 1 << lg_g
    rng = np.random.default_rng(lg * 32 + lg_g)
    for fn, p2p, ofn, dt, p in ((ntt_emu.emu_ntt_slab_gl64, ntt_emu.emu_ntt_slab_p2p_gl64, oracle.ntt_gl64, np.uint64, 2**64 - 2**32 + 1),
                                (ntt_emu.emu_ntt_slab_bb31, ntt_emu.emu_ntt_slab_p2p_bb31, oracle.ntt_bb31, np.uint32, 0x78000001)):
        x = rng.integers(0, p, size=1 << lg, dtype=dt)
        recv = [np.zeros((1 << lg) // G, dtype=dt) for _ in range(G)]
        ptrs = (C.c_void_p * G)(*[r.ctypes.data for r in recv])
        for r in range(G):
            loc = parallel.scatter_columns(x, lg, lg_g, r, s1=s1).reshape(-1).copy()
            assert p2p(loc.ctypes.data, ptrs, lg, lg_g, r, 0, lg_tile) == 0
        outs = []
        for r in range(G):
            scratch = np.zeros_like(recv[r])
            assert fn(2, recv[r].ctypes.data, scratch.ctypes.data, lg, lg_g, r, 0, lg_tile) == 0
            outs.append(recv[r])
        assert np.array_equal(parallel.gather_columns(outs, lg, lg_g, s1=s1), ofn(x, 0, False))


def test_slab_first_digit_matches_planner(ntt_emu, monkeypatch):
    from sppark_b200 import parallel
    monkeypatch.delenv("SPPARK_B200_NTT_SPLIT", raising=False)
    for max_r, field in ((12, 0), (11, 2)):
        for lg in range(2, 33):
            assert parallel.slab_first_digit(lg, field) == ntt_emu.emu_slab_first_digit(lg, max_r), (lg, max_r)


def _scalars(vals):
    return np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(4)] for v in vals], dtype=np.uint64).reshape(len(vals), 4)


def _run(oracle, emu, pts, sc, wbits, heavy):
    out = np.zeros(18, dtype=np.uint64)
    emu.emu_msm_bls12_381(out.ctypes.data, pts.ctypes.data, pts.shape[0], sc.ctypes.data, wbits, heavy)
    want = oracle.msm("bls12_381", pts, sc, "pippenger", ncpus=4)
    return np.array_equal(oracle.jac_to_affine("bls12_381", out), oracle.jac_to_affine("bls12_381", want))


@pytest.mark.parametrize("n,wbits,heavy", [(1, 0, 0), (2, 0, 0), (33, 0, 0), (300, 0, 0), (300, 5, 0),
                                           (300, 9, 3), (800, 7, 4), (500, 4, 2), (64, 13, 0)])
def test_msm_pipeline_logic(oracle, msm_emu, n, wbits, heavy):
    rnd = random.Random(n * 31 + wbits)
    pts = oracle.gen_points("bls12_381", 64)[np.arange(n) % 64].copy()
    if n > 3:
        pts[3] = 0
    assert _run(oracle, msm_emu, pts, _scalars([rnd.randrange(R_BLS) for _ in range(n)]), wbits, heavy)


@pytest.mark.parametrize("val", [0, 1, R_BLS - 1, (1 << 254) + 12345, 0x8000000080000000800000008000])
def test_msm_pipeline_adversarial_scalars(oracle, msm_emu, val):
    n = 150
    pts = oracle.gen_points("bls12_381", 4)[np.arange(n) % 4].copy()
    p = oracle.ff_consts("bls12_381_fp")["p"]
    y1 = sum(int(v) << (64 * i) for i, v in enumerate(pts[1][6:]))
    pts[5][6:] = [((p - y1) >> (64 * i)) & (2**64 - 1) for i in range(6)]     # a (P, -P) pair in one bucket
    assert _run(oracle, msm_emu, pts, _scalars([val] * n), 6, 8)
    assert _run(oracle, msm_emu, pts, _scalars([val] * n), 5, 0)


@pytest.mark.parametrize("n,wbits,heavy,nslices", [(300, 6, 0, 3), (501, 5, 3, 4), (64, 9, 2, 2), (40, 4, 0, 40)])
def test_msm_sliced_bucket_merging(oracle, msm_emu, n, wbits, heavy, nslices):
    """slices of points folded into persistent buckets (the host-pointer pipeline)"""
    rnd = random.Random(n + nslices)
    pts = oracle.gen_points("bls12_381", 16)[np.arange(n) % 16].copy()
    pts[3] = 0
    sc = _scalars([rnd.randrange(R_BLS) for _ in range(n)])
    sc[: n // 3] = sc[0]
    out = np.zeros(18, dtype=np.uint64)
    msm_emu.emu_msm_bls12_381_sliced(out.ctypes.data, pts.ctypes.data, n, sc.ctypes.data, wbits, heavy, nslices)
    want = oracle.msm("bls12_381", pts, sc, "pippenger", ncpus=4)
    assert np.array_equal(oracle.jac_to_affine("bls12_381", out), oracle.jac_to_affine("bls12_381", want))


def _run_pair(oracle, emu, pts, sc, wbits, heavy, nslices=1, curve="bls12_381"):
    nl = 6 if curve == "bls12_381" else 4
    out = np.zeros(3 * nl, dtype=np.uint64)
    fn = emu.emu_msm_bls12_381_pair if curve == "bls12_381" else emu.emu_msm_pallas_pair
    fn(out.ctypes.data, pts.ctypes.data, pts.shape[0], sc.ctypes.data, wbits, heavy, nslices)
    want = oracle.msm(curve, pts, sc, "pippenger", ncpus=4)
    return np.array_equal(oracle.jac_to_affine(curve, out), oracle.jac_to_affine(curve, want))


@pytest.mark.parametrize("n,wbits,heavy,nslices", [(1, 0, 0, 1), (2, 0, 0, 1), (3, 4, 0, 1), (33, 0, 0, 1), (300, 0, 0, 1),
                                                   (300, 5, 0, 1), (300, 9, 3, 1), (800, 7, 4, 1), (500, 4, 2, 1),
                                                   (64, 13, 0, 1), (501, 5, 3, 4), (300, 6, 0, 3), (1000, 3, 0, 2)])
def test_msm_batched_affine_prereduction(oracle, msm_emu, n, wbits, heavy, nslices):
    """msm_pair.cuh: every bucket list halved by batched affine pair sums (Montgomery's trick over
    16 pairs per thread and 32 thread totals per inversion), then the XYZZ accumulate in direct
    mode -- chord, tangent (equal points), cancellation (P, -P), infinity inputs, odd tails, heavy
    buckets left to the cooperative path, several slices, ragged last launch."""
    rnd = random.Random(n * 7 + wbits + nslices)
    pts = oracle.gen_points("bls12_381", 32)[np.arange(n) % 32].copy()
    if n > 3:
        pts[3] = 0
    assert _run_pair(oracle, msm_emu, pts, _scalars([rnd.randrange(R_BLS) for _ in range(n)]), wbits, heavy, nslices)


@pytest.mark.parametrize("val", [0, 1, R_BLS - 1, (1 << 254) + 12345, 0x8000000080000000800000008000])
def test_msm_batched_affine_adversarial(oracle, msm_emu, val):
    """all scalars equal: every pair of a bucket is (P_a, P_b) with many P_a == P_b (tangent) and a
    planted (P, -P) (cancellation) -- the cases where the denominator x2 - x1 vanishes"""
    n = 150
    pts = oracle.gen_points("bls12_381", 4)[np.arange(n) % 4].copy()
    p = oracle.ff_consts("bls12_381_fp")["p"]
    y1 = sum(int(v) << (64 * i) for i, v in enumerate(pts[1][6:]))
    pts[5][6:] = [((p - y1) >> (64 * i)) & (2**64 - 1) for i in range(6)]     # points 4,5 -> (P0, -P1); 1 vs 5 cancels when paired
    pts[8:12] = pts[8]                                                         # runs of equal points: tangent pairs
    pts[20] = 0
    pts[21] = 0                                                                # an (inf, inf) pair
    assert _run_pair(oracle, msm_emu, pts, _scalars([val] * n), 6, 1000)
    assert _run_pair(oracle, msm_emu, pts, _scalars([val] * n), 5, 8)          # the same buckets, now heavy
    same = np.tile(pts[7], (40, 1))
    assert _run_pair(oracle, msm_emu, same, _scalars([val] * 40), 6, 1000)     # one point 40 times: only tangents
    pm = same.copy()
    pm[1::2, 6:] = [((p - sum(int(v) << (64 * i) for i, v in enumerate(same[0][6:]))) >> (64 * i)) & (2**64 - 1) for i in range(6)]
    assert _run_pair(oracle, msm_emu, pm, _scalars([val] * 40), 6, 1000)       # (P, -P) pairs only


def test_msm_batched_affine_pallas(oracle, msm_emu):
    rnd = random.Random(4)
    r = oracle.ff_consts("vesta_fp")["p"]
    n = 400
    pts = oracle.gen_points("pallas", 16)[np.arange(n) % 16].copy()
    sc = _scalars([rnd.randrange(r) for _ in range(n)])
    sc[: n // 2] = sc[0]
    assert _run_pair(oracle, msm_emu, pts, sc, 6, 50, 2, curve="pallas")


def test_portable_field_branch(oracle, msm_emu):
    p = oracle.ff_consts("bls12_381_fp")["p"]
    R = 1 << 384
    rnd = random.Random(9)
    for _ in range(300):
        a, b = rnd.randrange(p), rnd.randrange(p)
        A, B, r = oracle.int_to_limbs(a, 6), oracle.int_to_limbs(b, 6), np.zeros(6, dtype=np.uint64)
        for op, exp in ((0, a * b * pow(R, -1, p) % p), (1, (a + b) % p), (2, (a - b) % p)):
            msm_emu.emu_fp_op(op, r.ctypes.data, A.ctypes.data, B.ctypes.data)
            assert oracle.limbs_to_int(r) == exp


def test_ntt_planner_random_splits(oracle, ntt_emu, monkeypatch):
    """Property test of the digit planner + pass kernel (CPU single-stepper): random sizes, digit
    splits, orders, directions and tile sizes must all give the oracle's transform."""
    rnd = random.Random(2024)
    for _ in range(60):
        lg = rnd.randint(2, 11)
        digits, left = [], lg
        while left:
            d = rnd.randint(1, min(left, 6))
            digits.append(d)
            left -= d
        monkeypatch.setenv("SPPARK_B200_NTT_SPLIT", ",".join(map(str, digits)))
        order, inv, lg_tile = rnd.randrange(5), rnd.randrange(2), rnd.choice([5, 7, 9, 14])
        x = np.array([rnd.randrange(2**64 - 2**32 + 1) for _ in range(1 << lg)], dtype=np.uint64)
        y = x.copy()
        ntt_emu.emu_ntt_gl64(y.ctypes.data, lg, order, inv, lg_tile)
        assert np.array_equal(y, oracle.ntt_gl64(x, order, bool(inv))), (lg, digits, order, inv, lg_tile)


def test_ntt_slab_random_shapes(oracle, ntt_emu, monkeypatch):
    """Same for the slab-sharded transform: random digit splits (one to four digits after the
    first), rank counts and tile sizes, staging route and fused-exchange route."""
    from sppark_b200 import parallel
    rnd = random.Random(77)
    done = 0
    while done < 25:
        lg = rnd.randint(4, 11)
        lg_g = rnd.randint(0, 3)
        digits, left = [], lg
        while left:
            d = rnd.randint(1, min(left, 5))
            digits.append(d)
            left -= d
        if len(digits) < 2 or digits[0] < lg_g or lg - digits[0] < lg_g:
            continue
        done += 1
        monkeypatch.setenv("SPPARK_B200_NTT_SPLIT", ",".join(map(str, digits)))
        s1, G, lg_tile = digits[0], 1 << lg_g, rnd.choice([5, 8, 14])
        x = np.array([rnd.randrange(2**64 - 2**32 + 1) for _ in range(1 << lg)], dtype=np.uint64)
        inv = rnd.randrange(2)
        recv = [np.zeros((1 << lg) // G, dtype=np.uint64) for _ in range(G)]
        if rnd.randrange(2):                                            # fused exchange
            ptrs = (C.c_void_p * G)(*[r.ctypes.data for r in recv])
            for r in range(G):
                loc = parallel.scatter_columns(x, lg, lg_g, r, s1=s1).reshape(-1).copy()
                assert ntt_emu.emu_ntt_slab_p2p_gl64(loc.ctypes.data, ptrs, lg, lg_g, r, inv, lg_tile) == 0
        else:                                                           # staging + all-to-all
            stag = []
            for r in range(G):
                loc = parallel.scatter_columns(x, lg, lg_g, r, s1=s1).reshape(-1).copy()
                st = np.zeros_like(loc)
                assert ntt_emu.emu_ntt_slab_gl64(1, loc.ctypes.data, st.ctypes.data, lg, lg_g, r, inv, lg_tile) == 0
                stag.append(st.reshape(G, -1))
            recv = [np.concatenate([stag[q][r] for q in range(G)]).copy() for r in range(G)]
        for r in range(G):
            scratch = np.zeros_like(recv[r])
            assert ntt_emu.emu_ntt_slab_gl64(2, recv[r].ctypes.data, scratch.ctypes.data, lg, lg_g, r, inv, lg_tile) == 0
        assert np.array_equal(parallel.gather_columns(recv, lg, lg_g, s1=s1), oracle.ntt_gl64(x, 0, bool(inv))), (lg, lg_g, digits, inv)


def test_msm_random_configurations(oracle, msm_emu):
    """Random window widths, heavy thresholds, slice counts, duplicate-heavy inputs; plain and
    batched-affine accumulation must both give the oracle's point."""
    rnd = random.Random(31337)
    base = oracle.gen_points("bls12_381", 24)
    for _ in range(24):
        n = rnd.randint(1, 400)
        wbits = rnd.choice([3, 4, 5, 7, 9, 12])
        heavy = rnd.choice([0, 2, 5, 50, 1000])
        nslices = rnd.choice([1, 1, 2, 5])
        pts = base[[rnd.randrange(rnd.choice([2, 24])) for _ in range(n)]].copy()
        if rnd.random() < 0.5:
            pts[rnd.randrange(n)] = 0
        vals = [rnd.randrange(R_BLS) for _ in range(n)]
        if rnd.random() < 0.4:
            vals = [vals[rnd.randrange(3)] for _ in range(n)]          # three distinct scalars: long buckets
        sc = _scalars(vals)
        want = oracle.jac_to_affine("bls12_381", oracle.msm("bls12_381", pts, sc, "pippenger", ncpus=4))
        out = np.zeros(18, dtype=np.uint64)
        msm_emu.emu_msm_bls12_381_sliced(out.ctypes.data, pts.ctypes.data, n, sc.ctypes.data, wbits, heavy, nslices)
        assert np.array_equal(oracle.jac_to_affine("bls12_381", out), want), ("plain", n, wbits, heavy, nslices)
        msm_emu.emu_msm_bls12_381_pair(out.ctypes.data, pts.ctypes.data, n, sc.ctypes.data, wbits, heavy, nslices)
        assert np.array_equal(oracle.jac_to_affine("bls12_381", out), want), ("pair", n, wbits, heavy, nslices)
