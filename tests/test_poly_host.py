"""Host-side pieces of the polynomial helpers (csrc/poly/poly.cuh), compiled with g++ and run here:
the Goldilocks data/constant arithmetic of arith<gl64> and the power tables scan_tab_fill hands to
the division kernels, against Python integers."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**64 - 2**32 + 1
SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <cuda_runtime.h>
#include "sppark_b200/csrc/poly/poly.cuh"
using namespace poly;
int main(int argc, char** argv)
{
    typedef arith<gl64> A;
    uint64_t z = strtoull(argv[1], nullptr, 10), a = strtoull(argv[2], nullptr, 10), b = strtoull(argv[3], nullptr, 10);
    // data x data, data x constant, constant x constant -> data, inverse
    printf("%llu %llu %llu %llu\n", (unsigned long long)A::dmul(a, b), (unsigned long long)A::cmul(a, A::konst(b)),
           (unsigned long long)A::cmul(1, A::kmul(A::konst(a), A::konst(b))), (unsigned long long)A::inv(a));
    scan_tab<uint64_t, 8, 256> t;
    scan_tab_fill<gl64, 8, 256>(t, A::konst(z));
    const uint64_t* w = (const uint64_t*)&t;
    for (size_t i = 0; i < sizeof(t) / 8; i++) printf("%llu\n", (unsigned long long)A::cmul(1, w[i]));   // constants -> plain
    return 0;
}
'''


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("polyhost")
    src, out = d / "t.cpp", d / "t"
    src.write_text(SRC)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", ROOT, "-I", "/usr/local/cuda/include", "-o", str(out), str(src)])
    return str(out)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_goldilocks_domains_and_power_tables(exe, seed):
    rnd = random.Random(seed)
    z, a, b = (rnd.randrange(1, P) for _ in range(3))
    if seed == 3:
        a, b = P - 1, 0xffffffff00000000                 # wrap-around corners of the plain product
    lines = subprocess.check_output([exe, str(z), str(a), str(b)], text=True).split("\n")
    dmul, cmul, kmul, inv = (int(v) for v in lines[0].split())
    assert dmul == cmul == kmul == a * b % P
    assert inv * a % P == 1
    w = [int(v) for v in lines[1:] if v]
    E, BS, NW = 8, 256, 8
    want = [pow(z, E * k, P) for k in range(33)]                          # wl
    want += [pow(z, 32 * E * k, P) for k in range(NW + 1)]                 # ww
    want += [pow(z, k, P) for k in range(E + 1)]                           # zp
    zt = pow(z, BS * E, P)
    want += [zt]                                                           # zt
    want += [pow(z, 1 << k, P) for k in range(10)] + [pow(z, BS, P)]       # zw, y
    want += [pow(zt, 1 << k, P) for k in range(10)] + [pow(zt, BS, P)]     # tw, yt
    assert w == want
