"""Host-side window chooser of the MSM (csrc/msm/msm_core.cuh, make_config / choose_wbits): plain
C++, compiled and run here.  Pins the configurations the measurements in DESIGN.md section 7 were
taken with, and the rule added for thin top windows."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "sppark_b200/csrc/msm/msm_core.cuh"
int main()
{
    for (int lg = 10; lg <= 28; lg++) {
        msm::Config c = msm::make_config((size_t)1 << lg);
        printf("%d %u %u %u\n", lg, c.wbits, c.nwins, c.heavy);
    }
    return 0;
}
'''


@pytest.fixture(scope="module")
def table(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg")
    src, exe = d / "cfg.cpp", d / "cfg"
    src.write_text(SRC)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", ROOT, "-I", "/usr/local/cuda/include", "-o", str(exe), str(src)])
    env = {k: v for k, v in os.environ.items() if not k.startswith("SPPARK_B200_MSM_")}
    rows = subprocess.check_output([str(exe)], text=True, env=env).split("\n")
    return {int(r.split()[0]): tuple(int(v) for v in r.split()[1:]) for r in rows if r}


def test_measured_configurations_are_the_ones_chosen(table):
    assert table[26][:2] == (20, 13)          # the headline: 13 windows of 20 bits
    assert table[24][:2] == (20, 13)          # per-GPU share at 4 GPUs, Pallas 2^24
    assert table[23][0] == 16                 # per-GPU share at 8 GPUs (18 before the thin-window term)
    assert table[16][0] == 12 and table[20][0] == 16


def test_windows_cover_the_scalar_and_large_jobs_avoid_thin_top_windows(table):
    for lg, (c, nwins, heavy) in table.items():
        assert nwins * c >= 256 and (nwins - 1) * c < 256
        assert 256 <= heavy <= 16384
        if lg >= 22:
            e = 256 - (nwins - 1) * c
            assert e > 10 or ((1 << lg) >> e) <= 2048, (lg, c, e)
