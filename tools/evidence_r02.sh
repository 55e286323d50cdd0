#!/bin/bash
# Round-2 evidence batch (run under gpurun, one GPU): ncu launch list of the bench command and
# `ncu --set full` summaries of the kernels round 1 had no capture for.  Writes to gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# 1. launch list of the bench command (short run)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches_r02.csv \
    python bench.py --steps 1 --warmup 1 --skip-cpu --skip-ref-gpu > $O/bench_under_ncu_r02.log 2>&1
python tools/launch_list.py $O/bench_launches_r02.csv $O/bench_launches_r02.md "ncu launch list of python bench.py --steps 1 --warmup 1 (round 2)" > /dev/null
# 2. BabyBear 2^24 NN (block-tile passes)
$NCU -k regex:pass_kernel -s 2 -c 2 -o /tmp/bb python tools/run_ntt_any.py bb31 24 0 2 > $O/ncu_bb31.log 2>&1
python tools/ncu_summary.py /tmp/bb.ncu-rep $O/ntt_bb31_block_r02.md "BabyBear NTT 2^24 NN, block-tile passes" > /dev/null
# 3. BabyBear 2^18 (warp-autonomous passes)
$NCU -k regex:pass_kernel_warp -s 3 -c 3 -o /tmp/bbw python tools/run_ntt_any.py bb31 18 0 2 > $O/ncu_bb31w.log 2>&1
python tools/ncu_summary.py /tmp/bbw.ncu-rep $O/ntt_bb31_warp_r02.md "BabyBear NTT 2^18 NN, warp-autonomous passes" > /dev/null
# 4. 256-bit (BLS12-381 fr) 2^20 NN
$NCU -k regex:pass_kernel -s 2 -c 2 -o /tmp/w256 python tools/run_ntt_any.py bls12_381_fr 20 0 2 > $O/ncu_256.log 2>&1
python tools/ncu_summary.py /tmp/w256.ncu-rep $O/ntt_256bit_r02.md "BLS12-381 scalar-field NTT 2^20 NN (256-bit Montgomery words)" > /dev/null
# 5. LDE and coset kernels (Goldilocks 2^22 -> 2^23)
$NCU -k regex:"lde_spread|coset_kernel|bitrev_copy" -c 3 -o /tmp/lde python tools/run_ntt_any.py gl64 23 0 1 lde > $O/ncu_lde.log 2>&1
python tools/ncu_summary.py /tmp/lde.ncu-rep $O/ntt_lde_r02.md "LDE spread / coset scaling kernels (Goldilocks 2^22 -> 2^23)" > /dev/null
# 6. MSM tail kernels at 2^24
$NCU -k regex:"scan_kernel|reduce1_kernel|combine_kernel|finish_kernel|heavy" -s 0 -c 8 -o /tmp/tail python tools/run_msm_once.py 24 1 > $O/ncu_msm_tail.log 2>&1
python tools/ncu_summary.py /tmp/tail.ncu-rep $O/msm_tail_r02.md "MSM 2^24: scan / reduce / combine / finish kernels" > /dev/null
ls -la $O/*_r02.md
