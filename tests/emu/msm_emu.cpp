// CPU single-stepper for the MSM pipeline -- TEST INFRASTRUCTURE ONLY.
// Executes the HD kernel bodies of sppark_b200/csrc/msm/msm_core.cuh (portable arithmetic
// branch) in the same order as msm_t::invoke_dev launches them.  Not linked into the product.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include "../../sppark_b200/csrc/ff/fields.cuh"
#include "../../sppark_b200/csrc/msm/msm_core.cuh"
#include "../../sppark_b200/csrc/msm/msm_pair.cuh"

using namespace msm;

template<class F>
static void emu_msm(uint32_t* out, const uint32_t* points_all, size_t npoints_all, const uint32_t* scalars_all,
                    uint32_t wbits, uint32_t heavy, uint32_t nslices = 1, bool pair = false)
{
    size_t npoints = npoints_all;
    constexpr uint32_t BW = 4 * F::N, JW = 3 * F::N;
    if (npoints == 0) { memset(out, 0, JW * 4); return; }
    Config cfg = make_config(npoints);
    if (wbits) { cfg.wbits = wbits; cfg.nwins = (256 + wbits - 1) / wbits; cfg.lg_nb = wbits - 1; }
    if (heavy) { cfg.heavy = heavy; cfg.heavy_chunk = 4 * heavy; }
    const size_t nslots = (size_t)cfg.nwins << cfg.lg_nb;
    std::vector<uint32_t> counts(nslots, 0), offsets(nslots), cursor(nslots), sorted((size_t)cfg.nwins * npoints_all);
    std::vector<uint32_t> buckets(nslots * BW, 0xdeadbeef), heavy_list;
    if (nslices == 0) nslices = 1;
    const size_t slice_n = (npoints_all + nslices - 1) / nslices;
    for (size_t first = 0, sl = 0; first < npoints_all; first += slice_n, sl++) {
    // ---- one slice: msm_t::slice() -------------------------------------------------------------
    npoints = std::min(slice_n, npoints_all - first);
    const uint32_t* points = points_all + first * 2 * F::N;
    const uint32_t* scalars = scalars_all + first * 8;
    cfg.npoints = (uint32_t)npoints;
    cfg.merge = sl ? 1 : 0;
    std::fill(counts.begin(), counts.end(), 0);
    heavy_list.clear();
    for (uint32_t i = 0; i < npoints; i++) count_body(cfg, scalars, counts.data(), i);
    for (uint32_t w = 0; w < cfg.nwins; w++) {                 // scan_kernel
        uint32_t run = 0;
        for (uint32_t b = 0; b < (1u << cfg.lg_nb); b++) {
            size_t t = ((size_t)w << cfg.lg_nb) + b;
            offsets[t] = cursor[t] = run;
            if (counts[t] > cfg.heavy) heavy_list.push_back((uint32_t)t);
            run += counts[t];
        }
    }
    for (uint32_t i = 0; i < npoints; i++) scatter_body(cfg, scalars, cursor.data(), sorted.data(), i, 0, cfg.nwins);
    uint32_t task_counter = 0;
    if (pair) {
        // msm_pair.cuh: lists -> pair sums (same launch sequence as msm_t::slice with pairing on)
        std::vector<uint32_t> counts1(nslots + 1, 1), off1(nslots), winbase(cfg.nwins + 1);   // sentinel slot stops pair_advance
        for (uint32_t t = 0; t < nslots; t++) pair_counts_body(cfg, counts.data(), counts1.data(), t);
        uint32_t total1 = 0;
        for (uint32_t w = 0; w < cfg.nwins; w++) {
            winbase[w] = total1;
            uint32_t run = 0;
            for (uint32_t b = 0; b < (1u << cfg.lg_nb); b++) {
                off1[((size_t)w << cfg.lg_nb) + b] = run;
                run += counts1[((size_t)w << cfg.lg_nb) + b];
            }
            total1 += run;
        }
        winbase[cfg.nwins] = total1;
        std::vector<uint32_t> sums((size_t)std::max(total1, 1u) * 2 * F::N);
        const uint32_t chunk = 5 * PAIR_K;                          // several launches, a ragged last one
        for (uint32_t o0 = 0; o0 < total1; o0 += chunk) {
            const uint32_t nth = chunk / PAIR_K;
            std::vector<uint32_t> pre((size_t)PAIR_K * nth * F::N), totals((size_t)nth * F::N);
            for (uint32_t tid = 0; tid < nth; tid++)
                pair_forward_body<F>(cfg, points, sorted.data(), offsets.data(), counts.data(), counts1.data(), off1.data(),
                                     winbase.data(), o0, nth, pre.data(), totals.data(), tid);
            for (uint32_t tid = 0; tid * PAIR_M < nth; tid++) pair_invert_body<F>(totals.data(), nth, tid);
            for (uint32_t tid = 0; tid < nth; tid++)
                pair_backward_body<F>(cfg, points, sorted.data(), offsets.data(), counts.data(), counts1.data(), off1.data(),
                                      winbase.data(), o0, nth, pre.data(), totals.data(), sums.data(), tid);
        }
        accumulate_body<F, true>(cfg, sums.data(), sorted.data(), offsets.data(), counts.data(), buckets.data(), &task_counter,
                           counts1.data(), off1.data(), winbase.data());
    } else
    accumulate_body<F>(cfg, points, sorted.data(), offsets.data(), counts.data(), buckets.data(), &task_counter);
    const uint32_t HT = 8;                                        // heavy_kernel with 8 "threads"
    for (uint32_t t : heavy_list) {
        const uint32_t* run = sorted.data() + (size_t)(t >> cfg.lg_nb) * cfg.npoints + offsets[t];
        std::vector<uint32_t> tree(HT * BW);
        ec::xyzz_t<F> acc[HT];
        for (uint32_t th = 0; th < HT; th++) {
            acc[th].set_inf();
            for (uint32_t k = th; k < counts[t]; k += HT) acc[th].madd(load_point<F>(points, run[k]));
            store_bucket<F>(tree.data(), th, acc[th]);
        }
        for (uint32_t d = HT / 2; d > 0; d >>= 1)
            for (uint32_t th = 0; th < d; th++) {
                acc[th].add(load_bucket<F>(tree.data(), th + d));
                store_bucket<F>(tree.data(), th, acc[th]);
            }
        if (cfg.merge) acc[0].add(load_bucket<F>(buckets.data(), t));
        store_bucket<F>(buckets.data(), t, acc[0]);
    }
    }   // slices
    const uint32_t lg_l = cfg.lg_nb > 3 ? cfg.lg_nb - 3 : 0;      // small chunks so that every level runs
    uint32_t per_win = 1u << (cfg.lg_nb - lg_l), items = cfg.nwins * per_win;
    std::vector<uint32_t> R[2], S[2];
    for (auto& v : R) v.assign((size_t)items * BW, 0);
    for (auto& v : S) v.assign((size_t)items * BW, 0);
    for (uint32_t it = 0; it < items; it++) reduce1_body<F>(cfg, buckets.data(), lg_l, R[0].data(), S[0].data(), it);
    uint32_t lg_span = lg_l, cur = 0;
    while (per_win > 1) {
        uint32_t lg_g = 31 - __builtin_clz(per_win);
        if (lg_g > 2) lg_g = 2;
        uint32_t G = 1u << lg_g, n = cfg.nwins * (per_win >> lg_g);
        for (uint32_t it = 0; it < n; it++)
            combine_body<F>(R[cur].data(), S[cur].data(), G, lg_span, R[cur ^ 1].data(), S[cur ^ 1].data(), it);
        per_win >>= lg_g; lg_span += lg_g; cur ^= 1;
    }
    finish_body<F>(cfg, R[cur].data(), out);
}

extern "C" void emu_msm_bls12_381(uint32_t* out, const uint32_t* points, size_t n, const uint32_t* scalars,
                                  uint32_t wbits, uint32_t heavy)
{   emu_msm<ff::bls12_381_fp_t>(out, points, n, scalars, wbits, heavy);   }
extern "C" void emu_msm_bls12_381_sliced(uint32_t* out, const uint32_t* points, size_t n, const uint32_t* scalars,
                                         uint32_t wbits, uint32_t heavy, uint32_t nslices)
{   emu_msm<ff::bls12_381_fp_t>(out, points, n, scalars, wbits, heavy, nslices);   }
extern "C" void emu_msm_bls12_381_pair(uint32_t* out, const uint32_t* points, size_t n, const uint32_t* scalars,
                                       uint32_t wbits, uint32_t heavy, uint32_t nslices)
{   emu_msm<ff::bls12_381_fp_t>(out, points, n, scalars, wbits, heavy, nslices, true);   }
extern "C" void emu_msm_pallas_pair(uint32_t* out, const uint32_t* points, size_t n, const uint32_t* scalars,
                                    uint32_t wbits, uint32_t heavy, uint32_t nslices)
{   emu_msm<ff::pallas_fp_t>(out, points, n, scalars, wbits, heavy, nslices, true);   }
extern "C" void emu_msm_pallas(uint32_t* out, const uint32_t* points, size_t n, const uint32_t* scalars,
                               uint32_t wbits, uint32_t heavy)
{   emu_msm<ff::pallas_fp_t>(out, points, n, scalars, wbits, heavy);   }

// field KATs through the portable branch: op 0 mul, 1 add, 2 sub
extern "C" void emu_fp_op(int op, uint32_t* r, const uint32_t* a, const uint32_t* b)
{
    ff::bls12_381_fp_t x, y, z;
    memcpy(x.l, a, 48); memcpy(y.l, b, 48);
    z = op == 0 ? x * y : op == 1 ? x + y : x - y;
    memcpy(r, z.l, 48);
}
