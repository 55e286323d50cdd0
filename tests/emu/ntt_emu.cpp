// CPU single-stepper for the NTT pass kernel -- TEST INFRASTRUCTURE ONLY.
//
// Runs the exact HD phase functions of sppark_b200/csrc/ntt/ntt_core.cuh with the plan of
// ntt_plan.hpp, one "thread" at a time, one phase at a time (a phase boundary is where the
// CUDA kernel has __syncthreads()).  Lets `pytest -m "not gpu"` validate the planner's index
// algebra and the butterfly schedule against the oracle without a GPU.  It is not linked
// into libsppark_b200.so and is not a fallback of any kind.
#include <cstdio>
#include <vector>
#include "../../sppark_b200/csrc/ff/gl64.cuh"
#include "../../sppark_b200/csrc/ff/bb31.cuh"
#include "../../sppark_b200/csrc/ff/mont_ntt.cuh"
#include "../../sppark_b200/csrc/ntt/ntt_plan.hpp"

using namespace ntt;

template<class F> struct HostTables {
    std::vector<typename F::T> dense, tlo, thi;
    Tables<F> view;
    HostTables(uint32_t lg_n, bool inverse)
    {
        typedef typename F::T T;
        T w_max = F::root_of_unity_max();
        if (inverse) w_max = F::inv(w_max);
        auto root = [&](uint32_t lg) {            // primitive 2^lg-th root
            T w = w_max;
            for (uint32_t i = F::MAX_LG; i > lg; i--) w = F::mul(w, w);
            return w;
        };
        dense.assign(1u << LG_DENSE, F::one());
        for (uint32_t lg_h = 0; lg_h < LG_DENSE; lg_h++) {
            uint32_t h = 1u << lg_h;
            T w = root(lg_h + 1), acc = F::one();
            for (uint32_t i = 0; i < h; i++, acc = F::mul(acc, w)) dense[h + i] = acc;
        }
        T wn = root(lg_n);
        tlo.resize(1u << LG_TLO);
        T acc = F::one();
        for (uint32_t i = 0; i < (1u << LG_TLO); i++, acc = F::mul(acc, wn)) tlo[i] = acc;
        uint32_t nhi = lg_n > LG_TLO ? 1u << (lg_n - LG_TLO) : 1;
        thi.resize(nhi);
        T step = acc;                              // w_N^(2^LG_TLO)
        acc = F::one();
        for (uint32_t i = 0; i < nhi; i++, acc = F::mul(acc, step)) thi[i] = acc;
        T half = F::inv(F::add(F::one(), F::one()));
        T ninv = F::one();
        for (uint32_t i = 0; i < lg_n; i++) ninv = F::mul(ninv, half);
        view = Tables<F>{dense.data(), tlo.data(), thi.data(), ninv};
    }
};

template<class F>
static int emu_run(typename F::T* data, uint32_t lg_n, int order, int inverse, uint32_t lg_tile)
{
    typedef typename F::T T;
    if (lg_n == 0) return 0;
    HostTables<F> tb(lg_n, inverse != 0);
    Plan plan = make_plan(lg_n, order, inverse != 0, lg_tile, 6, F::NTT_MAX_LG_R);
    std::vector<T> scratch(plan.needs_scratch ? (size_t)1 << lg_n : 0);
    T* buf[2] = {data, scratch.data()};
    for (const Pass& d : plan.passes) {
        uint32_t nthreads = tile_threads<F>(d);
        uint32_t ntiles = 1u << (lg_n - d.lg_r - d.lg_w);
        std::vector<T> smem(smem_elems(d));
        // out-of-place passes read src while other tiles write dst: they never alias.
        // in-place passes touch only their own tile.  Either way tile order is free.
        for (uint32_t t = 0; t < ntiles; t++) {
            const KDyn k{d};
            for (uint32_t tid = 0; tid < nthreads; tid++) phase_twiddles<F>(k, tb.view, smem.data(), tid, nthreads);
            for (uint32_t tid = 0; tid < nthreads; tid++) phase_load<F>(k, d, tb.view, buf[d.src], smem.data(), t, tid, nthreads);
            for (uint32_t s = 0; s < step_count<F>(d.lg_r); s++)
                for (uint32_t tid = 0; tid < nthreads; tid++)
                    phase_step_dyn<F>(k, smem.data(), s * F::LG_EPT, step_log_e<F>(d.lg_r, s), tid);
            for (uint32_t tid = 0; tid < nthreads; tid++) phase_store<F>(k, d, tb.view, buf[d.dst], smem.data(), t, tid, nthreads);
        }
    }
    return (int)plan.passes.size();
}

// one local stage of the slab-sharded transform on a "rank"'s buffers (see make_slab_plan):
// which = 1: in -> out (staging); which = 2: `in` transformed with `out` as scratch, result in `in`
template<class F>
static int emu_slab_pass(int which, const typename F::T* in, typename F::T* out, uint32_t lg_n, uint32_t lg_g,
                         uint32_t rank, int inverse, uint32_t lg_tile, void* const* peers = nullptr)
{
    typedef typename F::T T;
    HostTables<F> tb(lg_n, inverse != 0);
    SlabPlan sp;
    if (!make_slab_plan(sp, lg_n, lg_g, rank, inverse != 0, lg_tile, F::NTT_MAX_LG_R)) return -1;
    if (which == 2 && sp.needs_scratch && in == out) return -2;
    auto run = [&](const Pass& d, const T* src, T* dst) {
        uint32_t nthreads = tile_threads<F>(d), ntiles = 1u << (lg_n - lg_g - d.lg_r - d.lg_w);
        std::vector<T> smem(smem_elems(d));
        const KDyn k{d};
        for (uint32_t t = 0; t < ntiles; t++) {
            for (uint32_t tid = 0; tid < nthreads; tid++) phase_twiddles<F>(k, tb.view, smem.data(), tid, nthreads);
            for (uint32_t tid = 0; tid < nthreads; tid++) phase_load<F>(k, d, tb.view, src, smem.data(), t, tid, nthreads);
            for (uint32_t s = 0; s < step_count<F>(d.lg_r); s++)
                for (uint32_t tid = 0; tid < nthreads; tid++)
                    phase_step_dyn<F>(k, smem.data(), s * F::LG_EPT, step_log_e<F>(d.lg_r, s), tid);
            for (uint32_t tid = 0; tid < nthreads; tid++) phase_store<F>(k, d, tb.view, dst, smem.data(), t, tid, nthreads);
        }
    };
    if (which == 1 && peers) {                       // fused exchange: rows go straight to their receiver
        Pass d = sp.pass1;
        d.peer_on = 1;
        for (uint32_t q = 0; q < (1u << lg_g); q++) d.peer[q] = (uint64_t)(uintptr_t)peers[q];
        run(d, in, nullptr);
    } else if (which == 1) {
        run(sp.pass1, in, out);
    } else {
        T* buf[2] = {const_cast<T*>(in), out};
        for (const Pass& d : sp.after) run(d, buf[d.src], buf[d.dst]);
    }
    return 0;
}
extern "C" int emu_ntt_slab_p2p_gl64(const uint64_t* in, void* const* peers, uint32_t lg_n, uint32_t lg_g,
                                     uint32_t rank, int inverse, uint32_t lg_tile)
{   return emu_slab_pass<gl64>(1, in, nullptr, lg_n, lg_g, rank, inverse, lg_tile, peers);   }
extern "C" int emu_ntt_slab_p2p_bb31(const uint32_t* in, void* const* peers, uint32_t lg_n, uint32_t lg_g,
                                     uint32_t rank, int inverse, uint32_t lg_tile)
{   return emu_slab_pass<bb31>(1, in, nullptr, lg_n, lg_g, rank, inverse, lg_tile, peers);   }
extern "C" int emu_slab_first_digit(uint32_t lg_n, uint32_t max_lg_r) { return (int)slab_first_digit(lg_n, max_lg_r); }
extern "C" int emu_ntt_slab_gl64(int which, const uint64_t* in, uint64_t* out, uint32_t lg_n, uint32_t lg_g,
                                 uint32_t rank, int inverse, uint32_t lg_tile)
{   return emu_slab_pass<gl64>(which, in, out, lg_n, lg_g, rank, inverse, lg_tile);   }
extern "C" int emu_ntt_slab_bb31(int which, const uint32_t* in, uint32_t* out, uint32_t lg_n, uint32_t lg_g,
                                 uint32_t rank, int inverse, uint32_t lg_tile)
{   return emu_slab_pass<bb31>(which, in, out, lg_n, lg_g, rank, inverse, lg_tile);   }

extern "C" int emu_ntt_gl64(uint64_t* data, uint32_t lg_n, int order, int inverse, uint32_t lg_tile)
{   return emu_run<gl64>(data, lg_n, order, inverse, lg_tile);   }
extern "C" int emu_ntt_bb31(uint32_t* data, uint32_t lg_n, int order, int inverse, uint32_t lg_tile)
{   return emu_run<bb31>(data, lg_n, order, inverse, lg_tile);   }
extern "C" int emu_ntt_256(int field, uint32_t* data, uint32_t lg_n, int order, int inverse, uint32_t lg_tile)
{
    switch (field) {
    case 2: return emu_run<ff::bls12_381_fr_ntt>((ff::bls12_381_fr_ntt::T*)data, lg_n, order, inverse, lg_tile);
    case 3: return emu_run<ff::pallas_fr_ntt>((ff::pallas_fr_ntt::T*)data, lg_n, order, inverse, lg_tile);
    default: return emu_run<ff::vesta_fr_ntt>((ff::vesta_fr_ntt::T*)data, lg_n, order, inverse, lg_tile);
    }
}
