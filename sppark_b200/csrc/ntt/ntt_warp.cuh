// Warp-autonomous NTT pass for single-word fields (gl64, bb31): sub-NTTs of 2^R points, 4 <= R <= 8.
//
// Same job and same descriptor as the block-tile pass of ntt_core.cuh (`Pass`: a tile of 2^lg_w
// columns x 2^R rows, strided or contiguous, natural or bit-reversed rows, inter-pass twiddle at
// load or store, scaling, slab / peer routing), different execution shape: 2^(R-4) lanes own one
// column; every lane keeps 16 of its elements in registers.
//
//   step 0   the lane with row-residue b loads rows a*L + b (a = 0..15, L = 2^(R-4)) straight from
//            HBM into registers and runs a 16-point DFT over a: all its twiddles are powers of
//            w_16, eight per-field constants that sit in the constant bank -- no table lookups;
//   twist    Y_b[k0] *= w_(2^R)^(b*k0): the only table of the pass, 16*L words;
//   exchange a 16 x L transposition through a private slice of shared memory, __syncwarp() only;
//   step 1   16/L DFTs of L points (constants again), giving X[k0 + 16*k1];
//   store    inter-pass twiddle w_N^(k*col) generated as ONE running product per lane (k = l + L*j
//            walks an arithmetic progression: two table look-ups per lane instead of two per
//            element), then registers -> HBM.
//
// There is no block-level barrier and no block-level tile: a warp never waits for another warp, the
// HBM latency of one warp's loads is covered by the arithmetic of the other warps of the SM (up to
// 32 resident), and transforms of any size run the same code.  A 2^24 transform is three passes
// (8+8+8) that move 6 x 128 MiB; the reference's three <=10-stage steps plus its bit-reversal pass
// move 8 x 128 MiB (ntt/ntt.cuh:100-127,174-178).
//
// CUDA only.  The CPU single-stepper (tests/emu) keeps checking the planner through the block-tile
// phases, which implement the same descriptors; the GPU parity tests run this kernel.
#pragma once
#include "ntt_core.cuh"

namespace ntt {

constexpr uint32_t WARP_MIN_LG_R = 4, WARP_MAX_LG_R = 8;
// offset of the twist table of a 2^R pass inside Tables::mid (R = 5..8: 32, 64, 128, 256 words)
HD constexpr uint32_t mid_offset(uint32_t R) { return R <= 5 ? 0 : 16u * ((1u << (R - 4)) - 2); }
constexpr uint32_t MID_WORDS = 480;

#if defined(__CUDACC__)

// bit reversal of a value of at most 4 bits, loop-free so that it always folds to a constant
__device__ __forceinline__ constexpr uint32_t cbrev(uint32_t v, uint32_t bits)
{
    return (((v & 1u) << 3) | ((v & 2u) << 1) | ((v & 4u) >> 1) | ((v & 8u) >> 3)) >> (4 - bits);
}

// One stage of the in-place radix-2 DIT network on 16 registers.  Logical element i lives in
// register brev4(i); stage S pairs logical i and i + 2^(S-1) inside blocks of 2^S with the twiddle
// w_(2^S)^j = w16^(j * 16 / 2^S).  Running stages 1..4 turns 16 natural-order inputs (register r =
// input r) into a 16-point DFT (X[k] in register brev4(k)); running only stages 1..n performs
// 16 / 2^n independent 2^n-point DFTs on the logical blocks -- the same code serves both register
// steps of a pass.  Second operands of add/sub must be canonical: products are, the un-multiplied
// ones (j = 0) are tightened.
template<class F, uint32_t S>
__device__ __forceinline__ void dft_stage(typename F::T (&x)[16], const Tables<F>& tb)
{
    typedef typename F::T T;
    constexpr uint32_t half = 1u << (S - 1);
#pragma unroll
    for (uint32_t k = 0; k < 16; k += 2 * half) {
#pragma unroll
        for (uint32_t j = 0; j < half; j++) {
            const uint32_t i0 = cbrev(k + j, 4), i1 = cbrev(k + j + half, 4);
            const T t = j == 0 ? F::tight(x[i1]) : F::mul(x[i1], tb.w16[j * (8u >> (S - 1))]);
            const T u = x[i0];
            x[i0] = F::add(u, t);
            x[i1] = F::sub(u, t);
        }
    }
}

template<class T, uint32_t CPT> struct VecLoad;
template<> struct VecLoad<uint64_t, 2> {
    static __device__ __forceinline__ void st(uint64_t* p, uint64_t a, uint64_t b)
    {   *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);   }
};
template<> struct VecLoad<uint32_t, 2> {
    static __device__ __forceinline__ void st(uint32_t* p, uint32_t a, uint32_t b)
    {   *reinterpret_cast<uint2*>(p) = make_uint2(a, b);   }
};

// w_N^e from the two-level table, branch-free (thi[0] = 1)
template<class F> __device__ __forceinline__ typename F::T twiddle2(const Tables<F>& tb, uint32_t e)
{   return F::mul(tb.tlo[e & ((1u << LG_TLO) - 1)], tb.thi[e >> LG_TLO]);   }

// asynchronous global -> shared copy of one element (or of two adjacent ones): LDGSTS, the data
// never passes through registers and the issuing warp does not wait for it
template<uint32_t BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src)
{
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(dst), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// shared-memory words per warp and column slot: exchange / output staging, and the input stage
HD constexpr uint32_t warp_xchg_words(uint32_t R)
{
    return (32u >> (R - 4)) * (16u * ((1u << (R - 4)) + 1) + (1u << (R - 4)));
}
constexpr uint32_t WARP_STAGE_WORDS = 16 * 32;
HD constexpr uint32_t warp_smem_words(uint32_t R, uint32_t cpt) { return cpt * (warp_xchg_words(R) + WARP_STAGE_WORDS); }
// per CTA: the twist table of the pass (16 * 2^(R-4) words) in front of the warps' buffers
HD constexpr uint32_t cta_smem_words(uint32_t R, uint32_t cpt, uint32_t warps)
{   return (16u << (R - 4)) + warps * warp_smem_words(R, cpt);   }

#ifndef SPPARK_B200_NTT_WARP_MINB
#define SPPARK_B200_NTT_WARP_MINB 3
#endif
// TW = the pass's tw_mode (TW_NONE / TW_LOAD / TW_STORE), fixed at compile time so that each
// instantiation carries one copy of the twiddle code at most: the loop body must stay inside the
// instruction cache (warps run the code at different times, nothing is shared between them)
template<class F, uint32_t R, uint32_t CPT, uint32_t TW>
__global__ void __launch_bounds__(256, CPT == 1 ? SPPARK_B200_NTT_WARP_MINB : 2)
pass_kernel_warp(const Pass d, const Tables<F> tb, const typename F::T* __restrict__ in,
                 typename F::T* __restrict__ out, uint32_t ncols)
{
    typedef typename F::T T;
    static_assert(R >= WARP_MIN_LG_R && R <= WARP_MAX_LG_R, "sub-NTT size");
    constexpr uint32_t R1 = R - 4, L = 1u << R1, G = 16u >> R1, SPW = 32u >> R1;
    constexpr uint32_t SK = L + 1, SS = 16 * (L + 1) + L;      // conflict-free strides of the exchange
    constexpr uint32_t XW = SPW * SS, SW = WARP_STAGE_WORDS;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const uint32_t s = lane >> R1, l = lane & (L - 1);
    const uint32_t lrev = brev32(l, R1);
    // xb: exchange + output staging, element (q, sub-NTT s, row j, lane l); sb: input stage (a, lane, q)
    T* const midsh = reinterpret_cast<T*>(smem_raw);
    T* const xb = midsh + 16 * L + warp * warp_smem_words(R, CPT) + s * SS + l;
    T* const sb = midsh + 16 * L + warp * warp_smem_words(R, CPT) + CPT * XW + lane * CPT;
    if (R1 > 0) {
        // twist table w_(2^R)^(l * k0) of this pass: once per persistent CTA, one bulk asynchronous
        // copy (TMA) signalled through an mbarrier
        __shared__ uint64_t mid_bar;
        if (threadIdx.x == 0) mbar_init(&mid_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(midsh, tb.mid + mid_offset(R), (uint32_t)(16 * L * sizeof(T)), &mid_bar);
        mbar_wait(&mid_bar, 0);
    }

    const uint32_t nunits = (ncols + SPW * CPT - 1) / (SPW * CPT);
    const uint32_t stride = gridDim.x * warps;
    // loop-invariant pieces of the addresses; every element index fits 32 bits (<= 2^30 elements)
    const uint32_t in_lane_off = (d.in_rev ? (lrev << 4) : l) << d.in_lg_sa;
    const uint32_t out_lane_row = d.out_rev ? (lrev << 4) : l;
    const uint32_t out_lane_off = out_lane_row << d.out_lg_sa;
    const bool in_vec = CPT == 2 && d.in_lg_sc == 0 && d.lg_w >= 1 && d.in_lg_sa >= 1 && ncols >= 2;
    const bool out_plain = !d.peer_on && !d.out_split_bits;
    const bool out_vec = CPT == 2 && d.out_lg_sc == 0 && d.lg_w >= 1 && d.out_lg_sa >= 1 && ncols >= 2 && out_plain;

    // column geometry of a unit: lanes beyond the last column (transforms with fewer columns than a
    // warp holds) read the last column and store nothing
    auto column = [&](uint32_t unit, uint32_t q, uint32_t& ibase, uint32_t& obase) -> bool {
        uint32_t gc = (unit * SPW + s) * CPT + q;
        const bool act = gc < ncols;
        gc = act ? gc : (ncols >= CPT ? ncols - CPT + q : ncols - 1);
        const uint32_t t = gc >> d.lg_w, c = gc & ((1u << d.lg_w) - 1);
        ibase = (uint32_t)tile_base(t, d.in_lg_tlo, d.in_tl, d.in_th) + (c << d.in_lg_sc);
        obase = (uint32_t)tile_base(t, d.out_lg_tlo, d.out_tl, d.out_th) + (c << d.out_lg_sc);
        return act;
    };
    // stage <- natural rows a*L + l (a = 0..15) of this lane's columns, stored at row brev_R(.) if in_rev
    auto prefetch = [&](uint32_t unit) {
        uint32_t ib[CPT], ob;
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) column(unit, q, ib[q], ob);
        const uint32_t in_step = (d.in_rev ? 1u : L) << d.in_lg_sa;
#pragma unroll
        for (uint32_t a = 0; a < 16; a++) {
            const uint32_t off = in_lane_off + (d.in_rev ? cbrev(a, 4) : a) * in_step;
            if (CPT == 2 && in_vec) {
                cp_async<2 * sizeof(T)>(sb + a * (32 * CPT), in + (ib[0] + off));
            } else {
#pragma unroll
                for (uint32_t q = 0; q < CPT; q++) cp_async<sizeof(T)>(sb + a * (32 * CPT) + q, in + (ib[q] + off));
            }
        }
        cp_async_commit();
    };

    uint32_t unit = blockIdx.x * warps + warp;
    if (unit < nunits) prefetch(unit);
#pragma unroll 1
    for (; unit < nunits; unit += stride) {
        uint32_t ibase[CPT], obase[CPT];
        bool active[CPT];
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) active[q] = column(unit, q, ibase[q], obase[q]);

        // inter-pass twiddle of this unit's columns: w^((l + L*j) * colv << lsh) = tw * step^j
        // generated as four interleaved geometric sequences tw[v] * (step^4)^u, j = 4u + v, so that
        // the dependent chain is 4 products long instead of 16
        T tw[CPT][4], step4[CPT];
        if constexpr (TW != TW_NONE) {
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++) {
                const uint32_t colv = tw_column_value(d, ibase[q]);
                T t0 = twiddle2<F>(tb, (l * colv) << d.tw_lsh);
                const T st = twiddle2<F>(tb, (L * colv) << d.tw_lsh);
                if (d.scale) t0 = F::mul(t0, tb.ninv);
                const T st2 = F::mul(st, st);
                step4[q] = F::mul(st2, st2);
                tw[q][0] = t0;
                tw[q][1] = F::mul(t0, st);
                tw[q][2] = F::mul(t0, st2);
                tw[q][3] = F::mul(tw[q][1], st2);
            }
        }

        cp_async_wait_all();
        if constexpr (TW == TW_LOAD) {                     // RN plans: x[a] *= w^((a*L + l) * colv << lsh)
#pragma unroll 1
            for (uint32_t u = 0; u < 4; u++) {
#pragma unroll
                for (uint32_t v = 0; v < 4; v++) {
#pragma unroll
                    for (uint32_t q = 0; q < CPT; q++) {
                        T* p = sb + (4 * u + v) * (32 * CPT) + q;
                        *p = F::mul(F::load(*p), tw[q][v]);
                        if (u != 3) tw[q][v] = F::mul(tw[q][v], step4[q]);
                    }
                }
            }
        }
        T x[CPT][16];
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++)
#pragma unroll
            for (uint32_t a = 0; a < 16; a++) x[q][a] = F::load(sb[a * (32 * CPT) + q]);
        // the next unit's rows travel while this one is transformed
        if (unit + stride < nunits) prefetch(unit + stride);

        // ---- two register steps over ONE copy of the butterfly network ------------------------
        // step 0: 16-point DFT over a (all four stages); Y_l[k0] lands in register brev4(k0)
        // step 1: G DFTs of L points (stages 1..R1) on the exchanged data
#pragma unroll 1
        for (uint32_t st = 0; st < (R1 ? 2u : 1u); st++) {
            const uint32_t nstages = st == 0 ? 4 : R1;
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++) {
                dft_stage<F, 1>(x[q], tb);
                if (nstages >= 2) dft_stage<F, 2>(x[q], tb);
                if (nstages >= 3) dft_stage<F, 3>(x[q], tb);
                if (nstages >= 4) dft_stage<F, 4>(x[q], tb);
            }
            if (R1 > 0 && st == 0) {
                // twist by w_(2^R)^(l * k0) in shared memory (one copy of the multiplier in a rolled
                // loop), then the 16 x L exchange: lane l takes over k0 = l + L*g, g < G
#pragma unroll
                for (uint32_t q = 0; q < CPT; q++)
#pragma unroll
                    for (uint32_t r = 0; r < 16; r++) xb[q * XW + cbrev(r, 4) * SK] = x[q][r];
                const T* mid = midsh + l;
#pragma unroll 5
                for (uint32_t k0 = 1; k0 < 16; k0++) {
                    const T w = mid[k0 * L];
#pragma unroll
                    for (uint32_t q = 0; q < CPT; q++) xb[q * XW + k0 * SK] = F::mul(xb[q * XW + k0 * SK], w);
                }
#pragma unroll
                for (uint32_t q = 0; q < CPT; q++) xb[q * XW] = F::tight(xb[q * XW]);
                __syncwarp();
                // natural input b of block g goes to logical position g*L + brev(b) of the network
#pragma unroll
                for (uint32_t q = 0; q < CPT; q++)
#pragma unroll
                    for (uint32_t g = 0; g < G; g++)
#pragma unroll
                        for (uint32_t b = 0; b < L; b++)
                            x[q][cbrev(g * L + cbrev(b, R1), 4)] = (xb - l)[q * XW + (l + L * g) * SK + b];
                __syncwarp();
            }
        }

        // ---- store: X[k], k = l + L*j, j = g + G*k1, sits in register brev4(g*L + k1) ------------
        // staged in this lane's own shared-memory slots so that twiddle generation, scaling and the
        // address walk run as one rolled loop; stored row: k, or brev_R(k) = (brev(l) << 4) + brev4(j)
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++)
#pragma unroll
            for (uint32_t j = 0; j < 16; j++) xb[q * XW + j * SK] = x[q][cbrev((j % G) * L + j / G, 4)];
        const uint32_t mode = TW == TW_STORE ? 0 : (d.scale && TW == TW_NONE) ? 1 : 2;
#pragma unroll 1
        for (uint32_t u = 0; u < 4; u++) {
#pragma unroll
          for (uint32_t v = 0; v < 4; v++) {
            const uint32_t j = 4 * u + v;
            const uint32_t jrev = __brev(j) >> 28;
            T y[CPT];
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++) {
                y[q] = xb[q * XW + j * SK];
                if constexpr (TW == TW_STORE) {
                    y[q] = F::mul(y[q], tw[q][v]);
                    if (u != 3) tw[q][v] = F::mul(tw[q][v], step4[q]);
                } else if (mode == 1) {
                    y[q] = F::mul(y[q], tb.ninv);
                } else {
                    y[q] = F::canon(y[q]);
                }
            }
            if (out_plain) {
                const uint32_t off = out_lane_off + ((d.out_rev ? jrev : j * L) << d.out_lg_sa);
                if constexpr (CPT == 2) {
                    if (out_vec) {
                        VecLoad<T, 2>::st(out + (obase[0] + off), y[0], y[1]);
                        continue;
                    }
                }
#pragma unroll
                for (uint32_t q = 0; q < CPT; q++)
                    if (active[q]) out[obase[q] + off] = y[q];
            } else {
                const uint32_t row = out_lane_row + (d.out_rev ? jrev : j * L);    // stored row
#pragma unroll
                for (uint32_t q = 0; q < CPT; q++) {
                    if (!active[q]) continue;
                    if (d.peer_on) {
                        // fused exchange: row belongs to rank row >> peer_shift, whose receive buffer
                        // is mapped into this process (NVLink peer memory)
                        const uint32_t vl = row & ((1u << d.peer_shift) - 1);
                        T* dst = reinterpret_cast<T*>(d.peer[row >> d.peer_shift]);
                        dst[d.peer_block + obase[q] + ((uint64_t)vl << d.out_lg_sa)] = y[q];
                    } else {
                        const uint64_t row_off = ((uint64_t)(row >> d.out_split_bits) << d.out_split_shift) +
                                                 ((uint64_t)(row & ((1u << d.out_split_bits) - 1)) << d.out_lg_sa);
                        out[obase[q] + row_off] = y[q];
                    }
                }
            }
        }
        }
    }
}

#endif  // __CUDACC__

}  // namespace ntt
