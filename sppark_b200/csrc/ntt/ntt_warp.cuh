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

// In-place radix-2 DIT on registers x[BASE .. BASE + 2^LG): register BASE+a holds input a on
// entry, X[k] is left in register BASE + brev(k).  Twiddles: w_(2^s)^j = w16^(j * 16 / 2^s).
// TIGHT_IN: the inputs are legal second operands of add/sub already (canonical).
template<class F, uint32_t LG, uint32_t BASE, bool TIGHT_IN, uint32_t S, uint32_t NREG>
__device__ __forceinline__ void dft_stage(typename F::T (&x)[NREG], const Tables<F>& tb)
{
    typedef typename F::T T;
    constexpr uint32_t N = 1u << LG, half = 1u << (S - 1);
#pragma unroll
    for (uint32_t k = 0; k < N; k += 2 * half) {
#pragma unroll
        for (uint32_t j = 0; j < half; j++) {
            const uint32_t i0 = BASE + cbrev(k + j, LG), i1 = BASE + cbrev(k + j + half, LG);
            T t;
            if (j == 0) t = (S == 1 && TIGHT_IN) ? x[i1] : F::tight(x[i1]);
            else t = F::mul(x[i1], tb.w16[j * (8u >> (S - 1))]);
            const T u = x[i0];
            x[i0] = F::add(u, t);
            x[i1] = F::sub(u, t);
        }
    }
}
template<class F, uint32_t LG, uint32_t BASE, bool TIGHT_IN, uint32_t NREG>
__device__ __forceinline__ void dft_regs(typename F::T (&x)[NREG], const Tables<F>& tb)
{
    if constexpr (LG >= 1) dft_stage<F, LG, BASE, TIGHT_IN, 1>(x, tb);
    if constexpr (LG >= 2) dft_stage<F, LG, BASE, TIGHT_IN, 2>(x, tb);
    if constexpr (LG >= 3) dft_stage<F, LG, BASE, TIGHT_IN, 3>(x, tb);
    if constexpr (LG >= 4) dft_stage<F, LG, BASE, TIGHT_IN, 4>(x, tb);
}

// w_N^e from the two-level table, branch-free (thi[0] = 1)
template<class F> __device__ __forceinline__ typename F::T twiddle2(const Tables<F>& tb, uint32_t e)
{   return F::mul(tb.tlo[e & ((1u << LG_TLO) - 1)], tb.thi[e >> LG_TLO]);   }

template<class T, uint32_t CPT> struct VecLoad;
template<> struct VecLoad<uint64_t, 2> {
    static __device__ __forceinline__ void ld(const uint64_t* p, uint64_t& a, uint64_t& b)
    {   const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p); a = v.x; b = v.y;   }
    static __device__ __forceinline__ void st(uint64_t* p, uint64_t a, uint64_t b)
    {   *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);   }
};
template<> struct VecLoad<uint32_t, 2> {
    static __device__ __forceinline__ void ld(const uint32_t* p, uint32_t& a, uint32_t& b)
    {   const uint2 v = *reinterpret_cast<const uint2*>(p); a = v.x; b = v.y;   }
    static __device__ __forceinline__ void st(uint32_t* p, uint32_t a, uint32_t b)
    {   *reinterpret_cast<uint2*>(p) = make_uint2(a, b);   }
};

// shared-memory words one warp needs for its exchange (per column slot)
HD constexpr uint32_t warp_xchg_words(uint32_t R)
{
    return R <= 4 ? 0 : (32u >> (R - 4)) * (16u * ((1u << (R - 4)) + 1) + (1u << (R - 4)));
}

#ifndef SPPARK_B200_NTT_WARP_MINB
#define SPPARK_B200_NTT_WARP_MINB 3
#endif
template<class F, uint32_t R, uint32_t CPT>
__global__ void __launch_bounds__(256, CPT == 1 ? SPPARK_B200_NTT_WARP_MINB : 2)
pass_kernel_warp(const Pass d, const Tables<F> tb, const typename F::T* __restrict__ in,
                 typename F::T* __restrict__ out, uint32_t ncols)
{
    typedef typename F::T T;
    static_assert(R >= WARP_MIN_LG_R && R <= WARP_MAX_LG_R, "sub-NTT size");
    constexpr uint32_t R1 = R - 4, L = 1u << R1, G = 16u >> R1, SPW = 32u >> R1;
    constexpr uint32_t SK = L + 1, SS = 16 * (L + 1) + L;      // conflict-free strides of the exchange
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t s = lane >> R1, l = lane & (L - 1);
    const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + warp;
    if (unit * (SPW * CPT) >= ncols) return;                   // warp-uniform
    const uint32_t gc0 = (unit * SPW + s) * CPT;               // first of this lane's columns
    const uint32_t lrev = brev32(l, R1);

    // All element indices fit 32 bits (transforms of at most 2^30 elements).  Lanes beyond the last
    // column (transforms with fewer columns than a warp holds) read the last column and store
    // nothing.
    T x[CPT][16];
    bool active[CPT];
    uint32_t ibase[CPT], obase[CPT];
#pragma unroll
    for (uint32_t q = 0; q < CPT; q++) {
        uint32_t gc = gc0 + q;
        active[q] = gc < ncols;
        gc = active[q] ? gc : ncols - 1;
        const uint32_t t = gc >> d.lg_w, c = gc & ((1u << d.lg_w) - 1);
        ibase[q] = (uint32_t)tile_base(t, d.in_lg_tlo, d.in_tl, d.in_th) + (c << d.in_lg_sc);
        obase[q] = (uint32_t)tile_base(t, d.out_lg_tlo, d.out_tl, d.out_th) + (c << d.out_lg_sc);
    }

    // ---- load: register a <- natural row a*L + l, stored at row brev_R(a*L + l) if in_rev -----
    {
        const uint32_t lane_off = (d.in_rev ? (lrev << 4) : l) << d.in_lg_sa;
        const uint32_t nat_step = L << d.in_lg_sa, rev_step = 1u << d.in_lg_sa;
        const bool vec = CPT == 2 && d.in_lg_sc == 0 && d.lg_w >= 1 && d.in_lg_sa >= 1 && active[CPT - 1];
#pragma unroll
        for (uint32_t a = 0; a < 16; a++) {
            const uint32_t off = lane_off + (d.in_rev ? cbrev(a, 4) * rev_step : a * nat_step);
            if constexpr (CPT == 2) {
                if (vec) {
                    VecLoad<T, 2>::ld(in + (ibase[0] + off), x[0][a], x[1][a]);
                    continue;
                }
            }
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++) x[q][a] = in[ibase[q] + off];
        }
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++)
#pragma unroll
            for (uint32_t a = 0; a < 16; a++) x[q][a] = F::load(x[q][a]);
    }

    // ---- inter-pass twiddle at load (RN plans): x[a] *= w^((a*L + l) * colv << lsh) ------------
    if (d.tw_mode == TW_LOAD) {
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) {
            const uint32_t colv = tw_column_value(d, ibase[q]);
            T tw = twiddle2<F>(tb, (l * colv) << d.tw_lsh);
            const T step = twiddle2<F>(tb, (L * colv) << d.tw_lsh);
#pragma unroll
            for (uint32_t a = 0; a < 16; a++) {
                x[q][a] = F::mul(x[q][a], tw);
                if (a != 15) tw = F::mul(tw, step);
            }
        }
    }

    // ---- step 0: 16-point DFT over a; register r now holds Y_l[k0], k0 = brev4(r) -----------
#pragma unroll
    for (uint32_t q = 0; q < CPT; q++)
        dft_regs<F, 4, 0, false>(x[q], tb);

    if constexpr (R1 > 0) {
        // ---- twist by w_(2^R)^(l * k0), then the 16 x L exchange ---------------------------------
        const T* mid = tb.mid + mid_offset(R);
        T* buf = reinterpret_cast<T*>(smem_raw) + warp * (CPT * SPW * SS);
#pragma unroll
        for (uint32_t r = 1; r < 16; r++) {
            const T w = mid[cbrev(r, 4) * L + l];
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++) x[q][r] = F::mul(x[q][r], w);
        }
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) {
            x[q][0] = F::tight(x[q][0]);
            T* bq = buf + q * (SPW * SS) + s * SS;
#pragma unroll
            for (uint32_t r = 0; r < 16; r++) bq[cbrev(r, 4) * SK + l] = x[q][r];
        }
        __syncwarp();
        // lane l now owns k0 = l + L*g, g < G; register g*L + b <- Z_b[k0]
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) {
            const T* bq = buf + q * (SPW * SS) + s * SS;
#pragma unroll
            for (uint32_t g = 0; g < G; g++)
#pragma unroll
                for (uint32_t b = 0; b < L; b++) x[q][g * L + b] = bq[(l + L * g) * SK + b];
        }
        // ---- step 1: G DFTs of L points; register g*L + r holds X[k0 + 16*brev(r)] -----------
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) {
            if constexpr (G == 1) dft_regs<F, R1, 0, true>(x[q], tb);
            if constexpr (G >= 2) { dft_regs<F, R1, 0, true>(x[q], tb); dft_regs<F, R1, L, true>(x[q], tb); }
            if constexpr (G >= 4) { dft_regs<F, R1, 2 * L, true>(x[q], tb); dft_regs<F, R1, 3 * L, true>(x[q], tb); }
            if constexpr (G >= 8) {
                dft_regs<F, R1, 4 * L, true>(x[q], tb); dft_regs<F, R1, 5 * L, true>(x[q], tb);
                dft_regs<F, R1, 6 * L, true>(x[q], tb); dft_regs<F, R1, 7 * L, true>(x[q], tb);
            }
        }
    }

    // ---- store: natural output row k = l + L*j, j = 0..15 -------------------------------------
    // register of j: R1 > 0: g*L + brev(k1) with j = g + G*k1;  R1 == 0: brev4(j)
    // stored row:    k, or brev_R(k) = (brev(l) << 4) + (brev(g) << R1) + brev(k1) if out_rev
    const uint32_t lane_row = d.out_rev ? (lrev << 4) : l;
    const uint32_t lane_off = lane_row << d.out_lg_sa;
    const uint32_t nat_step = L << d.out_lg_sa, rev_step = 1u << d.out_lg_sa;
    const bool plain = !d.peer_on && !d.out_split_bits;
    const bool vec = CPT == 2 && d.out_lg_sc == 0 && d.lg_w >= 1 && d.out_lg_sa >= 1 && active[CPT - 1] && plain;
    T tw[CPT], step[CPT];
    const bool twisted = d.tw_mode == TW_STORE;
    if (twisted) {
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) {
            const uint32_t colv = tw_column_value(d, ibase[q]);
            tw[q] = twiddle2<F>(tb, (l * colv) << d.tw_lsh);
            step[q] = twiddle2<F>(tb, (L * colv) << d.tw_lsh);
            if (d.scale) tw[q] = F::mul(tw[q], tb.ninv);
        }
    }
#pragma unroll
    for (uint32_t j = 0; j < 16; j++) {
        const uint32_t g = j % G, k1 = j / G;
        const uint32_t reg = R1 > 0 ? g * L + cbrev(k1, R1) : cbrev(j, 4);
        const uint32_t vrev = R1 > 0 ? (cbrev(g, 4 - R1) << R1) + cbrev(k1, R1) : cbrev(j, 4);
        T y[CPT];
#pragma unroll
        for (uint32_t q = 0; q < CPT; q++) {
            y[q] = x[q][reg];
            if (twisted) {
                y[q] = F::mul(y[q], tw[q]);
                if (j != 15) tw[q] = F::mul(tw[q], step[q]);
            } else if (d.scale) {
                y[q] = F::mul(y[q], tb.ninv);
            } else {
                y[q] = F::canon(y[q]);
            }
        }
        if (plain) {
            const uint32_t off = lane_off + (d.out_rev ? vrev * rev_step : j * nat_step);
            if constexpr (CPT == 2) {
                if (vec) {
                    VecLoad<T, 2>::st(out + (obase[0] + off), y[0], y[1]);
                    continue;
                }
            }
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++)
                if (active[q]) out[obase[q] + off] = y[q];
        } else {
            const uint32_t v = lane_row + (d.out_rev ? vrev : L * j);      // stored row
#pragma unroll
            for (uint32_t q = 0; q < CPT; q++) {
                if (!active[q]) continue;
                if (d.peer_on) {
                    // fused exchange: row v belongs to rank v >> peer_shift; one of at most 8 peers,
                    // selected without indexing the parameter array dynamically
                    const uint32_t vl = v & ((1u << d.peer_shift) - 1), dst_rank = v >> d.peer_shift;
                    uint64_t base = d.peer[0];
#pragma unroll
                    for (uint32_t r = 1; r < 8; r++) base = dst_rank == r ? d.peer[r] : base;
                    reinterpret_cast<T*>(base)[d.peer_block + obase[q] + ((uint64_t)vl << d.out_lg_sa)] = y[q];
                } else {
                    const uint64_t row_off = ((uint64_t)(v >> d.out_split_bits) << d.out_split_shift) +
                                             ((uint64_t)(v & ((1u << d.out_split_bits) - 1)) << d.out_lg_sa);
                    out[obase[q] + row_off] = y[q];
                }
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace ntt
