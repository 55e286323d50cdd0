// Polynomial helpers over the NTT fields: inclusive prefix sum / product, division by (x - z),
// multi-point evaluation, batch inversion.  SURVEY.md section 8 row f4: the reference's
// polynomial/{prefix_op,div_by_x_minus_z,evaluate}.cuh and ff/batch_inversion.hpp.
//
// The reference writes each of these as ONE cooperative kernel with grid-wide barriers between a
// local phase, a cross-block carry phase and an apply phase.  Here:
//  * prefix_op and div_by_x_minus_z are the same single-pass scan with decoupled look-back: tiles
//    are handed out by an atomic ticket, each tile publishes its aggregate and then its inclusive
//    prefix next to a flag word, and a warp looks back over 32 predecessors at a time.  One read
//    and one write of the data, no grid barrier, no cooperative launch.  Division by (x - z) is
//    the scan of  b[i] = c[i] + z * b[i+1]  from the top coefficient down: the carry that crosses
//    k elements is weighted by z^k, and since every level of the hierarchy (thread, lane, warp,
//    tile) spans a fixed number of elements the weights are a handful of constants held in
//    shared memory.
//  * evaluate is a weighted tree reduction (Horner per thread, then  sum_t s_t * (x^E)^t  by
//    shuffles) with one partial per CTA and a one-CTA finish per point.
//  * batch inversion shares ONE field inversion per CTA: prefix and suffix products over the CTA
//    give every thread the inverse of its own chunk product.
// All arithmetic is on the field's memory format (arith<F> below), the same the NTT entry points
// use, so buffers pass between NTT and these helpers unchanged.
#pragma once
#include "../ff/gl64.cuh"
#include "../ff/bb31.cuh"
#include "../ff/mont_ntt.cuh"
#include "../util/gpu.cuh"

namespace poly {

// ---- arithmetic on the memory format -------------------------------------------------------------
// "data" values are what sits in memory; "constants" (z, x and their powers) may live in another
// domain when that saves work: cmul(data, constant) -> data, kmul(constant, constant) -> constant.
template<class F> struct arith {                       // Montgomery fields: one closed domain
    typedef typename F::T T;
    static HD T zero() { T z{}; return z; }
    static HD T one() { return F::one(); }
    static HD T load(const T& a) { return F::load(a); }
    static HD T add(const T& a, const T& b) { return F::add(a, b); }
    static HD T dmul(const T& a, const T& b) { return F::mul(a, b); }
    static HD T cmul(const T& a, const T& c) { return F::mul(a, c); }
    static HD T konst(const T& a) { return F::load(a); }
    static HD T kone() { return F::one(); }
    static HD T kmul(const T& a, const T& b) { return F::mul(a, b); }
    static HD T inv(const T& a) { return F::inv(a); }
};
// Goldilocks: data are plain canonical words, constants are kept times 2^64 (see ff/gl64.cuh) so
// that data * constant is a single multiplication-free-reduction product
template<> struct arith<gl64> {
    typedef uint64_t T;
    static constexpr uint64_t R2 = 0xfffffffe00000001ULL;          // 2^128 mod p
    static HD T zero() { return 0; }
    static HD T one() { return 1; }
    static HD T load(T a) { return gl64::canon(a); }
    static HD T add(T a, T b) { return gl64::canon(gl64::add(a, b)); }
    static HD T dmul(T a, T b) { return gl64::mul(gl64::mul(a, b), R2); }
    static HD T cmul(T a, T c) { return gl64::mul(a, c); }
    static HD T konst(T a) { return gl64::to_mont(a); }
    static HD T kone() { return gl64::one(); }
    static HD T kmul(T a, T b) { return gl64::mul(a, b); }
    static HD T inv(T a) { return gl64::mul(gl64::pow(gl64::to_mont(a), gl64::P - 2), 1); }
};

HD bool is_zero(uint32_t a) { return a == 0; }
HD bool is_zero(uint64_t a) { return a == 0; }
template<class T> HD bool is_zero(const T& a)
{
    uint32_t acc = 0;
#pragma unroll
    for (size_t i = 0; i < sizeof(a.l) / sizeof(a.l[0]); i++) acc |= a.l[i];
    return acc == 0;
}

template<class F> HD typename F::T kpow(typename F::T c, uint64_t e)
{
    typedef arith<F> A;
    typename F::T r = A::kone();
    for (; e; e >>= 1, c = A::kmul(c, c))
        if (e & 1) r = A::kmul(r, c);
    return r;
}

#if defined(__CUDACC__)
// ---- whole-element warp shuffles and L2 loads (role of the reference's ff/shfl.cuh) --------------
DEV uint32_t shfl_up(uint32_t v, uint32_t d) { return __shfl_up_sync(0xffffffffu, v, d); }
DEV uint64_t shfl_up(uint64_t v, uint32_t d) { return __shfl_up_sync(0xffffffffu, v, d); }
DEV uint32_t shfl_down(uint32_t v, uint32_t d) { return __shfl_down_sync(0xffffffffu, v, d); }
DEV uint64_t shfl_down(uint64_t v, uint32_t d) { return __shfl_down_sync(0xffffffffu, v, d); }
DEV uint32_t shfl_idx(uint32_t v, uint32_t l) { return __shfl_sync(0xffffffffu, v, l); }
DEV uint64_t shfl_idx(uint64_t v, uint32_t l) { return __shfl_sync(0xffffffffu, v, l); }
template<class T> DEV T shfl_up(const T& v, uint32_t d)
{
    T r;
#pragma unroll
    for (size_t i = 0; i < sizeof(v.l) / 4; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
template<class T> DEV T shfl_down(const T& v, uint32_t d)
{
    T r;
#pragma unroll
    for (size_t i = 0; i < sizeof(v.l) / 4; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}
template<class T> DEV T shfl_idx(const T& v, uint32_t l)
{
    T r;
#pragma unroll
    for (size_t i = 0; i < sizeof(v.l) / 4; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], l);
    return r;
}
DEV uint32_t ld_cg(const uint32_t* p) { return __ldcg(p); }
DEV uint64_t ld_cg(const uint64_t* p) { return __ldcg((const unsigned long long*)p); }
template<class T> DEV T ld_cg(const T* p)
{
    static_assert(sizeof(T) % 16 == 0, "wide elements are 16-byte multiples");
    T r;
    const uint4* q = (const uint4*)p;
#pragma unroll
    for (size_t i = 0; i < sizeof(T) / 16; i++) ((uint4*)&r)[i] = __ldcg(q + i);
    return r;
}

enum { OP_ADD = 0, OP_MUL = 1, OP_DIV = 2 };

// carry (what precedes, in scan order) joined with x; w = z^(number of elements x spans), OP_DIV only
template<class F, int OP> DEV typename F::T join(const typename F::T& carry, const typename F::T& x,
                                                 const typename F::T& w)
{
    typedef arith<F> A;
    if (OP == OP_ADD) return A::add(carry, x);
    if (OP == OP_MUL) return A::dmul(carry, x);
    return A::add(x, A::cmul(carry, w));
}

// ---- single-pass scan ------------------------------------------------------------------------------
// Scan order t = 0, 1, ...: OP_ADD / OP_MUL walk memory upwards, OP_DIV walks it downwards from
// the top coefficient (memory index len-1-t).  Thread = E consecutive scan positions, warp = 32 E,
// tile = BS E; the last tile in scan order may be ragged (its tail is identity and is not stored).
// Tile status: flags[k] = 0 nothing yet, 1 agg[k] valid, 2 incl[k] valid.
template<class F, int OP, int E, int BS>
__global__ __launch_bounds__(BS) void scan_kernel(typename F::T* out, const typename F::T* in, size_t len,
                                                  typename F::T z, int rotate, uint32_t ntiles,
                                                  uint32_t* counter, volatile uint32_t* flags,
                                                  typename F::T* agg, typename F::T* incl,
                                                  typename F::T* edge)
{
    typedef arith<F> A;
    typedef typename F::T T;
    constexpr int NW = BS / 32;
    constexpr uint32_t TILE = BS * E;
    static_assert(BS >= 256 && NW <= 32, "setup below spreads the constant table over 256 threads");
    __shared__ T s_wl[33];          // z^(E k): a carry crossing k threads
    __shared__ T s_ww[NW + 1];      // z^(32 E k): crossing k warps
    __shared__ T s_zp[E + 1];       // z^k
    __shared__ T s_zt[6];           // z^(TILE 2^k): crossing 2^k tiles
    __shared__ T s_agg[NW];
    __shared__ T s_carry;
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T ident = OP == OP_MUL ? A::one() : A::zero();

    if (OP == OP_DIV) {
        if (tid < 33) s_wl[tid] = kpow<F>(z, (uint64_t)E * tid);
        else if (tid >= 64 && tid < 64 + NW + 1) s_ww[tid - 64] = kpow<F>(z, (uint64_t)32 * E * (tid - 64));
        else if (tid >= 128 && tid < 128 + E + 1) s_zp[tid - 128] = kpow<F>(z, tid - 128);
        else if (tid >= 192 && tid < 198) s_zt[tid - 192] = kpow<F>(z, (uint64_t)TILE << (tid - 192));
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;

        const size_t base = (size_t)tile * TILE + (size_t)tid * E;
        T v[E];
#pragma unroll
        for (int j = 0; j < E; j++) {
            size_t pos = base + j;
            v[j] = ident;
            if (pos < len) v[j] = A::load(in[OP == OP_DIV ? len - 1 - pos : pos]);
        }
#pragma unroll
        for (int j = 1; j < E; j++) v[j] = join<F, OP>(v[j - 1], v[j], z);

        T inc = v[E - 1];
#pragma unroll
        for (uint32_t off = 1; off < 32; off <<= 1) {
            T t = shfl_up(inc, off);
            if (lane >= off) inc = join<F, OP>(t, inc, s_wl[off]);
        }
        T lane_excl = shfl_up(inc, 1);
        if (lane == 0) lane_excl = ident;
        if (lane == 31) s_agg[warp] = inc;
        __syncthreads();

        if (warp == 0) {
            T w = lane < NW ? s_agg[lane] : ident;
#pragma unroll
            for (uint32_t off = 1; off < NW; off <<= 1) {
                T t = shfl_up(w, off);
                if (lane >= off) w = join<F, OP>(t, w, s_ww[off]);
            }
            const T tile_agg = shfl_idx(w, NW - 1);
            T acc = ident;                                   // lane 0: everything before this tile
            if (tile > 0) {
                if (lane == 0) {
                    agg[tile] = tile_agg;
                    __threadfence();
                    flags[tile] = 1;
                }
                T wacc = A::kone();                          // z^(TILE * tiles acc spans)
                for (int64_t k0 = (int64_t)tile - 1;; k0 -= 32) {
                    const int64_t k = k0 - lane;             // lane 0 looks at the nearest predecessor
                    uint32_t f;
                    do {
                        f = k < 0 ? 2u : flags[k];
                    } while (__any_sync(0xffffffffu, f == 0));
                    __threadfence();
                    T val = ident;
                    if (k >= 0) val = f == 2 ? ld_cg(incl + k) : ld_cg(agg + k);
                    const uint32_t m = __ballot_sync(0xffffffffu, f == 2);
                    const uint32_t jstar = __ffs(m) - 1;     // nearest inclusive prefix (m == 0: 0xffffffff)
                    if (lane > jstar) val = ident;
#pragma unroll
                    for (uint32_t kk = 0, off = 1; off < 32; kk++, off <<= 1) {
                        T t = shfl_down(val, off);
                        if (lane + off < 32) val = join<F, OP>(t, val, s_zt[kk]);
                    }
                    if (lane == 0) acc = join<F, OP>(val, acc, wacc);
                    if (m) break;
                    if (OP == OP_DIV) wacc = A::kmul(wacc, s_zt[5]);
                }
            }
            if (lane == 0) {
                s_carry = acc;
                if (tile + 1 < ntiles) {
                    incl[tile] = join<F, OP>(acc, tile_agg, s_zt[0]);
                    __threadfence();
                    flags[tile] = 2;
                }
            }
            if (lane < NW) s_agg[lane] = w;
        }
        __syncthreads();

        const T wexcl = warp ? s_agg[warp - 1] : ident;
        const T wc = join<F, OP>(s_carry, wexcl, s_ww[warp]);          // value entering this warp
        const T cin = join<F, OP>(wc, lane_excl, s_wl[lane]);          // value entering this thread
#pragma unroll
        for (int j = 0; j < E; j++) {
            size_t pos = base + j;
            T r = join<F, OP>(cin, v[j], s_zp[j + 1]);
            if (pos >= len) continue;
            if (OP != OP_DIV) out[pos] = r;
            else if (!rotate) out[len - 1 - pos] = r;
            else if (pos == len - 1) out[len - 1] = r;                 // the remainder goes last
            else if (j == E - 1 && tid == BS - 1) edge[tile] = r;      // slot still unread by the next tile
            else out[len - 2 - pos] = r;
        }
    }
}

// rotate=true: the quotient coefficient at a tile's last scan position lands in the first input
// slot of the NEXT tile, which that tile may not have read yet; it is parked and stored here
template<class T>
__global__ void scan_edge_kernel(T* out, const T* edge, size_t len, uint32_t ntiles, uint32_t tile_elems)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k + 1 >= ntiles) return;
    size_t pos = (size_t)(k + 1) * tile_elems - 1;
    out[len - 2 - pos] = edge[k];
}

// ---- sum_t val_t * base^t over the CTA (thread 0 holds the result) ---------------------------------
template<class F, int BS>
DEV typename F::T block_wreduce(typename F::T val, const typename F::T& base, typename F::T* s_w /*[10]*/,
                                typename F::T* s_x /*[BS/32]*/)
{
    typedef arith<F> A;
    typedef typename F::T T;
    constexpr int NW = BS / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();                                            // s_w / s_x free again
    if (tid < 10) {
        T b = base;
        for (uint32_t i = 0; i < tid; i++) b = A::kmul(b, b);   // base^(2^tid)
        s_w[tid] = b;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t k = 0, off = 1; off < 32; k++, off <<= 1) {
        T t = shfl_down(val, off);
        val = A::add(val, A::cmul(t, s_w[k]));
    }
    if (lane == 0) s_x[warp] = val;
    __syncthreads();
    if (warp == 0) {
        val = lane < NW ? s_x[lane] : A::zero();
#pragma unroll
        for (uint32_t k = 5, off = 1; off < NW; k++, off <<= 1) {
            T t = shfl_down(val, off);
            val = A::add(val, A::cmul(t, s_w[k]));
        }
    }
    return val;
}

// one partial per (point, CTA): sum over the CTA's BS*E coefficients of c_i x^(i - first)
template<class F, int E, int BS>
__global__ __launch_bounds__(BS) void evaluate_partial_kernel(typename F::T* partial, const typename F::T* x,
                                                              uint32_t npoints, const typename F::T* coeffs,
                                                              size_t len)
{
    typedef arith<F> A;
    typedef typename F::T T;
    __shared__ T s_w[10];
    __shared__ T s_x[BS / 32];
    const size_t base = ((size_t)blockIdx.x * BS + threadIdx.x) * E;
    T c[E];
#pragma unroll
    for (int j = 0; j < E; j++) {
        c[j] = A::zero();
        if (base + j < len) c[j] = A::load(coeffs[base + j]);
    }
    for (uint32_t p = 0; p < npoints; p++) {
        const T xk = A::konst(x[p]);
        T s = c[E - 1];
#pragma unroll
        for (int j = E - 2; j >= 0; j--) s = A::add(A::cmul(s, xk), c[j]);
        T r = block_wreduce<F, BS>(s, kpow<F>(xk, E), s_w, s_x);
        if (threadIdx.x == 0) partial[(size_t)p * gridDim.x + blockIdx.x] = r;
    }
}

// ret[p] = sum_k partial[p][k] * (x^tile_elems)^k, one CTA per point
template<class F, int BS>
__global__ __launch_bounds__(BS) void evaluate_finish_kernel(typename F::T* ret, const typename F::T* partial,
                                                             const typename F::T* x, uint32_t nparts,
                                                             uint32_t tile_elems)
{
    typedef arith<F> A;
    typedef typename F::T T;
    __shared__ T s_w[10];
    __shared__ T s_x[BS / 32];
    const uint32_t p = blockIdx.x, per = (nparts + BS - 1) / BS;
    const T xt = kpow<F>(A::konst(x[p]), tile_elems);
    const T* mine = partial + (size_t)p * nparts;
    T s = A::zero();
    for (uint32_t j = per; j-- > 0;) {
        uint32_t k = threadIdx.x * per + j;
        s = A::cmul(s, xt);
        if (k < nparts) s = A::add(s, mine[k]);
    }
    T r = block_wreduce<F, BS>(s, kpow<F>(xt, per), s_w, s_x);
    if (threadIdx.x == 0) ret[p] = r;
}

// ---- batch inversion -------------------------------------------------------------------------------
// ff/batch_inversion.hpp:14-51 for a caller's own kernel: out[i] = 1/inp[i], zero where inp[i] is
// zero, one field inversion for the N elements (Montgomery's trick; zeros are stepped over by
// multiplying with one instead, the reference's csel/czero)
template<class F, int N>
DEV void batch_inversion(typename F::T out[N], const typename F::T inp[N])
{
    typedef arith<F> A;
    typedef typename F::T T;
    T acc = A::one();
#pragma unroll
    for (int i = 0; i < N; i++) {
        out[i] = acc;                                          // product of what precedes
        acc = A::dmul(acc, is_zero(inp[i]) ? A::one() : inp[i]);
    }
    T inv = A::inv(acc);
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        const bool zero = is_zero(inp[i]);
        T next = A::dmul(inv, zero ? A::one() : inp[i]);
        out[i] = zero ? A::zero() : A::dmul(inv, out[i]);
        inv = next;
    }
}

// the array form: every CTA shares one inversion among its BS*N elements
template<class F, int N, int BS>
__global__ __launch_bounds__(BS) void batch_inverse_kernel(typename F::T* out, const typename F::T* in, size_t len)
{
    typedef arith<F> A;
    typedef typename F::T T;
    constexpr int NW = BS / 32;
    __shared__ T s_agg[NW], s_pre[NW], s_suf[NW];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t nchunks = (len + (size_t)BS * N - 1) / ((size_t)BS * N);

    for (size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const size_t base = (chunk * BS + tid) * N;
        T x[N], pre[N];
        T a = A::one();
#pragma unroll
        for (int i = 0; i < N; i++) {
            x[i] = A::one();
            if (base + i < len) x[i] = A::load(in[base + i]);
            pre[i] = a;
            a = A::dmul(a, is_zero(x[i]) ? A::one() : x[i]);
        }
        // product of every OTHER thread's chunk = (lanes before) (lanes after) (warps before) (warps after)
        T pi = a, si = a;
#pragma unroll
        for (uint32_t off = 1; off < 32; off <<= 1) {
            T t = shfl_up(pi, off);
            if (lane >= off) pi = A::dmul(t, pi);
            T u = shfl_down(si, off);
            if (lane + off < 32) si = A::dmul(u, si);
        }
        T before = shfl_up(pi, 1), after = shfl_down(si, 1);
        if (lane == 0) before = A::one();
        if (lane == 31) after = A::one();
        __syncthreads();                                        // previous chunk's s_* consumed
        if (lane == 31) s_agg[warp] = pi;
        __syncthreads();
        if (warp == 0) {
            T wa = lane < NW ? s_agg[lane] : A::one();
            T wp = wa, ws = wa;
#pragma unroll
            for (uint32_t off = 1; off < 32; off <<= 1) {
                T t = shfl_up(wp, off);
                if (lane >= off) wp = A::dmul(t, wp);
                T u = shfl_down(ws, off);
                if (lane + off < 32) ws = A::dmul(u, ws);
            }
            const T total_inv = A::inv(shfl_idx(wp, 31));       // the one inversion
            T wb = shfl_up(wp, 1), wf = shfl_down(ws, 1);
            if (lane == 0) wb = A::one();
            if (lane == 31) wf = A::one();
            if (lane < NW) {
                s_pre[lane] = wb;
                s_suf[lane] = A::dmul(wf, total_inv);
            }
        }
        __syncthreads();
        T inv = A::dmul(A::dmul(before, after), A::dmul(s_pre[warp], s_suf[warp]));   // 1 / a
#pragma unroll
        for (int i = N - 1; i >= 0; i--) {
            const bool zero = is_zero(x[i]);
            T r = zero ? A::zero() : A::dmul(inv, pre[i]);
            inv = A::dmul(inv, zero ? A::one() : x[i]);
            if (base + i < len) out[base + i] = r;
        }
    }
}
#endif  // __CUDACC__

}  // namespace poly
