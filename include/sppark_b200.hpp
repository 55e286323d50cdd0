// sppark_b200.hpp -- the reference's C++ surface on top of the C ABI (include/sppark_b200.h).
//
// Host-only, header-only C++17; needs neither nvcc nor the CUDA headers.  It keeps the names a
// C++ caller of supranational/sppark uses ("C++-level names to keep"):
//   field / curve value types  fp_t, fr_t, fp2_t, Affine_t, Affine_inf_t, jacobian_t, xyzz_t
//                              (ff/bls12-381.hpp:91-139, ff/pasta.hpp:82-103, ff/goldilocks.hpp,
//                               ff/baby_bear.hpp, ec/affine_t.hpp:19-122, ec/jacobian_t.hpp:16-58,
//                               ec/xyzz_t.hpp:14-101)
//   mult_pippenger<bucket_t>() msm/pippenger.cuh:730-747
//   msm_t<...>::invoke()       msm/pippenger.cuh:351-395,448-571 (host-pointer overloads)
//   NTT::Base / Base_dev_ptr / LDE / LDE_aux / LDE_powers / LDE_expand, InputOutputOrder, Direction, Type
//                              ntt/ntt.cuh:33-36,216-244,283-350
//   gpu_t, stream_t, select_gpu(), ngpus(), cuda_available()   util/gpu_t.cuh:20-24,57-267
// The value types are LAYOUT types (the bytes that cross the boundary); arithmetic happens on the
// GPU.  Select the field set as the reference does, with -DFEATURE_BLS12_381 / FEATURE_PALLAS /
// FEATURE_VESTA / FEATURE_BN254 / FEATURE_BLS12_377 / FEATURE_GOLDILOCKS / FEATURE_BABY_BEAR; include/compat/ holds forwarding headers
// under the reference's own file names so that its poc glue (poc/msm-cuda/cuda/pippenger_inf.cu,
// poc/ntt-cuda/cuda/ntt_api.cu) compiles unmodified against this library.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include "sppark_b200.h"

#ifndef SPPARK_FFI
# define SPPARK_FFI extern "C" __attribute__((visibility("default")))
#endif

namespace sppark_b200 {

// A field element as it lies in memory: WORDS 32-bit little-endian limbs (Montgomery form for the
// mont_t fields, canonical for Goldilocks).  NTT_FIELD / MSM_CURVE are the C-ABI selectors of the
// transforms defined over it (-1: none).
template<int TAG, size_t WORDS, int NTT_FIELD, int MSM_CURVE>
struct alignas(WORDS % 2 ? 4 : 8) felem_t {
    uint32_t limb[WORDS];
    static constexpr size_t n = WORDS;
    static constexpr int ntt_field = NTT_FIELD, msm_curve = MSM_CURVE;
    static constexpr unsigned degree = 1;
    using mem_t = felem_t;
    felem_t() = default;
    explicit felem_t(const uint32_t* p) { memcpy(limb, p, sizeof(limb)); }
    explicit felem_t(uint64_t v) { memset(limb, 0, sizeof(limb)); memcpy(limb, &v, WORDS == 1 ? 4 : 8); }
    uint32_t& operator[](size_t i) { return limb[i]; }
    const uint32_t& operator[](size_t i) const { return limb[i]; }
    void zero() { memset(limb, 0, sizeof(limb)); }
    bool is_zero() const
    {
        uint32_t acc = 0;
        for (size_t i = 0; i < WORDS; i++) acc |= limb[i];
        return acc == 0;
    }
    friend bool operator==(const felem_t& a, const felem_t& b) { return memcmp(a.limb, b.limb, sizeof(a.limb)) == 0; }
    friend bool operator!=(const felem_t& a, const felem_t& b) { return !(a == b); }
};

}  // namespace sppark_b200

// ---- field sets, chosen like the reference's ff/*.hpp ---------------------------------------
#if defined(FEATURE_BLS12_381)
typedef sppark_b200::felem_t<0, 12, -1, SPPARK_CURVE_BLS12_381_G1> fp_t;
typedef sppark_b200::felem_t<1, 8, SPPARK_FIELD_BLS12_381_FR, -1> fr_t;
typedef sppark_b200::felem_t<2, 24, -1, SPPARK_CURVE_BLS12_381_G2> fp2_t;          // (c0, c1)
#elif defined(FEATURE_PALLAS)
typedef sppark_b200::felem_t<3, 8, -1, SPPARK_CURVE_PALLAS> fp_t;                   // pallas_t
typedef sppark_b200::felem_t<4, 8, SPPARK_FIELD_PALLAS_FR, -1> fr_t;                // vesta_t
#elif defined(FEATURE_VESTA)
typedef sppark_b200::felem_t<5, 8, -1, SPPARK_CURVE_VESTA> fp_t;                    // vesta_t
typedef sppark_b200::felem_t<6, 8, SPPARK_FIELD_VESTA_FR, -1> fr_t;                 // pallas_t
#elif defined(FEATURE_GOLDILOCKS)
typedef sppark_b200::felem_t<7, 2, SPPARK_FIELD_GL64, -1> fr_t;                     // gl64_t
typedef fr_t gl64_t;
#elif defined(FEATURE_BABY_BEAR)
typedef sppark_b200::felem_t<8, 1, SPPARK_FIELD_BB31, -1> fr_t;                     // bb31_t
typedef fr_t bb31_t;
#elif defined(FEATURE_BN254)
typedef sppark_b200::felem_t<9, 8, -1, SPPARK_CURVE_BN254_G1> fp_t;                // ff/alt_bn128.hpp
typedef sppark_b200::felem_t<10, 8, SPPARK_FIELD_BN254_FR, -1> fr_t;
typedef sppark_b200::felem_t<13, 16, -1, SPPARK_CURVE_BN254_G2> fp2_t;             // ff/alt_bn128-fp2.hpp, (c0, c1)
#elif defined(FEATURE_BLS12_377)
typedef sppark_b200::felem_t<11, 12, -1, SPPARK_CURVE_BLS12_377_G1> fp_t;          // ff/bls12-377.hpp
typedef sppark_b200::felem_t<12, 8, SPPARK_FIELD_BLS12_377_FR, -1> fr_t;
typedef sppark_b200::felem_t<14, 24, -1, SPPARK_CURVE_BLS12_377_G2> fp2_t;         // ff/bls12-377-fp2.hpp, (c0, c1)
#elif defined(FEATURE_MERSENNE31)
# error "sppark_b200: Mersenne31 is not instantiated in this library (DESIGN.md section 1)"
#endif

// ---- curve point layouts (ec/*.hpp) -----------------------------------------------------------
template<class field_t> struct Affine_t {
    field_t X, Y;                                          // infinity: X == Y == 0
    bool is_inf() const { return X.is_zero() && Y.is_zero(); }
};
template<class field_t> struct Affine_inf_t {
    field_t X, Y;
    bool inf;                                              // arkworks GroupAffine { x, y, infinity }
    bool is_inf() const { return inf; }
};
template<class field_t> struct jacobian_t {
    field_t X, Y, Z;                                       // infinity: Z == 0
    using affine_t = Affine_t<field_t>;
    using affine_inf_t = Affine_inf_t<field_t>;
    void inf() { memset(this, 0, sizeof(*this)); }
    bool is_inf() const { return Z.is_zero(); }
};
template<class field_t> struct xyzz_t {
    field_t X, Y, ZZZ, ZZ;                                 // member order of ec/xyzz_t.hpp:17
    using affine_t = Affine_t<field_t>;
    using affine_inf_t = Affine_inf_t<field_t>;
    using mem_t = xyzz_t;
    static constexpr unsigned degree = field_t::degree;
    void inf() { memset(this, 0, sizeof(*this)); }
    bool is_inf() const { return ZZZ.is_zero() && ZZ.is_zero(); }
};

// ---- devices and streams (util/gpu_t.cuh) -------------------------------------------------------
class stream_t {
    void* stream;                                          // a cudaStream_t owned by the caller
public:
    explicit stream_t(void* cuda_stream = nullptr) : stream(cuda_stream) {}
    operator void*() const { return stream; }
};

class gpu_t {
    int gpu_id;
public:
    explicit gpu_t(int id) : gpu_id(id) {}
    int id() const { return gpu_id; }
    int cid() const { return gpu_id; }
    int sm_count() const { return sppark_b200_sm_count(gpu_id); }
};
inline size_t ngpus() { return sppark_b200_ngpus(); }
// id = -1: the calling thread's current device (util/all_gpus.cpp:44-50)
inline const gpu_t& select_gpu(int id = 0)
{
    static const gpu_t gpus[] = {gpu_t{-1}, gpu_t{0}, gpu_t{1}, gpu_t{2}, gpu_t{3}, gpu_t{4}, gpu_t{5}, gpu_t{6},
                                 gpu_t{7}, gpu_t{8}, gpu_t{9}, gpu_t{10}, gpu_t{11}, gpu_t{12}, gpu_t{13},
                                 gpu_t{14}, gpu_t{15}};
    return gpus[id < -1 || id > 15 ? 1 : id + 1];
}

// ---- MSM (msm/pippenger.cuh) ----------------------------------------------------------------------
template<class bucket_t, class point_t, class affine_t, class scalar_t>
class msm_t {
    int device_id;
    sppark_b200_msm_ctx* ctx = nullptr;                    // preloaded points, if any
    size_t preloaded = 0;
    RustError ctor_error{0, nullptr};
    msm_t(const msm_t&) = delete;
    msm_t& operator=(const msm_t&) = delete;
public:
    // the reference pre-sizes its scratch from npoints; scratch here is allocated per call from the
    // stream-ordered pool, so these constructors only record the device
    msm_t(std::nullptr_t = nullptr, size_t /*npoints*/ = 0, int device = -1) : device_id(device) {}
    explicit msm_t(size_t /*npoints*/, int device = -1) : device_id(device) {}
    // points preloaded on the (current) device: msm_t{points, npoints[, ffi_affine_sz]}
    // (msm/pippenger.cuh:377-380), then invoke(out, scalars[, mont]) per scalar vector
    msm_t(const affine_t points[], size_t npoints, size_t ffi_affine_sz = sizeof(affine_t), int device = -1)
        : device_id(device), preloaded(npoints)
    {
        typedef std::remove_cv_t<std::remove_reference_t<decltype(points[0].X)>> field_t;
        if constexpr (field_t::msm_curve >= 0)
            ctor_error = sppark_b200_msm_ctx_create(field_t::msm_curve, points, npoints, ffi_affine_sz, &ctx);
    }
    ~msm_t() { sppark_b200_msm_ctx_free(ctx); if (ctor_error.message) drop_error_message(ctor_error.message); }

    RustError invoke(point_t& out, const scalar_t scalars[], bool mont = true)
    {   return invoke(out, preloaded, scalars, mont);   }
    RustError invoke(point_t& out, size_t npoints, const scalar_t scalars[], bool mont = true)
    {
        if (ctx == nullptr) {
            out.inf();
            return ctor_error.code ? RustError{ctor_error.code, ctor_error.message ? strdup(ctor_error.message) : nullptr}
                                   : RustError{-1, strdup("sppark_b200: msm_t has no preloaded points")};
        }
        return sppark_b200_msm_ctx_invoke(ctx, &out, scalars, npoints, mont);
    }

    RustError invoke(point_t& out, const affine_t points[], size_t npoints, const scalar_t scalars[],
                     bool mont = true, size_t ffi_affine_sz = sizeof(affine_t))
    {
        typedef decltype(points[0].X) fe_ref;
        typedef std::remove_cv_t<std::remove_reference_t<fe_ref>> field_t;
        static_assert(sizeof(scalar_t) == 32, "scalars are 256-bit");
        static_assert(sizeof(point_t) == 3 * sizeof(field_t) && sizeof(bucket_t) == 4 * sizeof(field_t), "layout");
        (void)device_id;                                   // the MSM runs on the caller's current device
        if constexpr (field_t::msm_curve < 0) {            // a field no curve of this library is defined over
            out.inf();
            return RustError{-1, strdup("sppark_b200: no MSM is instantiated over this field")};
        } else {
            return sppark_b200_msm_ex(field_t::msm_curve, &out, points, npoints, scalars, ffi_affine_sz, mont);
        }
    }
    RustError invoke(point_t& out, const affine_t points[], size_t npoints, const scalar_t scalars[],
                     bool mont, size_t ffi_affine_sz, std::nullptr_t) = delete;
};

template<class bucket_t, class point_t, class affine_t, class scalar_t>
static RustError mult_pippenger(point_t* out, const affine_t points[], size_t npoints,
                                const scalar_t scalars[], bool mont = true,
                                size_t ffi_affine_sz = sizeof(affine_t))
{
    msm_t<bucket_t, point_t, affine_t, scalar_t> msm{nullptr, npoints};
    return msm.invoke(*out, points, npoints, scalars, mont, ffi_affine_sz);
}

// ---- NTT (ntt/ntt.cuh) -------------------------------------------------------------------------------
#if defined(FEATURE_BLS12_381) || defined(FEATURE_PALLAS) || defined(FEATURE_VESTA) || \
    defined(FEATURE_GOLDILOCKS) || defined(FEATURE_BABY_BEAR) || defined(FEATURE_BN254) || \
    defined(FEATURE_BLS12_377)
class NTT {
public:
    enum class InputOutputOrder { NN, NR, RN, RR };
    enum class Direction { forward, inverse };
    enum class Type { standard, coset };
    enum class Algorithm { GS, CT };

    // in place on HOST memory
    static RustError Base(const gpu_t& gpu, fr_t* inout, uint32_t lg_domain_size, InputOutputOrder order,
                          Direction direction, Type type)
    {
        return sppark_b200_ntt(fr_t::ntt_field, (size_t)(gpu.id() < 0 ? 0 : gpu.id()), inout, lg_domain_size,
                               (int)order, (int)direction, (int)type);
    }
    // in place on DEVICE memory, enqueued on the caller's stream (the reference returns void and
    // throws; the error is returned here)
    static RustError Base_dev_ptr(stream_t& stream, fr_t* d_inout, uint32_t lg_domain_size,
                                  InputOutputOrder order, Direction direction, Type type)
    {
        return sppark_b200_ntt_dev(fr_t::ntt_field, d_inout, lg_domain_size, (int)order, (int)direction,
                                   (int)type, (void*)stream);
    }
    static RustError LDE_aux(const gpu_t& gpu, fr_t* inout, uint32_t lg_domain_size, uint32_t lg_blowup,
                             fr_t* aux_out = nullptr)
    {
        return sppark_b200_lde(fr_t::ntt_field, (size_t)(gpu.id() < 0 ? 0 : gpu.id()), inout, lg_domain_size,
                               lg_blowup, aux_out);
    }
    static RustError LDE(const gpu_t& gpu, fr_t* inout, uint32_t lg_domain_size, uint32_t lg_blowup)
    {   return LDE_aux(gpu, inout, lg_domain_size, lg_blowup);   }
    // device-pointer halves (ntt/ntt.cuh:352-365; errors are returned, not thrown)
    static RustError LDE_powers(stream_t& stream, fr_t* d_inout, uint32_t lg_domain_size)
    {   return sppark_b200_lde_powers_dev(fr_t::ntt_field, d_inout, lg_domain_size, (void*)stream);   }
    static RustError LDE_expand(stream_t& stream, fr_t* d_out, fr_t* d_in, uint32_t lg_domain_size, uint32_t lg_blowup)
    {   return sppark_b200_lde_expand_dev(fr_t::ntt_field, d_out, d_in, lg_domain_size, lg_blowup, (void*)stream);   }
};

// ---- polynomial/ and ff/batch_inversion.hpp -----------------------------------------------------------
// Same names and argument order as the reference's templates (polynomial/prefix_op.cuh:17-45,322;
// div_by_x_minus_z.cuh:445; evaluate.cuh:308,416); device arrays, enqueued on `s`.  The reference
// returns void and throws cuda_error; the error is returned here.
template<typename T_> struct Add { using T = T_; static constexpr int op = 0; };
template<typename T_> struct Multiply { using T = T_; static constexpr int op = 1; };

template<class Operation, typename T = typename Operation::T>
inline RustError prefix_op(T* d_out, const T* d_inp, size_t len, stream_t& s)
{   return sppark_b200_prefix_op_dev(T::ntt_field, Operation::op, d_out, d_inp, len, (void*)s);   }

template<bool rotate = false, typename T>
inline RustError div_by_x_minus_z(T d_inout[], size_t len, const T& z, stream_t& s)
{   return sppark_b200_div_by_x_minus_z_dev(T::ntt_field, d_inout, len, &z, rotate, (void*)s);   }

template<typename T>
inline RustError evaluate(T d_ret[], const T d_x[], size_t n, const T d_coeffs[], size_t len, stream_t& s)
{   return sppark_b200_evaluate_dev(T::ntt_field, d_ret, d_x, n, d_coeffs, len, (void*)s);   }

// the array form of batch_inversion<T, N>() (ff/batch_inversion.hpp:14): one inversion per CTA
template<typename T>
inline RustError batch_inversion(T d_out[], const T d_inp[], size_t len, stream_t& s)
{   return sppark_b200_batch_inverse_dev(T::ntt_field, d_out, d_inp, len, (void*)s);   }
#endif
