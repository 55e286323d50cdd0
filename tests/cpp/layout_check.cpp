// Compile-time check of the C++ layer's layouts against the reference's ABI:
// built with -fsyntax-only once per FEATURE_* by tests/test_cpp_layer.py.
#include "sppark_b200.hpp"

#if defined(FEATURE_BLS12_381)
static_assert(sizeof(fp_t) == 48 && sizeof(fr_t) == 32 && sizeof(fp2_t) == 96, "field sizes");
static_assert(sizeof(Affine_t<fp_t>) == 96, "blst_p1_affine");
static_assert(sizeof(Affine_inf_t<fp_t>) == 104 && offsetof(Affine_inf_t<fp_t>, inf) == 96, "ark G1Affine");
static_assert(sizeof(jacobian_t<fp_t>) == 144 && sizeof(xyzz_t<fp_t>) == 192, "blst_p1 / bucket");
static_assert(sizeof(Affine_inf_t<fp2_t>) == 200 && offsetof(Affine_inf_t<fp2_t>, inf) == 192, "ark G2Affine");
static_assert(sizeof(jacobian_t<fp2_t>) == 288, "ark G2Projective");
static_assert(offsetof(xyzz_t<fp_t>, ZZZ) == 96 && offsetof(xyzz_t<fp_t>, ZZ) == 144, "xyzz member order");
#elif defined(FEATURE_PALLAS) || defined(FEATURE_VESTA)
static_assert(sizeof(fp_t) == 32 && sizeof(fr_t) == 32, "field sizes");
static_assert(sizeof(Affine_t<fp_t>) == 64 && sizeof(Affine_inf_t<fp_t>) == 72, "affine");
static_assert(sizeof(jacobian_t<fp_t>) == 96 && sizeof(xyzz_t<fp_t>) == 128, "points");
#elif defined(FEATURE_BN254)
static_assert(sizeof(fp_t) == 32 && sizeof(fr_t) == 32 && sizeof(Affine_inf_t<fp_t>) == 72, "BN254");
static_assert(sizeof(jacobian_t<fp_t>) == 96, "G1Projective");
#elif defined(FEATURE_BLS12_377)
static_assert(sizeof(fp_t) == 48 && sizeof(fr_t) == 32 && sizeof(Affine_inf_t<fp_t>) == 104, "BLS12-377");
static_assert(sizeof(jacobian_t<fp_t>) == 144, "G1Projective");
#elif defined(FEATURE_GOLDILOCKS)
static_assert(sizeof(fr_t) == 8 && alignof(fr_t) == 8, "gl64_t");
#elif defined(FEATURE_BABY_BEAR)
static_assert(sizeof(fr_t) == 4, "bb31_t");
#endif
static_assert(sizeof(RustError) == 16 && sizeof(RustError::by_value) == 16, "RustError by value");

// the polynomial/ templates keep the reference's names and argument order
// (polynomial/prefix_op.cuh:322, div_by_x_minus_z.cuh:445, evaluate.cuh:308): instantiate each once
RustError (*const poly_add)(fr_t*, const fr_t*, size_t, stream_t&) = &prefix_op<Add<fr_t>>;
RustError (*const poly_mul)(fr_t*, const fr_t*, size_t, stream_t&) = &prefix_op<Multiply<fr_t>>;
RustError (*const poly_div)(fr_t*, size_t, const fr_t&, stream_t&) = &div_by_x_minus_z<false, fr_t>;
RustError (*const poly_rot)(fr_t*, size_t, const fr_t&, stream_t&) = &div_by_x_minus_z<true, fr_t>;
RustError (*const poly_eval)(fr_t*, const fr_t*, size_t, const fr_t*, size_t, stream_t&) = &evaluate<fr_t>;
RustError (*const poly_inv)(fr_t*, const fr_t*, size_t, stream_t&) = &batch_inversion<fr_t>;
int main() { return 0; }
