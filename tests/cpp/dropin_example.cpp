// A C++ caller written the way one is written against supranational/sppark's templates
// (cf. poc/msm-cuda/cuda/pippenger_inf.cu:20-34 and poc/ntt-cuda/cuda/ntt_api.cu:25-36), built
// (-DFEATURE_BLS12_381, or -DFEATURE_GOLDILOCKS -DNTT_ONLY for the Goldilocks transform)
// against include/sppark_b200.hpp + libsppark_b200.so.  tests/test_cpp_layer.py feeds it inputs
// through stdin-less binary files and checks the outputs against the oracle.
//   dropin_example msm <points.bin> <scalars.bin> <n> <mont 0|1> <out.bin>
//   dropin_example ctx <points.bin> <scalars.bin> <n> <mont 0|1> <out.bin>   (msm_t with preloaded points)
//   dropin_example ntt <data.bin> <lg> <order> <direction> <type>       (in place)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sppark_b200.hpp"

#ifndef NTT_ONLY
typedef jacobian_t<fp_t> point_t;
typedef xyzz_t<fp_t> bucket_t;
typedef bucket_t::affine_inf_t affine_t;
typedef fr_t scalar_t;
#endif

template<class T> static std::vector<T> slurp(const char* path, size_t n)
{
    std::vector<T> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}
template<class T> static void dump(const char* path, const T* p, size_t n)
{
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fclose(f);
}
static int report(RustError e)
{
    if (e.code != 0) {
        fprintf(stderr, "error %d: %s\n", e.code, e.message ? e.message : "");
        drop_error_message(e.message);
        return 1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (!cuda_available()) { fprintf(stderr, "no CUDA device\n"); return 3; }
#ifndef NTT_ONLY
    if (argc == 7 && argv[1][0] == 'm') {
        size_t n = strtoull(argv[4], nullptr, 10);
        auto points = slurp<affine_t>(argv[2], n);
        auto scalars = slurp<scalar_t>(argv[3], n);
        point_t out;
        RustError e = mult_pippenger<bucket_t>(&out, points.data(), n, scalars.data(), atoi(argv[5]) != 0,
                                               sizeof(affine_t));
        if (report(e)) return 1;
        dump(argv[6], &out, 1);
        return 0;
    }
    if (argc == 7 && argv[1][0] == 'c') {                  // preloaded points, two invocations
        size_t n = strtoull(argv[4], nullptr, 10);
        auto points = slurp<affine_t>(argv[2], n);
        auto scalars = slurp<scalar_t>(argv[3], n);
        msm_t<bucket_t, point_t, affine_t, scalar_t> msm{points.data(), n, sizeof(affine_t)};
        point_t out[2];
        if (report(msm.invoke(out[0], scalars.data(), atoi(argv[5]) != 0))) return 1;
        if (report(msm.invoke(out[1], n / 2, scalars.data(), atoi(argv[5]) != 0))) return 1;   // a prefix
        dump(argv[6], out, 2);
        return 0;
    }
#endif
    if (argc == 7 && argv[1][0] == 'n') {
        uint32_t lg = (uint32_t)atoi(argv[3]);
        auto data = slurp<fr_t>(argv[2], (size_t)1 << lg);
        RustError e = NTT::Base(select_gpu(0), data.data(), lg, (NTT::InputOutputOrder)atoi(argv[4]),
                                (NTT::Direction)atoi(argv[5]), (NTT::Type)atoi(argv[6]));
        if (report(e)) return 1;
        dump(argv[2], data.data(), data.size());
        return 0;
    }
    fprintf(stderr, "usage: see the header comment\n");
    return 2;
}
