// A caller written against the reference's polynomial headers (same include names, same template
// names and argument order) compiles against include/compat with plain g++: tests/test_cpp_layer.py.
#include <ff/goldilocks.hpp>
#include <util/gpu_t.cuh>
#include <polynomial/prefix_op.cuh>
#include <polynomial/div_by_x_minus_z.cuh>
#include <polynomial/evaluate.cuh>
#include <ff/batch_inversion.hpp>
int f(fr_t* d, size_t n, stream_t& s, const fr_t& z) {
    RustError e = prefix_op<Add<fr_t>>(d, d, n, s);
    e = div_by_x_minus_z<true>(d, n, z, s);
    e = evaluate(d, d, 1, d, n, s);
    return e.code;
}
int main() { return 0; }
