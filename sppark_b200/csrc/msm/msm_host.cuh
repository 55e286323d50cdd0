// MSM instantiations and their C-ABI entry points (include/sppark_b200.h).
//   mult_pippenger       poc/msm-cuda/cuda/pippenger.cu:20-25
//   mult_pippenger_inf   poc/msm-cuda/cuda/pippenger_inf.cu:28-34
// Host-pointer calls follow the reference's contract (msm/pippenger.cuh:730-747): scratch is
// allocated per call from the stream-ordered pool, scalars are 256-bit little-endian integers
// NOT in Montgomery form (mont=false), the result is a Jacobian point in Montgomery form.
#pragma once
#include "../ff/fields.cuh"
#include "msm.cuh"

namespace {

template<class F>
RustError msm_host(void* out, const void* points, size_t npoints, const void* scalars,
                   size_t stride, bool has_flag)
{
    constexpr size_t PB = 2 * F::N * 4, JB = 3 * F::N * 4;
    try {
        const gpu_t& gpu = select_gpu(-1);
        gpu.select();
        if (npoints == 0) { memset(out, 0, JB); return rust_ok(); }
        if (stride < PB + (has_flag ? 1 : 0))
            return rust_err(-(int)cudaErrorInvalidValue, "msm: affine stride too small");
        const stream_t &s0 = gpu[0], &s1 = gpu[1];
        dev_ptr_t<uint32_t> d_points(npoints * (PB / 4), s0), d_out(JB / 4, s0);
        dev_ptr_t<uint32_t> d_scalars(npoints * 8, s1);
        cudaEvent_t ev;
        CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        s1.HtoD(d_scalars, scalars, npoints * 32);
        CUDA_OK(cudaEventRecord(ev, s1));
        if (stride == PB && !has_flag) {
            s0.HtoD(d_points, points, npoints * PB);
        } else {
            dev_ptr_t<uint8_t> d_raw(npoints * stride, s0);
            s0.HtoD(d_raw, points, npoints * stride);
            uint32_t blocks = (uint32_t)std::min<size_t>((npoints + 255) / 256, (size_t)gpu.sm_count() * 16);
            msm::pack_points_kernel<<<blocks, 256, 0, s0>>>(d_raw, stride, PB / 4, has_flag, d_points,
                                                           (uint32_t)npoints);
            COUNT_LAUNCH();
            CUDA_OK(cudaGetLastError());
        }
        CUDA_OK(cudaStreamWaitEvent(s0, ev, 0));
        msm::msm_t<F> m(gpu);
        m.invoke_dev(d_out, d_points, npoints, d_scalars, s0);
        s0.DtoH(out, d_out, JB);
        s0.sync();
        s1.sync();
        cudaEventDestroy(ev);
    } catch (const cuda_error& e) {
        memset(out, 0, JB);                      // out->inf(), as the reference does on failure
        return rust_err(e.code(), e.what());
    } catch (const std::exception& e) {
        memset(out, 0, JB);
        return rust_err(-1, e.what());
    }
    return rust_ok();
}

template<class F>
RustError msm_dev(void* out, const void* d_points, size_t npoints, const void* d_scalars, void* stream)
{
    constexpr size_t JB = 3 * F::N * 4;
    try {
        const gpu_t& gpu = gpu_of_current_device();
        cudaStream_t s = (cudaStream_t)stream;
        uint32_t* d_out;
        CUDA_OK(cudaMallocAsync((void**)&d_out, JB, s));
        msm::msm_t<F> m(gpu);
        m.invoke_dev(d_out, (const uint32_t*)d_points, npoints, (const uint32_t*)d_scalars, s);
        CUDA_OK(cudaMemcpyAsync(out, d_out, JB, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaFreeAsync(d_out, s));
        CUDA_OK(cudaStreamSynchronize(s));
    } catch (const cuda_error& e) {
        memset(out, 0, JB);
        return rust_err(e.code(), e.what());
    } catch (const std::exception& e) {
        memset(out, 0, JB);
        return rust_err(-1, e.what());
    }
    return rust_ok();
}

}  // namespace
