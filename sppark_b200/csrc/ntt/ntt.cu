// NTT instantiations and their C-ABI entry points (include/sppark_b200.h).
#include "../ff/gl64.cuh"
#include "../ff/bb31.cuh"
#include "../ff/mont_ntt.cuh"
#include "ntt.cuh"

namespace ntt {
// lg_tile: log2(elements) of one CTA's shared-memory tile: 128 KiB of data for either field
template<> struct FieldId<gl64> { static constexpr uint32_t id = 1, lg_tile = 14; };
template<> struct FieldId<bb31> { static constexpr uint32_t id = 2, lg_tile = 14; };
// 256-bit Montgomery scalar fields: 2^11-element tiles (64 KiB) + 64 KiB of sub-NTT twiddles
template<> struct FieldId<ff::bls12_381_fr_ntt> { static constexpr uint32_t id = 3, lg_tile = 11; };
template<> struct FieldId<ff::pallas_fr_ntt> { static constexpr uint32_t id = 4, lg_tile = 11; };
template<> struct FieldId<ff::vesta_fr_ntt> { static constexpr uint32_t id = 5, lg_tile = 11; };
template<> struct FieldId<ff::bn254_fr_ntt> { static constexpr uint32_t id = 6, lg_tile = 11; };
template<> struct FieldId<ff::bls12_377_fr_ntt> { static constexpr uint32_t id = 7, lg_tile = 11; };
// ---- statically shaped twins of the passes the planner emits for the common sizes ----------
// key = (lg_r, lg_w, in row-fast, out row-fast, in_rev, out_rev, tw_mode)
template<class F, uint32_t R, uint32_t W, bool IRF, bool ORF, bool IREV, bool OREV, uint32_t TW>
static bool try_static(const Pass& d, const Tables<F>& tb, const typename F::T* in, typename F::T* out,
                       uint32_t ntiles, size_t smem, cudaStream_t stream)
{
    if (d.lg_r != R || d.lg_w != W || (d.in_lg_sa == 0) != IRF || (d.out_lg_sa == 0) != ORF ||
        (d.in_rev != 0) != IREV || (d.out_rev != 0) != OREV || d.tw_mode != TW)
        return false;
    typedef KStat<R, W, IRF, ORF, IREV, OREV, TW> K;
    // function attributes are per device (context): one flag per CUDA device and instantiation
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    static bool attr_done[64];
    if (!attr_done[dev & 63]) {
        CUDA_OK(cudaFuncSetAttribute(pass_kernel_static<F, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done[dev & 63] = true;
    }
    // one CTA per SM (a tile fills the shared memory), each walking ntiles / grid tiles
    static int sms_of[64];
    if (!sms_of[dev & 63]) CUDA_OK(cudaDeviceGetAttribute(&sms_of[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    const int sms = sms_of[dev & 63];
    uint32_t per_sm = smem <= 48 * 1024 ? 4 : smem <= 100 * 1024 ? 2 : 1;
    uint32_t grid = ntiles < (uint32_t)sms * per_sm ? ntiles : (uint32_t)sms * per_sm;
    pass_kernel_static<F, K><<<grid, tile_threads<F>(d), smem, stream>>>(d, tb, in, out, ntiles);
    return true;
}

// the seven (in-order, out-order, twiddle) combinations used by NN / NR / RN plans
template<class F, uint32_t R, uint32_t W>
static bool try_shapes(const Pass& d, const Tables<F>& tb, const typename F::T* in, typename F::T* out,
                       uint32_t ntiles, size_t smem, cudaStream_t stream)
{
    return try_static<F, R, W, false, true, false, false, TW_STORE>(d, tb, in, out, ntiles, smem, stream)    // NN first
        || try_static<F, R, W, false, false, false, false, TW_NONE>(d, tb, in, out, ntiles, smem, stream)    // NN last
        || try_static<F, R, W, false, false, false, false, TW_STORE>(d, tb, in, out, ntiles, smem, stream)   // NN middle
        || try_static<F, R, W, false, false, false, true, TW_STORE>(d, tb, in, out, ntiles, smem, stream)    // NR strided
        || try_static<F, R, W, true, true, false, true, TW_NONE>(d, tb, in, out, ntiles, smem, stream)       // NR last
        || try_static<F, R, W, true, true, true, false, TW_NONE>(d, tb, in, out, ntiles, smem, stream)       // RN first
        || try_static<F, R, W, false, false, true, false, TW_LOAD>(d, tb, in, out, ntiles, smem, stream);    // RN strided
}

template<class F> bool launch_static(const Pass& d, const Tables<F>& tb, const typename F::T* in,
                                     typename F::T* out, uint32_t ntiles, size_t smem, cudaStream_t stream)
{
    if (getenv("SPPARK_B200_NTT_GENERIC")) return false;
    if constexpr (F::LG_EPT != 4) return false;       // 256-bit fields run the run-time shaped kernel
    else return try_shapes<F, 12, 2>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 11, 3>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 10, 4>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 12, 1>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 11, 2>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 10, 3>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 10, 2>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 11, 1>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 11, 0>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 10, 1>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 10, 0>(d, tb, in, out, ntiles, smem, stream)
        || try_shapes<F, 9, 5>(d, tb, in, out, ntiles, smem, stream);    // 2^27 = 9+9+9 (BabyBear's maximum)
}

template bool launch_static<gl64>(const Pass&, const Tables<gl64>&, const uint64_t*, uint64_t*, uint32_t, size_t, cudaStream_t);
template bool launch_static<bb31>(const Pass&, const Tables<bb31>&, const uint32_t*, uint32_t*, uint32_t, size_t, cudaStream_t);


template bool launch_static<ff::bls12_381_fr_ntt>(const Pass&, const Tables<ff::bls12_381_fr_ntt>&, const ff::bls12_381_fr_ntt::T*, ff::bls12_381_fr_ntt::T*, uint32_t, size_t, cudaStream_t);
template bool launch_static<ff::pallas_fr_ntt>(const Pass&, const Tables<ff::pallas_fr_ntt>&, const ff::pallas_fr_ntt::T*, ff::pallas_fr_ntt::T*, uint32_t, size_t, cudaStream_t);
template bool launch_static<ff::vesta_fr_ntt>(const Pass&, const Tables<ff::vesta_fr_ntt>&, const ff::vesta_fr_ntt::T*, ff::vesta_fr_ntt::T*, uint32_t, size_t, cudaStream_t);
template bool launch_static<ff::bn254_fr_ntt>(const Pass&, const Tables<ff::bn254_fr_ntt>&, const ff::bn254_fr_ntt::T*, ff::bn254_fr_ntt::T*, uint32_t, size_t, cudaStream_t);
template bool launch_static<ff::bls12_377_fr_ntt>(const Pass&, const Tables<ff::bls12_377_fr_ntt>&, const ff::bls12_377_fr_ntt::T*, ff::bls12_377_fr_ntt::T*, uint32_t, size_t, cudaStream_t);

template class NTT<gl64>;
template class NTT<bb31>;
template class NTT<ff::bls12_381_fr_ntt>;
template class NTT<ff::pallas_fr_ntt>;
template class NTT<ff::vesta_fr_ntt>;
template class NTT<ff::bn254_fr_ntt>;
template class NTT<ff::bls12_377_fr_ntt>;
}  // namespace ntt

template<class F>
static RustError ntt_host(size_t device_id, void* inout, uint32_t lg, int order, int direction, int type)
{
    typedef ntt::NTT<F> N;
    if (order < 0 || order > 4 || direction < 0 || direction > 1 || type < 0 || type > 1)
        return rust_err(-(int)cudaErrorInvalidValue, "compute_ntt: bad order/direction/type");
    try {
        const gpu_t& gpu = select_gpu((int)device_id);
        return N::Base(gpu, (typename F::T*)inout, lg, (typename N::InputOutputOrder)order,
                       (typename N::Direction)direction, (typename N::Type)type);
    } catch (const cuda_error& e) {
        return rust_err(e.code(), e.what());
    } catch (const std::exception& e) {
        return rust_err(-1, e.what());
    }
}

template<class F>
static RustError ntt_dev(void* d_inout, uint32_t lg, int order, int direction, int type, void* stream)
{
    typedef ntt::NTT<F> N;
    if (order < 0 || order > 4 || direction < 0 || direction > 1 || type < 0 || type > 1)
        return rust_err(-(int)cudaErrorInvalidValue, "ntt_dev: bad order/direction/type");
    try {
        const gpu_t& gpu = gpu_of_current_device();
        N::Base_dev_ptr(gpu, (cudaStream_t)stream, (typename F::T*)d_inout, lg,
                        (typename N::InputOutputOrder)order, (typename N::Direction)direction,
                        (typename N::Type)type);
        return rust_ok();
    } catch (const cuda_error& e) {
        return rust_err(e.code(), e.what());
    } catch (const std::exception& e) {
        return rust_err(-1, e.what());
    }
}

template<class F>
static RustError ntt_slab(int which, const void* d_in, void* d_out, uint32_t lg, uint32_t lg_g, uint32_t rank,
                          int direction, void* stream, void* const* peers = nullptr)
{
    typedef ntt::NTT<F> N;
    if (direction < 0 || direction > 1 || (which != 1 && which != 2))
        return rust_err(-(int)cudaErrorInvalidValue, "ntt_slab_pass: bad direction / pass");
    try {
        N::slab_pass(gpu_of_current_device(), which, (const typename F::T*)d_in, (typename F::T*)d_out, lg, lg_g,
                     rank, (typename N::Direction)direction, (cudaStream_t)stream, peers);
        return rust_ok();
    } catch (const cuda_error& e) {
        return rust_err(e.code(), e.what());
    } catch (const std::exception& e) {
        return rust_err(-1, e.what());
    }
}

template<class F>
static RustError lde_host(size_t device_id, void* inout, uint32_t lg, uint32_t lg_blowup, void* aux)
{
    try {
        return ntt::NTT<F>::LDE(select_gpu((int)device_id), (typename F::T*)inout, lg, lg_blowup, (typename F::T*)aux);
    } catch (const cuda_error& e) {
        return rust_err(e.code(), e.what());
    } catch (const std::exception& e) {
        return rust_err(-1, e.what());
    }
}

extern "C" RustError sppark_b200_lde(int field, size_t device_id, void* inout, uint32_t lg, uint32_t lg_blowup, void* aux_out)
{
    switch (field) {
    case SPPARK_FIELD_GL64: return lde_host<gl64>(device_id, inout, lg, lg_blowup, aux_out);
    case SPPARK_FIELD_BB31: return lde_host<bb31>(device_id, inout, lg, lg_blowup, aux_out);
    case SPPARK_FIELD_BLS12_381_FR: return lde_host<ff::bls12_381_fr_ntt>(device_id, inout, lg, lg_blowup, aux_out);
    case SPPARK_FIELD_PALLAS_FR: return lde_host<ff::pallas_fr_ntt>(device_id, inout, lg, lg_blowup, aux_out);
    case SPPARK_FIELD_VESTA_FR: return lde_host<ff::vesta_fr_ntt>(device_id, inout, lg, lg_blowup, aux_out);
    case SPPARK_FIELD_BN254_FR: return lde_host<ff::bn254_fr_ntt>(device_id, inout, lg, lg_blowup, aux_out);
    case SPPARK_FIELD_BLS12_377_FR: return lde_host<ff::bls12_377_fr_ntt>(device_id, inout, lg, lg_blowup, aux_out);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_lde: unknown field");
    }
}

extern "C" RustError sppark_b200_ntt_slab_pass(int field, int which, const void* d_in, void* d_out,
                                               uint32_t lg, uint32_t lg_g, uint32_t rank, int direction, void* stream)
{
    switch (field) {
    case SPPARK_FIELD_GL64: return ntt_slab<gl64>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    case SPPARK_FIELD_BB31: return ntt_slab<bb31>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    case SPPARK_FIELD_BLS12_381_FR: return ntt_slab<ff::bls12_381_fr_ntt>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    case SPPARK_FIELD_PALLAS_FR: return ntt_slab<ff::pallas_fr_ntt>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    case SPPARK_FIELD_VESTA_FR: return ntt_slab<ff::vesta_fr_ntt>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    case SPPARK_FIELD_BN254_FR: return ntt_slab<ff::bn254_fr_ntt>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    case SPPARK_FIELD_BLS12_377_FR: return ntt_slab<ff::bls12_377_fr_ntt>(which, d_in, d_out, lg, lg_g, rank, direction, stream);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_ntt_slab_pass: unknown field");
    }
}

extern "C" RustError sppark_b200_ntt_slab_pass_p2p(int field, const void* d_in, void* const* peer_recv,
                                                   uint32_t lg, uint32_t lg_g, uint32_t rank, int direction, void* stream)
{
    if (peer_recv == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "ntt_slab_pass_p2p: no peer buffers");
    switch (field) {
    case SPPARK_FIELD_GL64: return ntt_slab<gl64>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    case SPPARK_FIELD_BB31: return ntt_slab<bb31>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    case SPPARK_FIELD_BLS12_381_FR: return ntt_slab<ff::bls12_381_fr_ntt>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    case SPPARK_FIELD_PALLAS_FR: return ntt_slab<ff::pallas_fr_ntt>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    case SPPARK_FIELD_VESTA_FR: return ntt_slab<ff::vesta_fr_ntt>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    case SPPARK_FIELD_BN254_FR: return ntt_slab<ff::bn254_fr_ntt>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    case SPPARK_FIELD_BLS12_377_FR: return ntt_slab<ff::bls12_377_fr_ntt>(1, d_in, nullptr, lg, lg_g, rank, direction, stream, peer_recv);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_ntt_slab_pass_p2p: unknown field");
    }
}

extern "C" RustError compute_ntt(size_t device_id, void* inout, uint32_t lg_domain_size,
                                 int ntt_order, int ntt_direction, int ntt_type)
{   return ntt_host<gl64>(device_id, inout, lg_domain_size, ntt_order, ntt_direction, ntt_type);   }

extern "C" RustError sppark_b200_ntt(int field, size_t device_id, void* inout, uint32_t lg,
                                     int order, int direction, int type)
{
    switch (field) {
    case SPPARK_FIELD_GL64: return ntt_host<gl64>(device_id, inout, lg, order, direction, type);
    case SPPARK_FIELD_BB31: return ntt_host<bb31>(device_id, inout, lg, order, direction, type);
    case SPPARK_FIELD_BLS12_381_FR: return ntt_host<ff::bls12_381_fr_ntt>(device_id, inout, lg, order, direction, type);
    case SPPARK_FIELD_PALLAS_FR: return ntt_host<ff::pallas_fr_ntt>(device_id, inout, lg, order, direction, type);
    case SPPARK_FIELD_VESTA_FR: return ntt_host<ff::vesta_fr_ntt>(device_id, inout, lg, order, direction, type);
    case SPPARK_FIELD_BN254_FR: return ntt_host<ff::bn254_fr_ntt>(device_id, inout, lg, order, direction, type);
    case SPPARK_FIELD_BLS12_377_FR: return ntt_host<ff::bls12_377_fr_ntt>(device_id, inout, lg, order, direction, type);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_ntt: unknown field");
    }
}

extern "C" RustError sppark_b200_ntt_dev(int field, void* d_inout, uint32_t lg, int order,
                                         int direction, int type, void* stream)
{
    switch (field) {
    case SPPARK_FIELD_GL64: return ntt_dev<gl64>(d_inout, lg, order, direction, type, stream);
    case SPPARK_FIELD_BB31: return ntt_dev<bb31>(d_inout, lg, order, direction, type, stream);
    case SPPARK_FIELD_BLS12_381_FR: return ntt_dev<ff::bls12_381_fr_ntt>(d_inout, lg, order, direction, type, stream);
    case SPPARK_FIELD_PALLAS_FR: return ntt_dev<ff::pallas_fr_ntt>(d_inout, lg, order, direction, type, stream);
    case SPPARK_FIELD_VESTA_FR: return ntt_dev<ff::vesta_fr_ntt>(d_inout, lg, order, direction, type, stream);
    case SPPARK_FIELD_BN254_FR: return ntt_dev<ff::bn254_fr_ntt>(d_inout, lg, order, direction, type, stream);
    case SPPARK_FIELD_BLS12_377_FR: return ntt_dev<ff::bls12_377_fr_ntt>(d_inout, lg, order, direction, type, stream);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_ntt_dev: unknown field");
    }
}
