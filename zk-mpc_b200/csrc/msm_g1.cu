// C ABI of the share MSM over G1 (include/mpc_cuda.h, "share MSM"); the pipeline is msm_impl.cuh with F = Fq.
#include "msm_impl.cuh"

namespace {

template <class F>
int32_t msm_handle_host(uint64_t handle, size_t offset, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                        uint8_t* out_inf, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_xy && out_inf && (n == 0 || scalars));
    BaseSnap v;
    MPC_TRY(registry_find(handle, g2, offset, n, &v));
    if (!v.ref->parts.empty()) return msm_sharded<F>(v, offset, scalars, nullptr, n, out_xy, out_inf);
    Scratch ss;
    Fr* dsc;
    MPC_TRY(ss.alloc(&dsc, n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dsc, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    TableRef tbl = table_of(v, offset);
    return msm_emit<F>((const Affine<F>*)v.bases + offset, v.inf ? v.inf + offset : nullptr, dsc, n, 0, nullptr, out_xy,
                       out_inf, s, &tbl);
}

template <class F>
int32_t generate(const uint32_t* gx, const uint32_t* gy, uint64_t seed, size_t first, size_t n, uint64_t* out,
                 void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(out);
    Affine<F> g;
    memcpy(&g.x, gx, sizeof(F));
    memcpy(&g.y, gy, sizeof(F));
    k_generate<F><<<(unsigned)((n + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, pick_stream(stream, s)>>>(
        g, seed, first, n, (Affine<F>*)out);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

template <class F>
int32_t register_dev(const uint64_t* bases_xy_dev, size_t n, uint64_t* handle, bool g2) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(handle && (n == 0 || bases_xy_dev));
    BaseRef v = std::make_shared<BaseVec>();
    v->bases = (void*)bases_xy_dev;
    v->n = n;
    v->g2 = g2;
    v->dev_index = current_device_index();
    v->cuda_device = current_device_info()->cuda_device;
    v->owned = false;
    *handle = registry_add(v);
    return MPC_CUDA_OK;
}

template <class F>
int32_t handle_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n, uint64_t* out_jac_dev,
                   void* stream, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_jac_dev && (n == 0 || scalars_mont_dev));
    BaseSnap v;
    MPC_TRY(registry_find(handle, g2, offset, n, &v));
    MPC_ARG_CHECK(v.ref->parts.empty());
    TableRef tbl = table_of(v, offset);
    return msm_emit<F>((const Affine<F>*)v.bases + offset, v.inf ? v.inf + offset : nullptr, (const Fr*)scalars_mont_dev, n,
                       1, (uint32_t*)out_jac_dev, nullptr, nullptr, pick_stream(stream, s), &tbl);
}

// resident scalars (e.g. h left on the device by the fused witness map), affine result on the host
template <class F>
int32_t handle_scalars_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n, uint64_t* out_xy,
                           uint8_t* out_inf, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_xy && out_inf && (n == 0 || scalars_mont_dev));
    BaseSnap v;
    MPC_TRY(registry_find(handle, g2, offset, n, &v));
    MPC_ARG_CHECK(v.ref->parts.empty());
    TableRef tbl = table_of(v, offset);
    return msm_emit<F>((const Affine<F>*)v.bases + offset, v.inf ? v.inf + offset : nullptr, (const Fr*)scalars_mont_dev, n,
                       0, nullptr, out_xy, out_inf, s, &tbl);
}

template <class F>
int32_t sum_partials(const uint64_t* jac_dev, uint32_t count, uint64_t* out_xy, uint8_t* out_inf, void* stream) {
    constexpr int N = sizeof(F) / 4;
    cudaStream_t s0;
    MPC_TRY(enter(&s0));
    cudaStream_t s = pick_stream(stream, s0);
    MPC_ARG_CHECK(out_xy && out_inf && (count == 0 || jac_dev));
    Scratch sx, so;
    XYZZ<F>* pts;
    uint32_t* out;
    MPC_TRY(sx.alloc(&pts, count, s));
    MPC_TRY(so.alloc(&out, 3 * N + 4, s));
    if (count) {
        k_jac_to_xyzz<F><<<(count + 127) / 128, 128, 0, s>>>((const Jac<F>*)jac_dev, count, pts);
        MPC_KERNEL_CHECK();
    }
    k_emit<F><<<1, 32, 0, s>>>(pts, count, 0, out);
    MPC_KERNEL_CHECK();
    uint32_t host[2 * N + 1];
    MPC_CUDA_TRY(cudaMemcpyAsync(host, out, sizeof(host), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_xy, host, 2 * N * sizeof(uint32_t));
    *out_inf = (uint8_t)host[2 * N];
    return MPC_CUDA_OK;
}

template <class F>
int32_t handle_sharded_dev(uint64_t handle, const uint64_t* const* scalars_dev, uint32_t parts, uint64_t* out_xy,
                           uint8_t* out_inf, bool g2) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(scalars_dev && out_xy && out_inf);
    BaseSnap v;
    MPC_TRY(registry_find(handle, g2, 0, 0, &v));
    MPC_ARG_CHECK(!v.ref->parts.empty() && v.ref->parts.size() == parts);
    return msm_sharded<F>(v, 0, nullptr, scalars_dev, v.n, out_xy, out_inf);
}

}  // namespace

#ifndef MSM_CURVE_G2

extern "C" {

int32_t mpc_cuda_msm_g1(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars_mont, size_t n,
                        uint64_t out_xy[12], uint8_t* out_inf) {
    return msm_host<Fq>(bases_xy, inf, scalars_mont, n, out_xy, out_inf);
}

int32_t mpc_cuda_msm_g1_register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle) {
    return register_bases<Fq>(bases_xy, inf, n, handle, false);
}

int32_t mpc_cuda_msm_g1_register_bases_sharded(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint32_t parts,
                                               uint64_t* handle) {
    return register_bases_sharded<Fq>(bases_xy, inf, n, parts, handle, false);
}

int32_t mpc_cuda_msm_g1_register_bases_dev(const uint64_t* bases_xy_dev, size_t n, uint64_t* handle) {
    return register_dev<Fq>(bases_xy_dev, n, handle, false);
}

int32_t mpc_cuda_msm_g1_precompute(uint64_t handle, uint32_t window_bits) {
    return precompute<Fq>(handle, window_bits, false);
}

int32_t mpc_cuda_msm_g1_handle(uint64_t handle, size_t offset, const uint64_t* scalars_mont, size_t n,
                               uint64_t out_xy[12], uint8_t* out_inf) {
    return msm_handle_host<Fq>(handle, offset, scalars_mont, n, out_xy, out_inf, false);
}

int32_t mpc_cuda_msm_g1_handle_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                   uint64_t* out_jac_dev, void* stream) {
    return handle_dev<Fq>(handle, offset, scalars_mont_dev, n, out_jac_dev, stream, false);
}

int32_t mpc_cuda_msm_g1_handle_scalars_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                           uint64_t out_xy[12], uint8_t* out_inf) {
    return handle_scalars_dev<Fq>(handle, offset, scalars_mont_dev, n, out_xy, out_inf, false);
}

int32_t mpc_cuda_msm_g1_handle_sharded_dev(uint64_t handle, const uint64_t* const* scalars_mont_dev, uint32_t parts,
                                           uint64_t out_xy[12], uint8_t* out_inf) {
    return handle_sharded_dev<Fq>(handle, scalars_mont_dev, parts, out_xy, out_inf, false);
}

int32_t mpc_cuda_g1_sum_partials_dev(const uint64_t* jac_dev, uint32_t count, uint64_t out_xy[12], uint8_t* out_inf,
                                     void* stream) {
    return sum_partials<Fq>(jac_dev, count, out_xy, out_inf, stream);
}

int32_t mpc_cuda_g1_generate_dev(uint64_t seed, size_t first, size_t n, uint64_t* out_xy_dev, void* stream) {
    return generate<Fq>(consts::G1_GEN_X, consts::G1_GEN_Y, seed, first, n, out_xy_dev, stream);
}

}  // extern "C"

#else  // MSM_CURVE_G2: the same entry points over Fq2 (compiled from msm_g2.cu)

extern "C" {

int32_t mpc_cuda_msm_g2(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars_mont, size_t n,
                        uint64_t out_xy[24], uint8_t* out_inf) {
    return msm_host<Fq2>(bases_xy, inf, scalars_mont, n, out_xy, out_inf);
}

int32_t mpc_cuda_msm_g2_register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle) {
    return register_bases<Fq2>(bases_xy, inf, n, handle, true);
}

int32_t mpc_cuda_msm_g2_register_bases_sharded(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint32_t parts,
                                               uint64_t* handle) {
    return register_bases_sharded<Fq2>(bases_xy, inf, n, parts, handle, true);
}

int32_t mpc_cuda_msm_g2_register_bases_dev(const uint64_t* bases_xy_dev, size_t n, uint64_t* handle) {
    return register_dev<Fq2>(bases_xy_dev, n, handle, true);
}

int32_t mpc_cuda_msm_g2_precompute(uint64_t handle, uint32_t window_bits) {
    return precompute<Fq2>(handle, window_bits, true);
}

int32_t mpc_cuda_msm_g2_handle(uint64_t handle, size_t offset, const uint64_t* scalars_mont, size_t n,
                               uint64_t out_xy[24], uint8_t* out_inf) {
    return msm_handle_host<Fq2>(handle, offset, scalars_mont, n, out_xy, out_inf, true);
}

int32_t mpc_cuda_msm_g2_handle_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                   uint64_t* out_jac_dev, void* stream) {
    return handle_dev<Fq2>(handle, offset, scalars_mont_dev, n, out_jac_dev, stream, true);
}

int32_t mpc_cuda_msm_g2_handle_scalars_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                           uint64_t out_xy[24], uint8_t* out_inf) {
    return handle_scalars_dev<Fq2>(handle, offset, scalars_mont_dev, n, out_xy, out_inf, true);
}

int32_t mpc_cuda_g2_sum_partials_dev(const uint64_t* jac_dev, uint32_t count, uint64_t out_xy[24], uint8_t* out_inf,
                                     void* stream) {
    return sum_partials<Fq2>(jac_dev, count, out_xy, out_inf, stream);
}

int32_t mpc_cuda_g2_generate_dev(uint64_t seed, size_t first, size_t n, uint64_t* out_xy_dev, void* stream) {
    return generate<Fq2>(consts::G2_GEN_X, consts::G2_GEN_Y, seed, first, n, out_xy_dev, stream);
}

}  // extern "C"

#endif
