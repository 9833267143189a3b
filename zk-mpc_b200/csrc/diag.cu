// Diagnostics: raw field kernels for the parity tests and integer-pipe microbenchmarks that give
// bench.py its measured IMAD roofline denominator (include/mpc_cuda.h, "diagnostics").
#include "common.cuh"

using namespace mpc;

namespace {

template <class F>
__global__ void k_field_op(uint32_t op, const F* __restrict__ a, const F* __restrict__ b, F* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = load_fe(a + i), y = F::zero(), r;
    if (b) y = load_fe(b + i);
    switch (op) {
        case 0: r = add(x, y); break;
        case 1: r = sub(x, y); break;
        case 2: r = mul(x, y); break;
        case 3: r = neg(x); break;
        case 4: r = inv(x); break;
        case 5: r = from_mont(x); break;
        case 6: r = to_mont(x); break;
        case 8: r = mul_narrow(x, y); break;
        case 9: r = inv_euclid(x); break;
        default: r = sqr(x); break;
    }
    store_fe(out + i, r);
}

template <class F>
int32_t field_op(uint32_t op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(a && out);
    Scratch sa, sb, so;
    F *da, *db = nullptr, *dout;
    MPC_TRY(sa.alloc(&da, n, s));
    MPC_TRY(so.alloc(&dout, n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(da, a, n * sizeof(F), cudaMemcpyHostToDevice, s));
    if (b) {
        MPC_TRY(sb.alloc(&db, n, s));
        MPC_CUDA_TRY(cudaMemcpyAsync(db, b, n * sizeof(F), cudaMemcpyHostToDevice, s));
    }
    k_field_op<F><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(op, da, db, dout, n);
    MPC_KERNEL_CHECK();
    MPC_CUDA_TRY(cudaMemcpyAsync(out, dout, n * sizeof(F), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

// ---- microbenchmarks ------------------------------------------------------------------------------
constexpr int MB_THREADS = 256;
constexpr int MB_ILP = 8;

// kind 0: 32-bit IMAD, MB_ILP independent dependent-chains per thread
__global__ void __launch_bounds__(MB_THREADS) k_mb_imad(uint32_t iters, uint32_t a, uint32_t b, uint32_t* sink) {
    uint32_t x[MB_ILP];
#pragma unroll
    for (int k = 0; k < MB_ILP; k++) x[k] = threadIdx.x + k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < MB_ILP; k++) x[k] = x[k] * a + b;
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < MB_ILP; k++) acc ^= x[k];
    if (acc == 0x12345678u) sink[0] = acc;
}

// kind 1: IMAD.WIDE.U32 carry chains shaped like one Montgomery row (6 columns), 2 independent chains
__global__ void __launch_bounds__(MB_THREADS) k_mb_wide(uint32_t iters, uint32_t a, uint32_t b, uint32_t* sink) {
    uint64_t e[6], o[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { e[k] = threadIdx.x + k; o[k] = threadIdx.x * 3 + k; }
    for (uint32_t it = 0; it < iters; it++) {
        e[0] = ptx::madw_cc(a, b, e[0]);
#pragma unroll
        for (int k = 1; k < 6; k++) e[k] = ptx::madwc_cc(a + k, b, e[k]);
        o[0] = ptx::madw_cc(b, a, o[0]);
#pragma unroll
        for (int k = 1; k < 6; k++) o[k] = ptx::madwc_cc(b + k, a, o[k]);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) acc ^= e[k] ^ o[k];
    if (acc == 0x12345678u) sink[0] = (uint32_t)acc;
}

// kind 6: FP64 fused multiply-add, MB_ILP independent chains per thread (is the DFMA pipe worth a 52-bit-limb
// Montgomery product next to the IMAD.WIDE one?)
__global__ void __launch_bounds__(MB_THREADS) k_mb_dfma(uint32_t iters, double a, double b, double* sink) {
    double x[MB_ILP];
#pragma unroll
    for (int k = 0; k < MB_ILP; k++) x[k] = (double)(threadIdx.x + k);
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < MB_ILP; k++) x[k] = __fma_rz(x[k], a, b);
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < MB_ILP; k++) acc += x[k];
    if (acc == 0.12345) sink[0] = acc;
}

// kinds 2..5: Montgomery products, one dependent chain per thread
template <class F, bool NARROW>
__global__ void __launch_bounds__(MB_THREADS) k_mb_mul(uint32_t iters, const F* __restrict__ in, F* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F x = load_fe(in + (i & 1023)), y = load_fe(in + ((i + 1) & 1023));
    for (uint32_t it = 0; it < iters; it++) x = NARROW ? mul_narrow(x, y) : mul_wide(x, y);
    if (x.v[0] == 0x12345678u && x.v[1] == 0x9abcdef0u) store_fe(out, x);
}

}  // namespace

extern "C" {

int32_t mpc_cuda_field_op(uint32_t field, uint32_t op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    MPC_ARG_CHECK(field <= 1 && op <= 9);
    return field == 0 ? field_op<Fr>(op, a, b, out, n) : field_op<Fq>(op, a, b, out, n);
}

int32_t mpc_cuda_microbench(uint32_t kind, uint32_t iters, double* gops) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(kind <= 6 && iters >= 1 && gops);
    const DeviceInfo* d = current_device_info();
    int blocks = d->sm_count * 8;
    Scratch sin, sout;
    uint32_t* buf;
    uint32_t* sink;
    MPC_TRY(sin.alloc(&buf, 1024 * 12 + 64, s));
    MPC_TRY(sout.alloc(&sink, 64, s));
    // operands: arbitrary reduced-looking values (top limb small) so products stay in range
    {
        static uint32_t host[1024 * 12];
        uint32_t z = 0x9E3779B9u;
        for (int i = 0; i < 1024 * 12; i++) { z = z * 1664525u + 1013904223u; host[i] = z; }
        for (int i = 0; i < 1024; i++) { host[i * 12 + 11] &= 0x00ffffffu; host[i * 8 + 7] &= 0x0fffffffu; }
        MPC_CUDA_TRY(cudaMemcpyAsync(buf, host, sizeof(host), cudaMemcpyHostToDevice, s));
    }
    cudaEvent_t e0, e1;
    MPC_CUDA_TRY(cudaEventCreate(&e0));
    MPC_CUDA_TRY(cudaEventCreate(&e1));
    double per_thread = 0;
    for (int rep = 0; rep < 2; rep++) {      // rep 0 warms up
        MPC_CUDA_TRY(cudaEventRecord(e0, s));
        switch (kind) {
            case 0: k_mb_imad<<<blocks, MB_THREADS, 0, s>>>(iters, 0x10dcdu, 12345u, sink); per_thread = (double)iters * MB_ILP; break;
            case 1: k_mb_wide<<<blocks, MB_THREADS, 0, s>>>(iters, 0x9e3779b1u, 0x85ebca6bu, sink); per_thread = (double)iters * 12; break;
            case 2: k_mb_mul<Fq, false><<<blocks, MB_THREADS, 0, s>>>(iters, (const Fq*)buf, (Fq*)sink); per_thread = iters; break;
            case 3: k_mb_mul<Fr, false><<<blocks, MB_THREADS, 0, s>>>(iters, (const Fr*)buf, (Fr*)sink); per_thread = iters; break;
            case 6: k_mb_dfma<<<blocks, MB_THREADS, 0, s>>>(iters, 1.0000001, 0.5, (double*)sink); per_thread = (double)iters * MB_ILP; break;
            case 4: k_mb_mul<Fq, true><<<blocks, MB_THREADS, 0, s>>>(iters, (const Fq*)buf, (Fq*)sink); per_thread = iters; break;
            default: k_mb_mul<Fr, true><<<blocks, MB_THREADS, 0, s>>>(iters, (const Fr*)buf, (Fr*)sink); per_thread = iters; break;
        }
        MPC_KERNEL_CHECK();
        MPC_CUDA_TRY(cudaEventRecord(e1, s));
        MPC_CUDA_TRY(cudaEventSynchronize(e1));
    }
    float ms = 0;
    MPC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gops = per_thread * (double)blocks * MB_THREADS / (ms * 1e-3) / 1e9;
    return MPC_CUDA_OK;
}

}  // extern "C"
