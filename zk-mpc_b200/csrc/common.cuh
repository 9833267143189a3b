// Library context shared by the translation units of libmpc_cuda.so: error reporting, per-thread
// device / party state, stream-ordered scratch memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdio.h>
#include <string.h>

#include "../../include/mpc_cuda.h"
#include "consts.cuh"
#include "fp.cuh"

using Fr = Fp<consts::FrParams>;
using Fq = Fp<consts::FqParams>;

namespace mpc {

// ---- errors -------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define MPC_CUDA_TRY(expr)                                                                       \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            mpc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MPC_CUDA_ERR_CUDA;                                                            \
        }                                                                                        \
    } while (0)

// after every kernel launch: surface launch errors and count the launch (mpc_cuda_launch_count)
#define MPC_KERNEL_CHECK()                          \
    do {                                            \
        MPC_CUDA_TRY(cudaGetLastError());           \
        mpc::count_launch();                        \
    } while (0)

#define MPC_TRY(expr)                  \
    do {                               \
        int32_t _rc = (expr);          \
        if (_rc != MPC_CUDA_OK) return _rc; \
    } while (0)

#define MPC_ARG_CHECK(cond)                                                     \
    do {                                                                        \
        if (!(cond)) {                                                          \
            mpc::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond); \
            return MPC_CUDA_ERR_ARG;                                            \
        }                                                                       \
    } while (0)

// ---- context ------------------------------------------------------------------------------------
struct DeviceInfo {
    int cuda_device;
    int sm_count;
};

// Makes the calling thread's mpc_cuda device current (lazy init on first use) and returns its
// per-thread stream.  Every exported entry point starts with this.
int32_t enter(cudaStream_t* stream_out);
const DeviceInfo* current_device_info();
int current_device_index();      // index into the init list (per-device caches are keyed by it)
int device_list_size();          // devices of the init list (0 before init)
int32_t aux_stream(cudaStream_t* out);   // the calling thread's second stream on its current device
bool is_leader();

// Multi-GPU entry points drive several devices from one host thread: a DeviceScope makes device `index`
// of the init list the calling thread's current mpc_cuda device (and CUDA device) until it goes out of
// scope, then restores the previous one.  stream() is the thread's own stream on that device.
struct DeviceScope {
    int saved;
    int32_t rc;
    cudaStream_t s = nullptr;
    explicit DeviceScope(int index);
    ~DeviceScope();
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
// cudaDeviceEnablePeerAccess between every ordered pair of the init list (idempotent, thread-safe);
// fails with MPC_CUDA_ERR_CUDA when a pair cannot reach each other
int32_t enable_peer_access();

void count_launch();

// stage timing with CUDA events on the launching stream, active while option "profile" is 1; read and
// reset with mpc_cuda_profile_read (bench.py uses it for the per-kernel roofline numbers)
void profile_begin(const char* name, cudaStream_t s);
void profile_end(const char* name, cudaStream_t s);
struct ProfileScope {
    const char* name;
    cudaStream_t s;
    ProfileScope(const char* n, cudaStream_t st) : name(n), s(st) { profile_begin(name, s); }
    ~ProfileScope() { profile_end(name, s); }
};

// tuning knobs set through mpc_cuda_set_option (0 = automatic)
extern std::atomic<int64_t> g_opt_msm_window_bits;
extern std::atomic<int64_t> g_opt_msm_task_len;
extern std::atomic<int64_t> g_opt_msm_host_chunks;
extern std::atomic<int64_t> g_opt_msm_affine;
extern std::atomic<int64_t> g_opt_msm_affine_split;
extern std::atomic<int64_t> g_opt_msm_reduce_chunk;
extern std::atomic<int64_t> g_opt_msm_reduce_warp_max;
extern std::atomic<int64_t> g_opt_profile;
extern std::atomic<int64_t> g_opt_ntt_generic;
extern std::atomic<int64_t> g_opt_ntt_occupancy;
extern std::atomic<int64_t> g_opt_ntt_graph;

inline cudaStream_t pick_stream(void* user, cudaStream_t mine) { return user ? (cudaStream_t)user : mine; }

// stream-ordered scratch (cudaMallocAsync pool of the current device)
template <class T>
inline int32_t scratch_alloc(T** p, size_t count, cudaStream_t s) {
    *p = nullptr;
    if (count == 0) count = 1;
    MPC_CUDA_TRY(cudaMallocAsync((void**)p, count * sizeof(T), s));
    return MPC_CUDA_OK;
}
inline void scratch_free(void* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

// RAII holder so early returns free scratch
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    Scratch() {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { scratch_free(p, s); }
    template <class T>
    int32_t alloc(T** out, size_t count, cudaStream_t stream) {
        s = stream;
        int32_t rc = scratch_alloc((T**)&p, count, stream);
        *out = (T*)p;
        return rc;
    }
};

// grid sizing: a multiple of the SM count, capped by the work available
inline int grid_for(size_t work_items, int threads, int ctas_per_sm) {
    const DeviceInfo* d = current_device_info();
    int sms = d ? d->sm_count : 148;
    size_t need = (work_items + threads - 1) / threads;
    size_t cap = (size_t)sms * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- 128-bit vectorised element I/O -----------------------------------------------------------------
// Field elements live in HBM as arrays of N 32-bit limbs (identical bytes to the reference's u64 limbs).
template <class F>
DEV F load_fe(const F* p) {
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) {
        uint4 t = q[i];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class F>
DEV F load_fe_ro(const F* p) {       // read-only path (LDG.E.128.CONSTANT)
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) {
        uint4 t = __ldg(q + i);
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class F>
DEV void store_fe(F* p, const F& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

}  // namespace mpc
