// Context, error reporting and memory helpers of libmpc_cuda.so (include/mpc_cuda.h, "context").
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpc {

static thread_local char t_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// global: the device list chosen at init (guarded); per thread: selected device index, party, stream
static std::mutex g_mu;
static std::vector<DeviceInfo> g_devices;
static bool g_inited = false;

static thread_local int t_dev_index = -1;       // index into g_devices
static thread_local uint32_t t_party = 0, t_parties = 1;
static thread_local cudaStream_t t_streams[64] = {};

static int32_t init_locked(const int32_t* devices, int32_t n_dev) {
    if (g_inited) return MPC_CUDA_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libmpc_cuda has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return MPC_CUDA_ERR_NO_DEVICE;
    }
    std::vector<int> ids;
    if (devices && n_dev > 0) {
        for (int i = 0; i < n_dev; i++) {
            if (devices[i] < 0 || devices[i] >= count) {
                set_error("device %d out of range (0..%d)", devices[i], count - 1);
                return MPC_CUDA_ERR_ARG;
            }
            ids.push_back(devices[i]);
        }
    } else {
        for (int i = 0; i < count; i++) ids.push_back(i);
    }
    if (ids.size() > 64) ids.resize(64);
    for (int id : ids) {
        cudaDeviceProp prop;
        MPC_CUDA_TRY(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10) {
            set_error("device %d is sm_%d%d; libmpc_cuda is built for sm_100a only", id, prop.major, prop.minor);
            return MPC_CUDA_ERR_NO_DEVICE;
        }
        DeviceInfo d;
        d.cuda_device = id;
        d.sm_count = prop.multiProcessorCount;
        g_devices.push_back(d);
        // keep freed scratch in the pool instead of returning it to the driver on every sync
        cudaMemPool_t pool;
        MPC_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, id));
        uint64_t threshold = UINT64_MAX;
        MPC_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    }
    g_inited = true;
    return MPC_CUDA_OK;
}

int32_t enter(cudaStream_t* stream_out) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        MPC_TRY(init_locked(nullptr, 0));
    }
    if (t_dev_index < 0) t_dev_index = (int)(t_party % g_devices.size());
    MPC_CUDA_TRY(cudaSetDevice(g_devices[t_dev_index].cuda_device));
    if (!t_streams[t_dev_index]) MPC_CUDA_TRY(cudaStreamCreateWithFlags(&t_streams[t_dev_index], cudaStreamNonBlocking));
    if (stream_out) *stream_out = t_streams[t_dev_index];
    return MPC_CUDA_OK;
}

const DeviceInfo* current_device_info() {
    if (t_dev_index < 0 || t_dev_index >= (int)g_devices.size()) return nullptr;
    return &g_devices[t_dev_index];
}

bool is_leader() { return t_party == 0; }

}  // namespace mpc

using namespace mpc;

extern "C" {

int32_t mpc_cuda_init(const int32_t* devices, int32_t n_dev) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(devices, n_dev);
}

int32_t mpc_cuda_set_party(uint32_t party_id, uint32_t n_parties) {
    MPC_ARG_CHECK(n_parties >= 1 && party_id < n_parties);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        MPC_TRY(init_locked(nullptr, 0));
    }
    t_party = party_id;
    t_parties = n_parties;
    t_dev_index = (int)(party_id % g_devices.size());
    return enter(nullptr);
}

int32_t mpc_cuda_set_device(int32_t dev_index) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        MPC_TRY(init_locked(nullptr, 0));
    }
    MPC_ARG_CHECK(dev_index >= 0 && dev_index < (int)g_devices.size());
    t_dev_index = dev_index;
    return enter(nullptr);
}

int32_t mpc_cuda_device_count(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (init_locked(nullptr, 0) != MPC_CUDA_OK) return 0;
    return (int32_t)g_devices.size();
}

const char* mpc_cuda_last_error(void) { return t_err; }
const char* mpc_cuda_version(void) { return "mpc_cuda 0.1 (sm_100a)"; }

int32_t mpc_cuda_malloc(void** dptr, size_t bytes) {
    MPC_ARG_CHECK(dptr != nullptr);
    MPC_TRY(enter(nullptr));
    MPC_CUDA_TRY(cudaMalloc(dptr, bytes ? bytes : 1));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_free(void* dptr) {
    MPC_TRY(enter(nullptr));
    MPC_CUDA_TRY(cudaFree(dptr));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_CUDA_TRY(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_stream_sync(void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_CUDA_TRY(cudaStreamSynchronize(pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

}  // extern "C"
