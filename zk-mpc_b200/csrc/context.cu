// Context, error reporting and memory helpers of libmpc_cuda.so (include/mpc_cuda.h, "context").
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
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
static thread_local cudaStream_t t_aux_streams[64] = {};     // second per-thread stream (producer side of pipelines)

// Builds the device list in a local vector and publishes it only when every device passed, so a failed
// attempt leaves nothing behind.  Once initialised, a request for a DIFFERENT explicit list is an error
// (the library keeps per-device caches keyed by list index); NULL/0 or the same list is a no-op.
static int32_t init_locked(const int32_t* devices, int32_t n_dev) {
    if (g_inited) {
        if (devices && n_dev > 0) {
            bool same = (size_t)n_dev == g_devices.size();
            for (int i = 0; same && i < n_dev; i++) same = devices[i] == g_devices[i].cuda_device;
            if (!same) {
                set_error("mpc_cuda_init: already initialised with a different device list (%zu devices, first = CUDA %d)",
                          g_devices.size(), g_devices.empty() ? -1 : g_devices[0].cuda_device);
                return MPC_CUDA_ERR_ARG;
            }
        }
        return MPC_CUDA_OK;
    }
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
            for (int j = 0; j < i; j++) {
                if (devices[j] == devices[i]) {
                    set_error("device %d listed twice", devices[i]);
                    return MPC_CUDA_ERR_ARG;
                }
            }
            ids.push_back(devices[i]);
        }
    } else {
        for (int i = 0; i < count; i++) ids.push_back(i);
    }
    if (ids.size() > 64) ids.resize(64);
    std::vector<DeviceInfo> found;
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
        found.push_back(d);
        // keep freed scratch in the pool instead of returning it to the driver on every sync
        cudaMemPool_t pool;
        MPC_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, id));
        uint64_t threshold = UINT64_MAX;
        MPC_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    }
    g_devices.swap(found);
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

int device_list_size() { return g_inited ? (int)g_devices.size() : 0; }

int32_t aux_stream(cudaStream_t* out) {
    MPC_TRY(enter(nullptr));
    if (!t_aux_streams[t_dev_index])
        MPC_CUDA_TRY(cudaStreamCreateWithFlags(&t_aux_streams[t_dev_index], cudaStreamNonBlocking));
    *out = t_aux_streams[t_dev_index];
    return MPC_CUDA_OK;
}

DeviceScope::DeviceScope(int index) : saved(t_dev_index), rc(MPC_CUDA_OK) {
    if (index < 0 || index >= (int)g_devices.size()) {
        set_error("device index %d outside the init list of %zu devices", index, g_devices.size());
        rc = MPC_CUDA_ERR_ARG;
        return;
    }
    t_dev_index = index;
    rc = enter(&s);
}

DeviceScope::~DeviceScope() {
    t_dev_index = saved;
    if (saved >= 0 && saved < (int)g_devices.size()) cudaSetDevice(g_devices[saved].cuda_device);
}

static bool g_peers_enabled = false;
int32_t enable_peer_access() {
    std::lock_guard<std::mutex> lk(g_mu);
    MPC_TRY(init_locked(nullptr, 0));
    if (g_peers_enabled) return MPC_CUDA_OK;
    int cur = 0;
    MPC_CUDA_TRY(cudaGetDevice(&cur));
    for (size_t a = 0; a < g_devices.size(); a++) {
        MPC_CUDA_TRY(cudaSetDevice(g_devices[a].cuda_device));
        for (size_t b = 0; b < g_devices.size(); b++) {
            if (a == b) continue;
            int can = 0;
            MPC_CUDA_TRY(cudaDeviceCanAccessPeer(&can, g_devices[a].cuda_device, g_devices[b].cuda_device));
            if (!can) {
                cudaSetDevice(cur);
                set_error("CUDA devices %d and %d have no peer access", g_devices[a].cuda_device, g_devices[b].cuda_device);
                return MPC_CUDA_ERR_CUDA;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(g_devices[b].cuda_device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) {
                cudaSetDevice(cur);
                set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", g_devices[a].cuda_device, g_devices[b].cuda_device,
                          cudaGetErrorString(e));
                return MPC_CUDA_ERR_CUDA;
            }
        }
    }
    // stream-ordered scratch (cudaMallocAsync) is invisible to peers by default: open every default pool to
    // every other device of the list, so peer copies and peer loads/stores work on scratch too
    for (size_t a = 0; a < g_devices.size(); a++) {
        cudaMemPool_t pool;
        MPC_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, g_devices[a].cuda_device));
        std::vector<cudaMemAccessDesc> desc;
        for (size_t b = 0; b < g_devices.size(); b++) {
            if (a == b) continue;
            cudaMemAccessDesc d = {};
            d.location.type = cudaMemLocationTypeDevice;
            d.location.id = g_devices[b].cuda_device;
            d.flags = cudaMemAccessFlagsProtReadWrite;
            desc.push_back(d);
        }
        if (!desc.empty()) MPC_CUDA_TRY(cudaMemPoolSetAccess(pool, desc.data(), desc.size()));
    }
    MPC_CUDA_TRY(cudaSetDevice(cur));
    g_peers_enabled = true;
    return MPC_CUDA_OK;
}

const DeviceInfo* current_device_info() {
    if (t_dev_index < 0 || t_dev_index >= (int)g_devices.size()) return nullptr;
    return &g_devices[t_dev_index];
}

int current_device_index() { return t_dev_index; }

std::atomic<int64_t> g_opt_profile{0};
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// pending (start, stop) event pairs of the calling thread, folded into per-name totals on read
struct Pending {
    std::string name;
    cudaEvent_t e0, e1;
    bool closed;
};
static thread_local std::vector<Pending> t_pending;
static thread_local std::map<std::string, std::pair<double, uint64_t>> t_totals;

void profile_begin(const char* name, cudaStream_t s) {
    if (!g_opt_profile.load(std::memory_order_relaxed)) return;
    Pending p;
    p.name = name;
    p.closed = false;
    if (cudaEventCreate(&p.e0) != cudaSuccess || cudaEventCreate(&p.e1) != cudaSuccess) return;
    cudaEventRecord(p.e0, s);
    t_pending.push_back(p);
}

void profile_end(const char* name, cudaStream_t s) {
    if (!g_opt_profile.load(std::memory_order_relaxed)) return;
    for (size_t i = t_pending.size(); i-- > 0;) {
        if (!t_pending[i].closed && t_pending[i].name == name) {
            cudaEventRecord(t_pending[i].e1, s);
            t_pending[i].closed = true;
            return;
        }
    }
}

static void profile_fold() {
    for (Pending& p : t_pending) {
        if (p.closed && cudaEventSynchronize(p.e1) == cudaSuccess) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
                t_totals[p.name].first += ms;
                t_totals[p.name].second += 1;
            }
        }
        cudaEventDestroy(p.e0);
        cudaEventDestroy(p.e1);
    }
    t_pending.clear();
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

int32_t mpc_cuda_set_option(const char* name, int64_t value) {
    MPC_ARG_CHECK(name != nullptr);
    if (!strcmp(name, "msm_window_bits")) {
        MPC_ARG_CHECK(value >= 0 && value <= 23);
        g_opt_msm_window_bits = value;
    } else if (!strcmp(name, "profile")) {
        g_opt_profile = value ? 1 : 0;
    } else if (!strcmp(name, "msm_task_len")) {
        MPC_ARG_CHECK(value >= 0 && value <= (1 << 20));
        g_opt_msm_task_len = value;
    } else if (!strcmp(name, "ntt_occupancy")) {
        MPC_ARG_CHECK(value >= 0 && value <= 2);
        g_opt_ntt_occupancy = value;
    } else if (!strcmp(name, "ntt_graph")) {
        MPC_ARG_CHECK(value >= 0 && value <= 2);
        g_opt_ntt_graph = value;
    } else if (!strcmp(name, "ntt_generic")) {
        g_opt_ntt_generic = value ? 1 : 0;
    } else if (!strcmp(name, "msm_affine")) {
        MPC_ARG_CHECK(value >= 0 && value <= 3);
        g_opt_msm_affine = value;
    } else if (!strcmp(name, "l2_fetch_granularity")) {
        // hint to the driver (cudaLimitMaxL2FetchGranularity) for the calling thread's device: 32, 64 or 128 bytes
        MPC_ARG_CHECK(value == 32 || value == 64 || value == 128);
        MPC_TRY(enter(nullptr));
        MPC_CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    } else if (!strcmp(name, "msm_reduce_warp_max")) {
        MPC_ARG_CHECK(value >= 0 && value <= (1 << 20));
        g_opt_msm_reduce_warp_max = value;
    } else if (!strcmp(name, "msm_reduce_chunk")) {
        MPC_ARG_CHECK(value >= 0 && value <= 4096 && (value & (value - 1)) == 0);
        g_opt_msm_reduce_chunk = value;
    } else if (!strcmp(name, "msm_affine_split")) {
        MPC_ARG_CHECK(value >= 0 && value <= 2);
        g_opt_msm_affine_split = value;
    } else if (!strcmp(name, "msm_host_chunks")) {
        MPC_ARG_CHECK(value >= 0 && value <= 16);
        g_opt_msm_host_chunks = value;
    } else {
        set_error("unknown option '%s'", name);
        return MPC_CUDA_ERR_ARG;
    }
    return MPC_CUDA_OK;
}

uint64_t mpc_cuda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int32_t mpc_cuda_profile_read(const char* name, double* ms_total, uint64_t* count) {
    MPC_ARG_CHECK(name && ms_total && count);
    profile_fold();
    auto it = t_totals.find(name);
    *ms_total = it == t_totals.end() ? 0.0 : it->second.first;
    *count = it == t_totals.end() ? 0 : it->second.second;
    if (it != t_totals.end()) t_totals.erase(it);
    return MPC_CUDA_OK;
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

int32_t mpc_cuda_host_alloc(void** hptr, size_t bytes) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(hptr != nullptr && bytes > 0);
    MPC_CUDA_TRY(cudaHostAlloc(hptr, bytes, cudaHostAllocPortable));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_host_free(void* hptr) {
    MPC_TRY(enter(nullptr));
    MPC_CUDA_TRY(cudaFreeHost(hptr));
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

int32_t mpc_cuda_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_memcpy2d_d2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (width == 0 || height == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(dst && src && dpitch >= width && spitch >= width);
    MPC_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToDevice, pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_memset_zero_dev(void* dptr, size_t bytes, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_CUDA_TRY(cudaMemsetAsync(dptr, 0, bytes, pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_stream_create(void** stream) {
    MPC_ARG_CHECK(stream != nullptr);
    MPC_TRY(enter(nullptr));
    cudaStream_t s;
    MPC_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_stream_destroy(void* stream) {
    MPC_ARG_CHECK(stream != nullptr);
    MPC_TRY(enter(nullptr));
    MPC_CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_stream_sync(void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_CUDA_TRY(cudaStreamSynchronize(pick_stream(stream, s)));
    return MPC_CUDA_OK;
}

}  // extern "C"
