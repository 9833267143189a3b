// Base-vector registry and the process-wide MSM tuning knobs (include/mpc_cuda.h, "share MSM").
#include <unordered_map>

#include "msm_registry.cuh"

namespace mpc {

std::atomic<int64_t> g_opt_msm_window_bits{0};
std::atomic<int64_t> g_opt_msm_task_len{0};
std::atomic<int64_t> g_opt_msm_host_chunks{0};
std::atomic<int64_t> g_opt_msm_affine{0};
std::atomic<int64_t> g_opt_msm_affine_split{0};
std::atomic<int64_t> g_opt_msm_reduce_chunk{0};
std::atomic<int64_t> g_opt_msm_reduce_warp_max{0};

static std::mutex g_bases_mu;
static std::unordered_map<uint64_t, BaseRef> g_bases;
static uint64_t g_next_handle = 1;

BaseVec::~BaseVec() {
    if (parts.empty() && (owned || table || !retired.empty())) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(cuda_device);
        if (owned) {
            if (bases) cudaFree(bases);
            if (inf) cudaFree(inf);
        }
        if (table) cudaFree(table);
        for (void* p : retired) cudaFree(p);
        cudaSetDevice(cur);
    }
}

uint64_t registry_add(const BaseRef& v) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    uint64_t h = g_next_handle++;
    g_bases[h] = v;
    return h;
}

BaseSnap snapshot_of(const BaseRef& v) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    BaseSnap s;
    s.ref = v;
    s.bases = v->bases;
    s.inf = v->inf;
    s.table = v->table;
    s.table_c = v->table_c;
    s.n = v->n;
    return s;
}

int32_t registry_find(uint64_t handle, bool g2, size_t offset, size_t n, BaseSnap* out) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    auto it = g_bases.find(handle);
    if (it == g_bases.end() || it->second->g2 != g2) {
        set_error("unknown %s base handle %llu", g2 ? "G2" : "G1", (unsigned long long)handle);
        return MPC_CUDA_ERR_HANDLE;
    }
    const BaseRef& v = it->second;
    if (offset > v->n || n > v->n - offset) {
        set_error("base range [%zu, %zu) outside registered vector of %zu points", offset, offset + n, v->n);
        return MPC_CUDA_ERR_ARG;
    }
    if (v->parts.empty() && v->cuda_device != current_device_info()->cuda_device) {
        set_error("base handle %llu lives on CUDA device %d, calling thread uses %d", (unsigned long long)handle,
                  v->cuda_device, current_device_info()->cuda_device);
        return MPC_CUDA_ERR_HANDLE;
    }
    out->ref = v;
    out->bases = v->bases;
    out->inf = v->inf;
    out->table = v->table;
    out->table_c = v->table_c;
    out->n = v->n;
    return MPC_CUDA_OK;
}

int32_t registry_set_table(uint64_t handle, void* table, uint32_t c) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    auto it = g_bases.find(handle);
    if (it == g_bases.end()) {
        set_error("base handle %llu released during precomputation", (unsigned long long)handle);
        return MPC_CUDA_ERR_HANDLE;
    }
    BaseVec& v = *it->second;
    if (v.table) v.retired.push_back(v.table);      // a concurrent MSM may still be about to launch on it
    v.table = table;
    v.table_c = c;
    return MPC_CUDA_OK;
}

}  // namespace mpc

using namespace mpc;

extern "C" int32_t mpc_cuda_msm_release_bases(uint64_t handle) {
    MPC_TRY(enter(nullptr));
    BaseRef v;
    {
        std::lock_guard<std::mutex> lk(g_bases_mu);
        auto it = g_bases.find(handle);
        if (it == g_bases.end()) {
            set_error("unknown base handle %llu", (unsigned long long)handle);
            return MPC_CUDA_ERR_HANDLE;
        }
        v = it->second;
        g_bases.erase(it);
    }
    v.reset();      // frees now unless a concurrent call still holds a reference (then when that call returns)
    return MPC_CUDA_OK;
}
