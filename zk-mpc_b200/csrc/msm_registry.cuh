// Registry of base vectors kept resident on the devices (pk.*_query, powers_of_g): shared by the G1 and G2
// translation units.  Entries are reference counted: a call copies the shared_ptr under the registry lock
// and holds it until its kernels are enqueued, so release / table replacement by another party thread can
// never free memory a launch is about to use (cudaFree itself waits for work already enqueued).
#pragma once
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpc {

struct BaseVec {
    void* bases = nullptr;       // Affine<F>[n] on `cuda_device`
    uint8_t* inf = nullptr;      // n flags or nullptr
    size_t n = 0;
    int dev_index = 0;           // index into the init list
    int cuda_device = 0;
    bool g2 = false;
    bool owned = true;
    void* table = nullptr;       // Affine<F>[nwin(table_c)][n]: 2^(table_c*w) * P_i, or nullptr
    uint32_t table_c = 0;
    std::vector<void*> retired;  // superseded tables: freed with the vector, never while it is published
    // point-range sharding over several devices: part k holds [lo[k], lo[k+1]) on device k
    std::vector<std::shared_ptr<BaseVec>> parts;
    std::vector<size_t> lo;
    ~BaseVec();
};
using BaseRef = std::shared_ptr<BaseVec>;

// consistent copy of the fields a launch needs, taken under the registry lock
struct BaseSnap {
    BaseRef ref;
    const void* bases = nullptr;
    const uint8_t* inf = nullptr;
    const void* table = nullptr;
    uint32_t table_c = 0;
    size_t n = 0;
};

uint64_t registry_add(const BaseRef& v);
// checks curve, range and (for single-device vectors) that the vector lives on the calling thread's device
int32_t registry_find(uint64_t handle, bool g2, size_t offset, size_t n, BaseSnap* out);
int32_t registry_set_table(uint64_t handle, void* table, uint32_t c);
BaseSnap snapshot_of(const BaseRef& v);

}  // namespace mpc
