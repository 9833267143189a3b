// Share MSM on G1 / G2: signed-digit Pippenger bucket method, one pipeline of kernels per call.
//
// Replaces   VariableBaseMSM::multi_scalar_mul  arkworks/algebra/ec/src/msm/variable_base.rs:12-106
//            AffineCurve::multi_scalar_mul       arkworks/algebra/ec/src/lib.rs:305-314  (into_repr of every scalar)
//            AffineMsm::msm                      mpc-algebra/src/share/msm.rs:33-37      (normalise to affine)
// Only the affine normal form of Σ sᵢ·Pᵢ is observable, so the device algorithm differs freely from the
// reference's serial loop (signed digits, XYZZ buckets, sorted point lists) and still returns the
// bit-identical point.
//
// Pipeline (all on one stream, no host synchronisation until the result is read):
//   k_digits          Montgomery -> canonical scalar, signed c-bit digits          n x 32 B in, nwin x n x 4 B out
//   k_hist            per (window, chunk) bucket histogram in shared memory
//   k_scan_window     bucket start offsets + per-chunk scatter cursors
//   k_task_scan / k_build_tasks   cut buckets into tasks of <= task_len points (load balance)
//   k_scatter         counting sort: point indices grouped by bucket
//   k_accumulate      HOT: each thread pulls tasks and sums its points with XYZZ mixed additions
//   k_finalize_*      join the partial sums of buckets that were split
//   k_bucket_reduce   Σ b·B_b per window by chunked running sums, k_window_sum, k_horner
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ec.cuh"
#include "msm_digits.cuh"

using namespace mpc;
using Fq2 = Fp2<consts::FqParams>;

namespace mpc {
// tuning knobs (mpc_cuda_set_option); 0 = automatic
int64_t g_opt_msm_window_bits = 0;
int64_t g_opt_msm_task_len = 0;
}  // namespace mpc

namespace {

constexpr uint32_t MAX_WINDOW_BITS = 23;       // bucket id <= 22 bits = 11 partition bits + 11 bin bits
constexpr int ACC_THREADS = 128;
#ifndef ACC_MIN_BLOCKS
#define ACC_MIN_BLOCKS 3   // 168 registers, 12 warps/SM; 4 CTAs (128 regs, small spills) measured 4 % slower
#endif
constexpr int SORT_THREADS = 1024;
constexpr uint32_t SMALL_MULTI_MAX = 64;

// ---- generic 128-bit I/O for plain structs of 32-bit limbs ------------------------------------------
template <class T>
DEV T load_pod_ro(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "size");
    T r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
        uint4 t = __ldg(q + i);
        w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
    return r;
}
template <class T>
DEV T load_pod(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "size");
    T r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
        uint4 t = q[i];
        w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
    return r;
}
template <class T>
DEV void store_pod(T* p, const T& v) {
    static_assert(sizeof(T) % 16 == 0, "size");
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// ---- plan -------------------------------------------------------------------------------------------
struct Plan {
    size_t n;                    // scalars of this call
    size_t sn;                   // entries per sorted window: n, or nwin*n when all windows share one bucket set
    uint32_t snwin;              // bucket sets: nwin, or 1 with a precomputed table of 2^(c*w)*P
    uint32_t c, nwin, nb;        // window bits, windows, buckets per window = 2^(c-1)
    uint32_t hi_bits, lo_bits;   // bucket id = (hi << lo_bits) | lo: level-1 partitions / level-2 bins
    uint32_t chunks;             // level-1 sort chunks per window
    size_t chunk_len;
    uint32_t task_len;           // max points per accumulate task
    size_t max_tasks;
    uint32_t scan_blocks;        // blocks of the task scan (SCAN_ITEMS buckets each)
    uint32_t red_m, red_t;       // bucket-reduce: red_t chunks of red_m buckets per window
    uint32_t sum_parts;          // window-sum: first-level blocks per window
};

constexpr uint32_t HI_BITS_MAX = 11;            // <= 2048 write streams per CTA in the level-1 scatter
constexpr uint32_t SCAN_PER_THREAD = 8;
constexpr uint32_t SCAN_ITEMS = 1024 * SCAN_PER_THREAD;

uint32_t log2_ceil(size_t x) {
    uint32_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

Plan make_plan(size_t n, int sm_count, uint32_t table_c) {
    Plan p;
    p.n = n;
    int64_t c = table_c ? table_c : g_opt_msm_window_bits;
    if (c <= 0) {
        // madds = n * windows against bucket work ~ windows * 2^(c-1) * 3; measured on B200 with
        // tools/tune_msm.py (2^24: c = 20, 2^20: c = 17, 2^16: c = 15).  Widths whose top window holds a
        // single bit (253 mod c == 1) are skipped: they put half of a window's entries in one bucket.
        int l = (int)log2_ceil(n);
        c = l >= 23 ? 20 : l >= 21 ? 19 : l >= 19 ? 17 : l >= 17 ? 16 : l >= 14 ? 15 : l >= 11 ? 11 : l >= 8 ? 8 : 4;
    }
    if (c < 3) c = 3;
    if (c > MAX_WINDOW_BITS) c = MAX_WINDOW_BITS;
    p.c = (uint32_t)c;
    p.nwin = msm::num_windows(p.c);
    p.nb = 1u << (p.c - 1);
    uint32_t kb = p.c - 1;
    p.hi_bits = kb < HI_BITS_MAX ? kb : HI_BITS_MAX;
    p.lo_bits = kb - p.hi_bits;
    p.snwin = table_c ? 1 : p.nwin;
    p.sn = table_c ? n * p.nwin : n;
    uint32_t want = (uint32_t)((4 * sm_count + p.snwin - 1) / p.snwin);
    size_t by_len = (p.sn + 8191) / 8192;
    p.chunks = (uint32_t)(by_len < want ? by_len : want);
    if (p.chunks < 1) p.chunks = 1;
    p.chunk_len = (p.sn + p.chunks - 1) / p.chunks;
    int64_t tl = g_opt_msm_task_len;
    if (tl <= 0) {
        // enough tasks to balance the resident threads, but not so short that joins dominate
        size_t entries = n * (size_t)p.nwin;
        size_t resident = (size_t)sm_count * 384;
        tl = (int64_t)(entries / (resident * 8));
        if (tl < 32) tl = 32;
        if (tl > 256) tl = 256;
    }
    p.task_len = (uint32_t)tl;
    size_t nbuckets = (size_t)p.snwin * p.nb;
    p.max_tasks = n * (size_t)p.nwin / p.task_len + nbuckets + 1;
    p.scan_blocks = (uint32_t)((nbuckets + SCAN_ITEMS - 1) / SCAN_ITEMS);
    // bucket-reduce chunk: short chunks (more threads, shorter serial chains) while the grid stays small,
    // longer ones once there are enough chunks to fill the machine
    p.red_m = p.nb > (1u << 16) ? 64 : 32;
    while (p.red_m > 4 && (size_t)p.snwin * (p.nb / p.red_m) < (size_t)sm_count * 128) p.red_m >>= 1;
    if (p.red_m > p.nb) p.red_m = p.nb;
    p.red_t = p.nb / p.red_m;
    p.sum_parts = (p.red_t + 1023) / 1024;       // <= 1024 chunk results per first-level block
    return p;
}

// ---- sort kernels -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_digits(const Fr* __restrict__ scalars, const uint8_t* __restrict__ inf,
                                                size_t n, uint32_t c, uint32_t nwin, uint32_t* __restrict__ digits) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr s = from_mont(load_fe_ro(scalars + i));            // into_repr (ec/src/lib.rs:308-310)
        bool skip = inf && inf[i];                            // infinity bases contribute nothing
        uint32_t carry = 0;
        for (uint32_t w = 0; w < nwin; w++) {
            uint32_t d = msm::signed_digit(s.v, w, c, nwin, carry);
            digits[(size_t)w * n + i] = skip ? 0u : d;
        }
    }
}

// exclusive scan of one value per thread across the block; returns the thread's prefix, *total = block sum
DEV uint32_t block_exclusive_scan(uint32_t v, uint32_t* sm /* blockDim.x */, uint32_t* total) {
    uint32_t t = threadIdx.x;
    sm[t] = v;
    __syncthreads();
    for (uint32_t off = 1; off < blockDim.x; off <<= 1) {
        uint32_t add = t >= off ? sm[t - off] : 0;
        __syncthreads();
        sm[t] += add;
        __syncthreads();
    }
    uint32_t incl = sm[t];
    if (total) *total = sm[blockDim.x - 1];
    __syncthreads();
    return incl - v;
}

// Two-level counting sort of the (window, point) entries by bucket id = digit magnitude - 1.
// Level 1 partitions each window by the low 11 bits of the bucket id (<= 2048 partitions, so a CTA keeps
// few open write streams and its 8-byte stores merge into full sectors in L2); level 2 finishes every
// partition inside one CTA by the remaining high bits, where it also derives the bucket start / size
// tables (a bucket's run may sit anywhere in the window's list; only (start, size) matter).
// The order inside a bucket is arbitrary (atomics), which is fine: group addition commutes.

// level 1a: per (window, chunk) histogram of the partition id
__global__ void __launch_bounds__(SORT_THREADS) k_hist1(const uint32_t* __restrict__ digits, size_t n,
                                                        uint32_t lo_bits, uint32_t hbins, size_t chunk_len,
                                                        uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t sm[];
    uint32_t w = blockIdx.y, ch = blockIdx.x;
    for (uint32_t b = threadIdx.x; b < hbins; b += blockDim.x) sm[b] = 0;
    __syncthreads();
    size_t lo = (size_t)ch * chunk_len, hi = lo + chunk_len < n ? lo + chunk_len : n;
    const uint32_t* d = digits + (size_t)w * n;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        uint32_t m = d[i] & ~msm::DIGIT_NEG;
        if (m) atomicAdd(&sm[(m - 1) & (hbins - 1)], 1u);
    }
    __syncthreads();
    uint32_t* out = hist + ((size_t)w * gridDim.x + ch) * hbins;
    for (uint32_t b = threadIdx.x; b < hbins; b += blockDim.x) out[b] = sm[b];
}

// level 1b: one block per window: partition starts, and hist[w][chunk][p] -> first slot of that chunk
__global__ void __launch_bounds__(1024) k_scan1(uint32_t* __restrict__ hist, uint32_t chunks, uint32_t hbins,
                                                uint32_t* __restrict__ part_start /* [nwin][hbins + 1] */) {
    __shared__ uint32_t sm[1024];
    uint32_t w = blockIdx.x;
    uint32_t per = (hbins + blockDim.x - 1) / blockDim.x;
    uint32_t b0 = threadIdx.x * per, b1 = b0 + per < hbins ? b0 + per : hbins;
    if (b0 > hbins) b0 = hbins;
    uint32_t* h = hist + (size_t)w * chunks * hbins;
    uint32_t tot = 0;
    for (uint32_t b = b0; b < b1; b++)
        for (uint32_t ch = 0; ch < chunks; ch++) tot += h[(size_t)ch * hbins + b];
    uint32_t total;
    uint32_t run = block_exclusive_scan(tot, sm, &total);
    for (uint32_t b = b0; b < b1; b++) {
        part_start[(size_t)w * (hbins + 1) + b] = run;
        for (uint32_t ch = 0; ch < chunks; ch++) {
            uint32_t t = h[(size_t)ch * hbins + b];
            h[(size_t)ch * hbins + b] = run;
            run += t;
        }
    }
    if (threadIdx.x == 0) part_start[(size_t)w * (hbins + 1) + hbins] = total;
}

// level 1c: scatter (point index | sign, low bucket bits) pairs into their partition
__global__ void __launch_bounds__(SORT_THREADS) k_scatter1(const uint32_t* __restrict__ digits, size_t n,
                                                           uint32_t lo_bits, uint32_t hbins, size_t chunk_len,
                                                           const uint32_t* __restrict__ cursors,
                                                           uint2* __restrict__ pairs, size_t inner, size_t stride,
                                                           size_t offset) {
    extern __shared__ uint32_t sm[];
    uint32_t w = blockIdx.y, ch = blockIdx.x;
    const uint32_t* cur = cursors + ((size_t)w * gridDim.x + ch) * hbins;
    for (uint32_t b = threadIdx.x; b < hbins; b += blockDim.x) sm[b] = cur[b];
    __syncthreads();
    size_t lo = (size_t)ch * chunk_len, hi = lo + chunk_len < n ? lo + chunk_len : n;
    const uint32_t* d = digits + (size_t)w * n;
    uint2* out = pairs + (size_t)w * n;
    const uint32_t hi_bits = 31 - __clz(hbins);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        uint32_t e = d[i], m = e & ~msm::DIGIT_NEG;
        if (m) {
            uint32_t key = m - 1;
            uint32_t pos = atomicAdd(&sm[key & (hbins - 1)], 1u);
            // point index: i itself, or (window, point) -> row of the precomputed table [nwin][stride]
            uint32_t pt = stride ? (uint32_t)((i / inner) * stride + offset + i % inner) : (uint32_t)i;
            out[pos] = make_uint2(pt | (e & msm::DIGIT_NEG), key >> hi_bits);
        }
    }
}

// level 2: one CTA per (partition, window): histogram of the low bits, bucket tables, final scatter.
// A partition can be huge (tiny top window: every entry of the window shares 2-3 buckets; 0/1-heavy
// scalars), so the CTA is wide (1024 threads) and keeps SORT2_ILP independent loads in flight per
// thread, and shared atomics are warp-aggregated so a hot counter is hit once per warp.
constexpr int SORT2_THREADS = 1024;
constexpr int SORT2_ILP = 4;

__global__ void __launch_bounds__(SORT2_THREADS) k_sort2(const uint2* __restrict__ pairs,
                                                         const uint32_t* __restrict__ part_start, size_t n,
                                                         uint32_t lo_bits, uint32_t hbins, uint32_t* __restrict__ sorted,
                                                         uint32_t* __restrict__ bucket_start,
                                                         uint32_t* __restrict__ bucket_size) {
    extern __shared__ uint32_t sm[];            // lbins counters, then blockDim.x scan slots
    const uint32_t lbins = 1u << lo_bits;
    uint32_t* cnt = sm;
    uint32_t* scan = sm + lbins;
    uint32_t w = blockIdx.y, p = blockIdx.x;
    uint32_t lo = part_start[(size_t)w * (hbins + 1) + p], hi = part_start[(size_t)w * (hbins + 1) + p + 1];
    const uint2* in = pairs + (size_t)w * n;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t step = blockDim.x * SORT2_ILP;
    for (uint32_t b = threadIdx.x; b < lbins; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    for (uint32_t i0 = lo; i0 < hi; i0 += step) {
        uint32_t key[SORT2_ILP];
#pragma unroll
        for (int u = 0; u < SORT2_ILP; u++) {
            uint32_t i = i0 + u * blockDim.x + threadIdx.x;
            key[u] = i < hi ? in[i].y : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < SORT2_ILP; u++) {
            uint32_t peers = __match_any_sync(0xffffffffu, key[u]);
            if (key[u] != 0xffffffffu && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&cnt[key[u]], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    uint32_t per = (lbins + blockDim.x - 1) / blockDim.x;
    uint32_t b0 = threadIdx.x * per, b1 = b0 + per < lbins ? b0 + per : lbins;
    if (b0 > lbins) b0 = lbins;
    uint32_t tot = 0;
    for (uint32_t b = b0; b < b1; b++) tot += cnt[b];
    uint32_t run = lo + block_exclusive_scan(tot, scan, nullptr);
    // bucket id = (bin << hi_bits) | partition: level 1 splits on the LOW bits of the bucket id, which
    // stay well spread when the digits are skewed towards small values (top window, small scalars)
    const uint32_t hi_bits = 31 - __clz(hbins);
    size_t g0 = ((size_t)w << (lo_bits + hi_bits)) + p;
    for (uint32_t b = b0; b < b1; b++) {
        uint32_t t = cnt[b];
        bucket_start[g0 + ((size_t)b << hi_bits)] = run;
        bucket_size[g0 + ((size_t)b << hi_bits)] = t;
        cnt[b] = run;
        run += t;
    }
    __syncthreads();
    uint32_t* out = sorted + (size_t)w * n;
    for (uint32_t i0 = lo; i0 < hi; i0 += step) {
        uint2 e[SORT2_ILP];
#pragma unroll
        for (int u = 0; u < SORT2_ILP; u++) {
            uint32_t i = i0 + u * blockDim.x + threadIdx.x;
            e[u] = i < hi ? in[i] : make_uint2(0, 0xffffffffu);
        }
#pragma unroll
        for (int u = 0; u < SORT2_ILP; u++) {
            uint32_t peers = __match_any_sync(0xffffffffu, e[u].y);
            uint32_t leader = (uint32_t)(__ffs(peers) - 1), base = 0;
            if (e[u].y != 0xffffffffu && lane == leader) base = atomicAdd(&cnt[e[u].y], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (e[u].y != 0xffffffffu) out[base + __popc(peers & ((1u << lane) - 1u))] = e[u].x;
        }
    }
}

// ---- task table: cut buckets into runs of <= task_len entries ------------------------------------------
// (a) per-block totals of ceil(size / task_len)
__global__ void __launch_bounds__(1024) k_task_sums(const uint32_t* __restrict__ bucket_size, size_t nbuckets,
                                                    uint32_t task_len, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sm[1024];
    size_t g0 = ((size_t)blockIdx.x * 1024 + threadIdx.x) * SCAN_PER_THREAD;
    uint32_t tot = 0;
    for (uint32_t k = 0; k < SCAN_PER_THREAD; k++)
        if (g0 + k < nbuckets) tot += (bucket_size[g0 + k] + task_len - 1) / task_len;
    uint32_t total;
    block_exclusive_scan(tot, sm, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// (b) single block: exclusive scan of the block totals in place; counters[0] = number of tasks
__global__ void __launch_bounds__(1024) k_task_offsets(uint32_t* __restrict__ block_sums, uint32_t nblocks,
                                                       uint32_t* __restrict__ counters) {
    __shared__ uint32_t sm[1024];
    uint32_t per = (nblocks + blockDim.x - 1) / blockDim.x;
    uint32_t b0 = threadIdx.x * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
    if (b0 > nblocks) b0 = nblocks;
    uint32_t tot = 0;
    for (uint32_t b = b0; b < b1; b++) tot += block_sums[b];
    uint32_t total;
    uint32_t run = block_exclusive_scan(tot, sm, &total);
    for (uint32_t b = b0; b < b1; b++) {
        uint32_t t = block_sums[b];
        block_sums[b] = run;
        run += t;
    }
    if (threadIdx.x == 0) counters[0] = total;
}

// (c) task = (first slot in the window's sorted list, length, bucket, bucket-has-a-single-task)
__global__ void __launch_bounds__(1024) k_build_tasks(const uint32_t* __restrict__ bucket_start,
                                                      const uint32_t* __restrict__ bucket_size,
                                                      const uint32_t* __restrict__ block_offs, size_t nbuckets,
                                                      uint32_t task_len, uint4* __restrict__ tasks,
                                                      uint32_t* __restrict__ task_start,
                                                      uint32_t* __restrict__ small_list, uint32_t* __restrict__ big_list,
                                                      uint32_t* __restrict__ counters) {
    __shared__ uint32_t sm[1024];
    size_t g0 = ((size_t)blockIdx.x * 1024 + threadIdx.x) * SCAN_PER_THREAD;
    uint32_t tot = 0;
    for (uint32_t k = 0; k < SCAN_PER_THREAD; k++)
        if (g0 + k < nbuckets) tot += (bucket_size[g0 + k] + task_len - 1) / task_len;
    uint32_t ts = block_offs[blockIdx.x] + block_exclusive_scan(tot, sm, nullptr);
    for (uint32_t k = 0; k < SCAN_PER_THREAD; k++) {
        size_t g = g0 + k;
        if (g >= nbuckets) break;
        uint32_t size = bucket_size[g];
        uint32_t nt = (size + task_len - 1) / task_len, pos = bucket_start[g];
        task_start[g] = ts;
        for (uint32_t j = 0; j < nt; j++) {
            uint32_t len = size - j * task_len < task_len ? size - j * task_len : task_len;
            tasks[ts + j] = make_uint4(pos + j * task_len, len, (uint32_t)g, nt == 1 ? 1u : 0u);
        }
        ts += nt;
        if (nt > 1) {
            if (nt <= SMALL_MULTI_MAX) small_list[atomicAdd(&counters[2], 1u)] = (uint32_t)g;
            else big_list[atomicAdd(&counters[3], 1u)] = (uint32_t)g;
        }
    }
}

// ---- the hot kernel -----------------------------------------------------------------------------------
// Every thread repeatedly claims a task (a run of <= task_len sorted entries of one bucket) and adds the
// referenced affine bases into an XYZZ accumulator.  The claim is folded into the point loop, so the
// lanes of a warp stay converged on the mixed addition whatever the task lengths are.
template <class F>
__global__ void __launch_bounds__(ACC_THREADS, ACC_MIN_BLOCKS) k_accumulate(const Affine<F>* __restrict__ bases,
                                                            const uint32_t* __restrict__ sorted, size_t n, uint32_t nb,
                                                            const uint4* __restrict__ tasks,
                                                            uint32_t* __restrict__ counters,
                                                            XYZZ<F>* __restrict__ buckets,
                                                            XYZZ<F>* __restrict__ partials) {
    const uint32_t total = counters[0];
    XYZZ<F> acc = XYZZ<F>::infinity();
    uint32_t k = 0, len = 0, task_id = 0;
    uint4 t = make_uint4(0, 0, 0, 0);
    const uint32_t* list = sorted;
    while (true) {
        if (k == len) {
            if (len) store_pod(t.w ? buckets + t.z : partials + task_id, acc);
            task_id = atomicAdd(&counters[1], 1u);
            if (task_id >= total) break;
            t = __ldg(tasks + task_id);
            k = 0;
            len = t.y;
            list = sorted + (size_t)(t.z / nb) * n + t.x;
            acc = XYZZ<F>::infinity();
        }
        uint32_t e = __ldg(list + k);
        k++;
        Affine<F> p = load_pod_ro(bases + (e & ~msm::DIGIT_NEG));
        if (e & msm::DIGIT_NEG) p.y = neg(p.y);
        xyzz_madd(acc, p.x, p.y);
    }
}

// buckets split into 2..SMALL_MULTI_MAX tasks: one thread joins them
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_finalize_small(const uint32_t* __restrict__ small_list,
                                                                const uint32_t* __restrict__ counters,
                                                                const uint32_t* __restrict__ bucket_size,
                                                                const uint32_t* __restrict__ task_start,
                                                                uint32_t task_len, const XYZZ<F>* __restrict__ partials,
                                                                XYZZ<F>* __restrict__ buckets) {
    uint32_t count = counters[2];
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint32_t g = small_list[i];
        uint32_t nt = (bucket_size[g] + task_len - 1) / task_len, ts = task_start[g];
        XYZZ<F> acc = load_pod(partials + ts);
        for (uint32_t k = 1; k < nt; k++) {
            XYZZ<F> q = load_pod(partials + ts + k);
            xyzz_add(acc, q);
        }
        store_pod(buckets + g, acc);
    }
}

// heavily loaded buckets (skewed scalars): one block per bucket, strided partial sums + tree
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_finalize_big(const uint32_t* __restrict__ big_list,
                                                              const uint32_t* __restrict__ counters,
                                                              const uint32_t* __restrict__ bucket_size,
                                                              const uint32_t* __restrict__ task_start,
                                                              uint32_t task_len, const XYZZ<F>* __restrict__ partials,
                                                              XYZZ<F>* __restrict__ buckets) {
    extern __shared__ uint4 sm_raw[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(sm_raw);
    uint32_t count = counters[3];
    for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
        uint32_t g = big_list[i];
        uint32_t nt = (bucket_size[g] + task_len - 1) / task_len, ts = task_start[g];
        XYZZ<F> acc = XYZZ<F>::infinity();
        for (uint32_t k = threadIdx.x; k < nt; k += blockDim.x) {
            XYZZ<F> q = load_pod(partials + ts + k);
            xyzz_add(acc, q);
        }
        store_pod(sm + threadIdx.x, acc);
        __syncthreads();
        for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                XYZZ<F> a = load_pod(sm + threadIdx.x), b = load_pod(sm + threadIdx.x + s);
                xyzz_add(a, b);
                store_pod(sm + threadIdx.x, a);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            XYZZ<F> a = load_pod(sm);
            store_pod(buckets + g, a);
        }
        __syncthreads();
    }
}

// window sum Σ_b b·B_b, b = 1..nb: thread (w, t) covers buckets b = t·m + j, j = 1..m, with the
// running-sum trick (variable_base.rs:82-86) inside its chunk and adds (t·m)·Σ_j B once.
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_bucket_reduce(const XYZZ<F>* __restrict__ buckets, uint32_t nwin,
                                                               uint32_t nb, uint32_t m, uint32_t T,
                                                               XYZZ<F>* __restrict__ chunk_res) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nwin * T) return;
    uint32_t w = id / T, t = id % T;
    const XYZZ<F>* b = buckets + (size_t)w * nb + (size_t)t * m;
    XYZZ<F> running = XYZZ<F>::infinity(), acc = XYZZ<F>::infinity();
    for (uint32_t j = m; j-- > 0;) {
        XYZZ<F> q = load_pod(b + j);
        xyzz_add(running, q);
        xyzz_add(acc, running);
    }
    if (t) {
        XYZZ<F> s = xyzz_mul_small(running, (uint64_t)t * m);
        xyzz_add(acc, s);
    }
    store_pod(chunk_res + id, acc);
}

// block (part, w) sums items [part*len, (part+1)*len) of window w's `count` inputs -> out[w*parts + part]
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_window_sum(const XYZZ<F>* __restrict__ in, uint32_t count,
                                                            uint32_t len, XYZZ<F>* __restrict__ out) {
    extern __shared__ uint4 sm_raw[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(sm_raw);
    uint32_t w = blockIdx.y, part = blockIdx.x;
    uint32_t lo = part * len, hi = lo + len < count ? lo + len : count;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
        XYZZ<F> q = load_pod(in + (size_t)w * count + k);
        xyzz_add(acc, q);
    }
    store_pod(sm + threadIdx.x, acc);
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            XYZZ<F> a = load_pod(sm + threadIdx.x), b = load_pod(sm + threadIdx.x + s);
            xyzz_add(a, b);
            store_pod(sm + threadIdx.x, a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        XYZZ<F> a = load_pod(sm);
        store_pod(out + (size_t)w * gridDim.x + part, a);
    }
}

// Σ_w 2^(c·w)·S_w, high to low (variable_base.rs:92-105)
template <class F>
__global__ void k_horner(const XYZZ<F>* __restrict__ window_sum, uint32_t nwin, uint32_t c, XYZZ<F>* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    XYZZ<F> acc = load_pod(window_sum + nwin - 1);
    for (uint32_t w = nwin - 1; w-- > 0;) {
        for (uint32_t k = 0; k < c; k++) xyzz_dbl(acc);
        XYZZ<F> q = load_pod(window_sum + w);
        xyzz_add(acc, q);
    }
    store_pod(out, acc);
}

// ---- result emission ----------------------------------------------------------------------------------
// mode 0: affine x|y (2 F) followed by one 32-bit infinity flag;  mode 1: Jacobian x|y|z (3 F)
template <class F>
__global__ void k_emit(const XYZZ<F>* __restrict__ in, uint32_t count, uint32_t mode, uint32_t* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t i = 0; i < count; i++) {
        XYZZ<F> q = load_pod(in + i);
        xyzz_add(acc, q);
    }
    constexpr int N = sizeof(F) / 4;
    if (mode == 0) {
        F x, y;
        bool finite = xyzz_to_affine(acc, x, y);
        const uint32_t *px = reinterpret_cast<const uint32_t*>(&x), *py = reinterpret_cast<const uint32_t*>(&y);
        for (int i = 0; i < N; i++) { out[i] = px[i]; out[N + i] = py[i]; }
        out[2 * N] = finite ? 0u : 1u;
    } else {
        Jac<F> j = xyzz_to_jac(acc);
        const uint32_t* pj = reinterpret_cast<const uint32_t*>(&j);
        for (int i = 0; i < 3 * N; i++) out[i] = pj[i];
    }
}

// Jacobian partials (x|y|z each) -> XYZZ array
template <class F>
__global__ void k_jac_to_xyzz(const Jac<F>* __restrict__ in, uint32_t count, XYZZ<F>* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Jac<F> j = load_pod(in + i);
    XYZZ<F> p = jac_to_xyzz(j);
    store_pod(out + i, p);
}

DEV uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// synthetic CRS: out[i] = k_i * gen
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_generate(Affine<F> gen, uint64_t seed, size_t first, size_t n,
                                                          Affine<F>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = mix64(seed + (uint64_t)(first + i + 1) * 0x9E3779B97F4A7C15ull);
    if (k == 0) k = 1;
    XYZZ<F> acc = XYZZ<F>::infinity();
    // work on register copies: handing references into the 192-byte by-value parameter to the inlined
    // group law produced wrong G2 points with nvcc 12.9 (found by the GPU parity test)
    const F gx = gen.x, gy = gen.y;
    bool started = false;
    for (int b = 63; b >= 0; b--) {
        if (started) xyzz_dbl(acc);
        if ((k >> b) & 1) { xyzz_madd(acc, gx, gy); started = true; }
    }
    Affine<F> r;
    xyzz_to_affine(acc, r.x, r.y);
    store_pod(out + i, r);
}

// ---- host pipeline ------------------------------------------------------------------------------------
template <class K>
int32_t allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) MPC_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return MPC_CUDA_OK;
}

// A registered base vector may carry a table T[w][i] = 2^(c*w) * P_i (affine).  Then every window adds
// into ONE shared bucket set (Σ_w 2^(cw) d_w P = Σ_w d_w T[w]), the bucket reduction runs once instead of
// per window and the Horner doublings disappear, which makes wide windows (c = 22) pay: ~25 % fewer
// mixed additions at n = 2^24.  180 GB of HBM is what makes a 12x CRS copy affordable.
struct TableRef {
    const void* table = nullptr;   // Affine<F>[nwin][stride]
    uint32_t c = 0;
    size_t stride = 0, offset = 0;
};

// result (one XYZZ on the device) = Σ scalars[i] * bases[i]
template <class F>
int32_t msm_run(const Affine<F>* bases, const uint8_t* inf, const Fr* scalars, size_t n, XYZZ<F>* result,
                cudaStream_t s, const TableRef* tbl = nullptr) {
    if (n == 0) {
        MPC_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(XYZZ<F>), s));
        return MPC_CUDA_OK;
    }
    MPC_ARG_CHECK(n < ((size_t)1 << 31));
    const DeviceInfo* dev = current_device_info();
    const bool use_table = tbl && tbl->table;
    Plan p = make_plan(n, dev->sm_count, use_table ? tbl->c : 0);
    const size_t sn = p.sn;
    const uint32_t snwin = p.snwin;
    size_t nbuckets = (size_t)snwin * p.nb;
    MPC_ARG_CHECK(sn < ((size_t)1 << 31) && (!use_table || (size_t)p.nwin * tbl->stride < ((size_t)1 << 31)));
    MPC_ARG_CHECK(nbuckets < ((size_t)1 << 31) && (size_t)p.nwin * n / p.task_len + nbuckets < ((size_t)1 << 32));

    Scratch s_digits, s_sorted, s_pairs, s_hist, s_part, s_bstart, s_bsize, s_tstart, s_bsums, s_tasks, s_small, s_big,
        s_cnt, s_buckets, s_partials, s_chunk, s_wpart, s_wsum;
    uint32_t *digits, *sorted, *hist, *part_start, *bstart, *bsize, *tstart, *bsums, *small_list, *big_list, *counters;
    uint2* pairs;
    uint4* tasks;
    XYZZ<F>*buckets, *partials, *chunk_res, *wpart, *wsum;
    const uint32_t hbins = 1u << p.hi_bits, lbins = 1u << p.lo_bits;
    MPC_TRY(s_digits.alloc(&digits, (size_t)p.nwin * n, s));
    MPC_TRY(s_sorted.alloc(&sorted, (size_t)p.nwin * n, s));
    MPC_TRY(s_pairs.alloc(&pairs, (size_t)p.nwin * n, s));
    MPC_TRY(s_hist.alloc(&hist, (size_t)snwin * p.chunks * hbins, s));
    MPC_TRY(s_part.alloc(&part_start, (size_t)snwin * (hbins + 1), s));
    MPC_TRY(s_bstart.alloc(&bstart, nbuckets, s));
    MPC_TRY(s_bsize.alloc(&bsize, nbuckets, s));
    MPC_TRY(s_tstart.alloc(&tstart, nbuckets, s));
    MPC_TRY(s_bsums.alloc(&bsums, p.scan_blocks, s));
    MPC_TRY(s_tasks.alloc(&tasks, p.max_tasks, s));
    MPC_TRY(s_small.alloc(&small_list, nbuckets, s));
    MPC_TRY(s_big.alloc(&big_list, nbuckets, s));
    MPC_TRY(s_cnt.alloc(&counters, 8, s));
    MPC_TRY(s_buckets.alloc(&buckets, nbuckets, s));
    MPC_TRY(s_partials.alloc(&partials, p.max_tasks, s));
    MPC_TRY(s_chunk.alloc(&chunk_res, (size_t)snwin * p.red_t, s));
    MPC_TRY(s_wpart.alloc(&wpart, (size_t)snwin * p.sum_parts, s));
    MPC_TRY(s_wsum.alloc(&wsum, snwin, s));

    ProfileScope prof_total("msm_total", s);
    profile_begin("msm_sort", s);
    MPC_CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(uint32_t), s));
    MPC_CUDA_TRY(cudaMemsetAsync(buckets, 0, nbuckets * sizeof(XYZZ<F>), s));      // all-zero = infinity

    // digits[w*n + i]: with a table the flat array IS one window of nwin*n entries
    k_digits<<<grid_for(n, 256, 8), 256, 0, s>>>(scalars, inf, n, p.c, p.nwin, digits);
    MPC_KERNEL_CHECK();

    dim3 grid1(p.chunks, snwin);
    k_hist1<<<grid1, SORT_THREADS, hbins * sizeof(uint32_t), s>>>(digits, sn, p.lo_bits, hbins, p.chunk_len, hist);
    MPC_KERNEL_CHECK();
    k_scan1<<<snwin, 1024, 0, s>>>(hist, p.chunks, hbins, part_start);
    MPC_KERNEL_CHECK();
    k_scatter1<<<grid1, SORT_THREADS, hbins * sizeof(uint32_t), s>>>(digits, sn, p.lo_bits, hbins, p.chunk_len, hist, pairs,
                                                                    n, use_table ? tbl->stride : 0,
                                                                    use_table ? tbl->offset : 0);
    MPC_KERNEL_CHECK();
    k_sort2<<<dim3(hbins, snwin), SORT2_THREADS, (lbins + SORT2_THREADS) * sizeof(uint32_t), s>>>(pairs, part_start, sn, p.lo_bits, hbins, sorted,
                                                                             bstart, bsize);
    MPC_KERNEL_CHECK();
    k_task_sums<<<p.scan_blocks, 1024, 0, s>>>(bsize, nbuckets, p.task_len, bsums);
    MPC_KERNEL_CHECK();
    k_task_offsets<<<1, 1024, 0, s>>>(bsums, p.scan_blocks, counters);
    MPC_KERNEL_CHECK();
    k_build_tasks<<<p.scan_blocks, 1024, 0, s>>>(bstart, bsize, bsums, nbuckets, p.task_len, tasks, tstart, small_list,
                                                 big_list, counters);
    MPC_KERNEL_CHECK();
    profile_end("msm_sort", s);

    int acc_blocks = 0;
    MPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_blocks, k_accumulate<F>, ACC_THREADS, 0));
    if (acc_blocks < 1) acc_blocks = 1;
    profile_begin("msm_accumulate", s);
    k_accumulate<F><<<dev->sm_count * acc_blocks, ACC_THREADS, 0, s>>>(
        use_table ? (const Affine<F>*)tbl->table : bases, sorted, sn, p.nb, tasks, counters, buckets, partials);
    MPC_KERNEL_CHECK();
    profile_end("msm_accumulate", s);
    profile_begin("msm_reduce", s);

    k_finalize_small<F><<<dev->sm_count * 2, ACC_THREADS, 0, s>>>(small_list, counters, bsize, tstart, p.task_len,
                                                                  partials, buckets);
    MPC_KERNEL_CHECK();
    size_t tree_smem = ACC_THREADS * sizeof(XYZZ<F>);
    MPC_TRY(allow_smem(k_finalize_big<F>, tree_smem));
    MPC_TRY(allow_smem(k_window_sum<F>, tree_smem));
    k_finalize_big<F><<<dev->sm_count, ACC_THREADS, tree_smem, s>>>(big_list, counters, bsize, tstart, p.task_len,
                                                                    partials, buckets);
    MPC_KERNEL_CHECK();

    uint32_t red_threads = snwin * p.red_t;
    k_bucket_reduce<F><<<(red_threads + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, s>>>(buckets, snwin, p.nb,
                                                                                            p.red_m, p.red_t, chunk_res);
    MPC_KERNEL_CHECK();
    k_window_sum<F><<<dim3(p.sum_parts, snwin), ACC_THREADS, tree_smem, s>>>(chunk_res, p.red_t, 1024, wpart);
    MPC_KERNEL_CHECK();
    k_window_sum<F><<<dim3(1, snwin), ACC_THREADS, tree_smem, s>>>(wpart, p.sum_parts, p.sum_parts, wsum);
    MPC_KERNEL_CHECK();
    k_horner<F><<<1, 32, 0, s>>>(wsum, snwin, p.c, result);
    MPC_KERNEL_CHECK();
    profile_end("msm_reduce", s);
    return MPC_CUDA_OK;
}

// ---- registered base vectors --------------------------------------------------------------------------
struct BaseVec {
    void* bases = nullptr;       // Affine<F>[n] on `cuda_device`
    uint8_t* inf = nullptr;      // n flags or nullptr
    size_t n = 0;
    int cuda_device = 0;
    bool g2 = false;
    bool owned = true;
    void* table = nullptr;       // Affine<F>[nwin(table_c)][n]: 2^(table_c*w) * P_i, or nullptr
    uint32_t table_c = 0;
};
std::mutex g_bases_mu;
std::unordered_map<uint64_t, BaseVec> g_bases;
uint64_t g_next_handle = 1;

template <class F>
int32_t register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(handle && (n == 0 || bases_xy));
    BaseVec v;
    v.n = n;
    v.g2 = g2;
    v.cuda_device = current_device_info()->cuda_device;
    MPC_CUDA_TRY(cudaMalloc(&v.bases, (n ? n : 1) * sizeof(Affine<F>)));
    MPC_CUDA_TRY(cudaMemcpyAsync(v.bases, bases_xy, n * sizeof(Affine<F>), cudaMemcpyHostToDevice, s));
    if (inf) {
        bool any = false;
        for (size_t i = 0; i < n && !any; i++) any = inf[i] != 0;
        if (any) {
            MPC_CUDA_TRY(cudaMalloc((void**)&v.inf, n));
            MPC_CUDA_TRY(cudaMemcpyAsync(v.inf, inf, n, cudaMemcpyHostToDevice, s));
        }
    }
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> lk(g_bases_mu);
    *handle = g_next_handle++;
    g_bases[*handle] = v;
    return MPC_CUDA_OK;
}

int32_t find_bases(uint64_t handle, bool g2, size_t offset, size_t n, BaseVec* out) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    auto it = g_bases.find(handle);
    if (it == g_bases.end() || it->second.g2 != g2) {
        set_error("unknown %s base handle %llu", g2 ? "G2" : "G1", (unsigned long long)handle);
        return MPC_CUDA_ERR_HANDLE;
    }
    if (offset > it->second.n || n > it->second.n - offset) {
        set_error("base range [%zu, %zu) outside registered vector of %zu points", offset, offset + n, it->second.n);
        return MPC_CUDA_ERR_ARG;
    }
    if (it->second.cuda_device != current_device_info()->cuda_device) {
        set_error("base handle %llu lives on CUDA device %d, calling thread uses %d", (unsigned long long)handle,
                  it->second.cuda_device, current_device_info()->cuda_device);
        return MPC_CUDA_ERR_HANDLE;
    }
    *out = it->second;
    return MPC_CUDA_OK;
}

TableRef table_of(const BaseVec& v, size_t offset) {
    TableRef t;
    t.table = v.table;
    t.c = v.table_c;
    t.stride = v.n;
    t.offset = offset;
    return t;
}

// T[w][i] = 2^(c*w) * P_i in affine form, one thread per base (c doublings per window, one inversion per
// entry; run once per registered CRS)
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_precompute(const Affine<F>* __restrict__ bases,
                                                            const uint8_t* __restrict__ inf, size_t n, uint32_t c,
                                                            uint32_t nwin, Affine<F>* __restrict__ table) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> pt = load_pod_ro(bases + i);
    store_pod(table + i, pt);
    if (inf && inf[i]) return;                 // never referenced: its digits are zero
    XYZZ<F> acc;
    acc.x = pt.x; acc.y = pt.y; acc.zz = F::one(); acc.zzz = F::one();
    for (uint32_t w = 1; w < nwin; w++) {
        for (uint32_t k = 0; k < c; k++) xyzz_dbl(acc);
        Affine<F> r;
        xyzz_to_affine(acc, r.x, r.y);
        store_pod(table + (size_t)w * n + i, r);
        acc.x = r.x; acc.y = r.y; acc.zz = F::one(); acc.zzz = F::one();
    }
}

template <class F>
int32_t precompute(uint64_t handle, uint32_t window_bits, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(window_bits == 0 || (window_bits >= 3 && window_bits <= MAX_WINDOW_BITS));
    BaseVec v;
    MPC_TRY(find_bases(handle, g2, 0, 0, &v));
    uint32_t c = window_bits;
    if (c == 0) {
        uint32_t l = log2_ceil(v.n ? v.n : 1);
        c = l >= 23 ? 23 : l >= 22 ? 22 : l >= 20 ? 20 : l >= 18 ? 17 : l >= 16 ? 15 : l > 7 ? l - 3 : 4;
        if (msm::SCALAR_BITS % c == 1) c--;       // a one-bit top window would put n/2 entries in one bucket
    }
    uint32_t nwin = msm::num_windows(c);
    MPC_ARG_CHECK((size_t)nwin * v.n < ((size_t)1 << 31));
    void* table = nullptr;
    MPC_CUDA_TRY(cudaMalloc(&table, (size_t)nwin * (v.n ? v.n : 1) * sizeof(Affine<F>)));
    if (v.n) {
        k_precompute<F><<<(unsigned)((v.n + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, s>>>(
            (const Affine<F>*)v.bases, v.inf, v.n, c, nwin, (Affine<F>*)table);
        MPC_KERNEL_CHECK();
    }
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    void* old = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_bases_mu);
        auto it = g_bases.find(handle);
        if (it == g_bases.end()) {
            cudaFree(table);
            set_error("base handle %llu released during precomputation", (unsigned long long)handle);
            return MPC_CUDA_ERR_HANDLE;
        }
        old = it->second.table;
        it->second.table = table;
        it->second.table_c = c;
    }
    if (old) {
        MPC_CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(old);
    }
    return MPC_CUDA_OK;
}

// run + emit; scalars already on the device.  host_out: affine limbs + flag copied back and synchronised.
template <class F>
int32_t msm_emit(const Affine<F>* bases, const uint8_t* inf, const Fr* scalars_dev, size_t n, uint32_t mode,
                 uint32_t* out_dev, uint64_t* host_xy, uint8_t* host_inf, cudaStream_t s, const TableRef* tbl = nullptr) {
    constexpr int N = sizeof(F) / 4;
    Scratch s_res, s_out;
    XYZZ<F>* res;
    MPC_TRY(s_res.alloc(&res, 1, s));
    MPC_TRY(msm_run<F>(bases, inf, scalars_dev, n, res, s, tbl));
    uint32_t* out = out_dev;
    if (!out) MPC_TRY(s_out.alloc(&out, 3 * N + 4, s));
    k_emit<F><<<1, 32, 0, s>>>(res, 1, mode, out);
    MPC_KERNEL_CHECK();
    if (host_xy) {
        uint32_t host[2 * N + 1];
        MPC_CUDA_TRY(cudaMemcpyAsync(host, out, sizeof(host), cudaMemcpyDeviceToHost, s));
        MPC_CUDA_TRY(cudaStreamSynchronize(s));
        memcpy(host_xy, host, 2 * N * sizeof(uint32_t));
        if (host_inf) *host_inf = (uint8_t)host[2 * N];
    }
    return MPC_CUDA_OK;
}

template <class F>
int32_t msm_host(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                 uint8_t* out_inf) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_xy && out_inf && (n == 0 || (bases_xy && scalars)));
    Scratch sb, si, ss;
    Affine<F>* db;
    uint8_t* di = nullptr;
    Fr* dsc;
    MPC_TRY(sb.alloc(&db, n, s));
    MPC_TRY(ss.alloc(&dsc, n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(db, bases_xy, n * sizeof(Affine<F>), cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dsc, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    if (inf && n) {
        MPC_TRY(si.alloc(&di, n, s));
        MPC_CUDA_TRY(cudaMemcpyAsync(di, inf, n, cudaMemcpyHostToDevice, s));
    }
    return msm_emit<F>(db, di, dsc, n, 0, nullptr, out_xy, out_inf, s);
}

template <class F>
int32_t msm_handle_host(uint64_t handle, size_t offset, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                        uint8_t* out_inf, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_xy && out_inf && (n == 0 || scalars));
    BaseVec v;
    MPC_TRY(find_bases(handle, g2, offset, n, &v));
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

}  // namespace

extern "C" {

int32_t mpc_cuda_msm_g1(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars_mont, size_t n,
                        uint64_t out_xy[12], uint8_t* out_inf) {
    return msm_host<Fq>(bases_xy, inf, scalars_mont, n, out_xy, out_inf);
}

int32_t mpc_cuda_msm_g2(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars_mont, size_t n,
                        uint64_t out_xy[24], uint8_t* out_inf) {
    return msm_host<Fq2>(bases_xy, inf, scalars_mont, n, out_xy, out_inf);
}

int32_t mpc_cuda_msm_g1_register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle) {
    return register_bases<Fq>(bases_xy, inf, n, handle, false);
}

int32_t mpc_cuda_msm_g2_register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle) {
    return register_bases<Fq2>(bases_xy, inf, n, handle, true);
}

int32_t mpc_cuda_msm_g1_register_bases_dev(const uint64_t* bases_xy_dev, size_t n, uint64_t* handle) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(handle && (n == 0 || bases_xy_dev));
    BaseVec v;
    v.bases = (void*)bases_xy_dev;
    v.n = n;
    v.cuda_device = current_device_info()->cuda_device;
    v.owned = false;
    std::lock_guard<std::mutex> lk(g_bases_mu);
    *handle = g_next_handle++;
    g_bases[*handle] = v;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_msm_release_bases(uint64_t handle) {
    MPC_TRY(enter(nullptr));
    BaseVec v;
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
    int cur = 0;
    MPC_CUDA_TRY(cudaGetDevice(&cur));
    MPC_CUDA_TRY(cudaSetDevice(v.cuda_device));
    if (v.owned) {
        cudaFree(v.bases);
        if (v.inf) cudaFree(v.inf);
    }
    if (v.table) cudaFree(v.table);
    MPC_CUDA_TRY(cudaSetDevice(cur));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_msm_g1_precompute(uint64_t handle, uint32_t window_bits) {
    return precompute<Fq>(handle, window_bits, false);
}

int32_t mpc_cuda_msm_g2_precompute(uint64_t handle, uint32_t window_bits) {
    return precompute<Fq2>(handle, window_bits, true);
}

int32_t mpc_cuda_msm_g1_handle(uint64_t handle, size_t offset, const uint64_t* scalars_mont, size_t n,
                               uint64_t out_xy[12], uint8_t* out_inf) {
    return msm_handle_host<Fq>(handle, offset, scalars_mont, n, out_xy, out_inf, false);
}

int32_t mpc_cuda_msm_g2_handle(uint64_t handle, size_t offset, const uint64_t* scalars_mont, size_t n,
                               uint64_t out_xy[24], uint8_t* out_inf) {
    return msm_handle_host<Fq2>(handle, offset, scalars_mont, n, out_xy, out_inf, true);
}

int32_t mpc_cuda_msm_g1_handle_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                   uint64_t* out_jac_dev, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_jac_dev && (n == 0 || scalars_mont_dev));
    BaseVec v;
    MPC_TRY(find_bases(handle, false, offset, n, &v));
    TableRef tbl = table_of(v, offset);
    return msm_emit<Fq>((const Affine<Fq>*)v.bases + offset, v.inf ? v.inf + offset : nullptr,
                        (const Fr*)scalars_mont_dev, n, 1, (uint32_t*)out_jac_dev, nullptr, nullptr,
                        pick_stream(stream, s), &tbl);
}

int32_t mpc_cuda_g1_sum_partials_dev(const uint64_t* jac_dev, uint32_t count, uint64_t out_xy[12], uint8_t* out_inf,
                                     void* stream) {
    cudaStream_t s0;
    MPC_TRY(enter(&s0));
    cudaStream_t s = pick_stream(stream, s0);
    MPC_ARG_CHECK(out_xy && out_inf && (count == 0 || jac_dev));
    Scratch sx, so;
    XYZZ<Fq>* pts;
    uint32_t* out;
    MPC_TRY(sx.alloc(&pts, count, s));
    MPC_TRY(so.alloc(&out, 40, s));
    if (count) {
        k_jac_to_xyzz<Fq><<<(count + 127) / 128, 128, 0, s>>>((const Jac<Fq>*)jac_dev, count, pts);
        MPC_KERNEL_CHECK();
    }
    k_emit<Fq><<<1, 32, 0, s>>>(pts, count, 0, out);
    MPC_KERNEL_CHECK();
    uint32_t host[25];
    MPC_CUDA_TRY(cudaMemcpyAsync(host, out, sizeof(host), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_xy, host, 24 * sizeof(uint32_t));
    *out_inf = (uint8_t)host[24];
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_g1_generate_dev(uint64_t seed, size_t first, size_t n, uint64_t* out_xy_dev, void* stream) {
    return generate<Fq>(consts::G1_GEN_X, consts::G1_GEN_Y, seed, first, n, out_xy_dev, stream);
}

int32_t mpc_cuda_g2_generate_dev(uint64_t seed, size_t first, size_t n, uint64_t* out_xy_dev, void* stream) {
    return generate<Fq2>(consts::G2_GEN_X, consts::G2_GEN_Y, seed, first, n, out_xy_dev, stream);
}

}  // extern "C"
