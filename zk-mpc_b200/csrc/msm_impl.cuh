// Share MSM on G1 / G2: signed-digit Pippenger bucket method, one pipeline of kernels per call.
// Implementation header: included by msm_g1.cu (F = Fq) and msm_g2.cu (F = Fq2), which export the C ABI.
//
// Replaces   VariableBaseMSM::multi_scalar_mul  arkworks/algebra/ec/src/msm/variable_base.rs:12-106
//            AffineCurve::multi_scalar_mul       arkworks/algebra/ec/src/lib.rs:305-314  (into_repr of every scalar)
//            AffineMsm::msm                      mpc-algebra/src/share/msm.rs:33-37      (normalise to affine)
// Only the affine normal form of Σ sᵢ·Pᵢ is observable, so the device algorithm differs freely from the
// reference's serial loop (signed digits, XYZZ buckets, sorted point lists) and still returns the
// bit-identical point.
//
// Pipeline (all on one stream, no host synchronisation until the result is read).  The points may arrive in
// several CHUNKS (point ranges) that add into the same bucket set, so a host-buffer call overlaps the PCIe
// copy of chunk j+1 with the sort and accumulation of chunk j:
//   per chunk   k_digits          Montgomery -> canonical scalar, signed c-bit digits
//               k_hist1/k_scan1/k_scatter1/k_sort2   two-level counting sort by bucket
//               k_task_*          cut buckets into tasks of <= task_len points (load balance)
//               k_accumulate      HOT: each thread pulls tasks and sums its points with XYZZ mixed additions
//               k_finalize_*      join the partial sums of buckets that were split
//   once        k_bucket_reduce   Σ b·B_b per window by chunked running sums, k_window_sum
//               k_horner          warp-cooperative Σ_w 2^(cw) S_w        k_emit   affine / Jacobian result
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "msm_digits.cuh"
#include "msm_registry.cuh"

using namespace mpc;
using Fq2 = Fp2<consts::FqParams>;

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
    size_t n;                    // scalars of the whole MSM
    size_t cap;                  // longest chunk (point range) the scratch is sized for
    size_t sn_cap;               // entries per sorted window of a longest chunk: cap, or nwin*cap with a table
    uint32_t snwin;              // bucket sets: nwin, or 1 with a precomputed table of 2^(c*w)*P
    uint32_t c, nwin, nb;        // window bits, windows, buckets per window = 2^(c-1)
    uint32_t hi_bits, lo_bits;   // bucket id = (hi << lo_bits) | lo: level-1 partitions / level-2 bins
    uint32_t chunks;             // level-1 sort chunks per window
    uint32_t task_len;           // max points per accumulate task
    bool aff_auto = true;        // the pre-reduction rounds were chosen by the size rule: short chunks choose again
    size_t max_tasks;
    uint32_t scan_blocks;        // blocks of the task scan (SCAN_ITEMS buckets each)
    uint32_t red_m, red_t;       // bucket-reduce: red_t chunks of red_m buckets per window
    uint32_t sum_parts;          // window-sum: first-level blocks per window
    uint32_t aff;                // affine pre-reduction rounds (0 or 2): bucket runs padded to 2^aff entries
    size_t snp;                  // slots per window of the sorted list (= sn_cap when aff == 0)
};

constexpr uint32_t HI_BITS_MAX = 11;            // <= 2048 write streams per CTA in the level-1 scatter
constexpr uint32_t SCAN_PER_THREAD = 8;
constexpr uint32_t SCAN_ITEMS = 1024 * SCAN_PER_THREAD;

inline uint32_t log2_ceil(size_t x) {
    uint32_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

inline Plan make_plan(size_t n, size_t cap, int sm_count, uint32_t table_c) {
    Plan p;
    p.n = n;
    p.cap = cap;
    int64_t c = table_c ? table_c : g_opt_msm_window_bits.load(std::memory_order_relaxed);
    if (c <= 0) {
        // madds = n * windows against bucket work ~ windows * 2^(c-1) * 3; measured on B200 with
        // tools/tune_msm.py (2^24: c = 20, 2^20: c = 17, 2^16: c = 15).  Widths whose top window holds a
        // single bit (253 mod c == 1) are skipped: they put half of a window's entries in one bucket.
        int l = (int)log2_ceil(n);
        c = l >= 23 ? 20 : l >= 21 ? 19 : l >= 19 ? 17 : l >= 17 ? 16 : l >= 14 ? 15 : l >= 11 ? 9 : l >= 8 ? 8 : 4;
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
    p.sn_cap = table_c ? cap * p.nwin : cap;
    uint32_t want = (uint32_t)((4 * sm_count + p.snwin - 1) / p.snwin);
    size_t by_len = (p.sn_cap + 8191) / 8192;
    p.chunks = (uint32_t)(by_len < want ? by_len : want);
    if (p.chunks < 1) p.chunks = 1;
    // Affine pre-reduction: measured on B200 (tools/affine_ab.py, profiles/r2_affine_ab.jsonl) two rounds win from
    // ~2^27 sorted entries (2^24 points: 108.5 -> 96.3 ms plain, 90.8 -> 82.8 ms with the table), one round between
    // 2^25.5 and 2^27 (2^22 points: 34.0 -> 31.7 ms), none below (the shared inversions stop amortising); the
    // padding of every run to 2^aff entries also needs buckets of >= 8 entries on average.
    size_t nbuckets_all = (size_t)p.snwin * p.nb, entries_all = cap * (size_t)p.nwin;
    int64_t aff_opt = g_opt_msm_affine.load(std::memory_order_relaxed);
    p.aff = 0;
    p.aff_auto = aff_opt == 0;
    if (aff_opt == 0 && entries_all >= 8 * nbuckets_all)
        p.aff = entries_all >= ((size_t)3 << 25) ? 2 : entries_all >= ((size_t)3 << 24) ? 1 : 0;
    else if (aff_opt >= 2) p.aff = (uint32_t)(aff_opt - 1);              // 2 -> one round, 3 -> two rounds
    p.snp = p.sn_cap;
    if (p.aff) {
        size_t padm = ((size_t)1 << p.aff) - 1, hb = (size_t)1 << p.hi_bits, lb = (size_t)1 << p.lo_bits;
        p.snp = (p.sn_cap + hb * (padm * lb + padm + 1) + padm + padm) & ~padm;
    }
    int64_t tl = g_opt_msm_task_len.load(std::memory_order_relaxed);
    if (tl <= 0) {
        // enough tasks to balance the resident threads, but not so short that joins dominate
        size_t entries = (cap * (size_t)p.nwin) >> p.aff;
        size_t resident = (size_t)sm_count * 384;
        tl = (int64_t)(entries / (resident * 8));
        if (tl < 16) tl = 16;                       // 2^13 points, table: 1.04 ms at 16 against 1.13 at 32 (G2: 3.13 / 3.63)
        if (tl > 256) tl = 256;
    }
    p.task_len = (uint32_t)tl;
    size_t nbuckets = (size_t)p.snwin * p.nb;
    p.max_tasks = cap * (size_t)p.nwin / p.task_len + nbuckets + 1;
    p.scan_blocks = (uint32_t)((nbuckets + SCAN_ITEMS - 1) / SCAN_ITEMS);
    // bucket-reduce chunk: short chunks (more threads, shorter serial chains) while the grid stays small,
    // longer ones once there are enough chunks to fill the machine
    p.red_m = p.nb > (1u << 16) ? 64 : 32;
    while (p.red_m > 4 && (size_t)p.snwin * (p.nb / p.red_m) < (size_t)sm_count * 128) p.red_m >>= 1;
    const int64_t red_opt = g_opt_msm_reduce_chunk.load(std::memory_order_relaxed);
    if (red_opt > 0) p.red_m = (uint32_t)red_opt;
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

// pad_log > 0 (affine pre-reduction, below): every bucket's run starts at a multiple of 2^pad_log in the output
// list and is padded to such a multiple with holes (the list is pre-filled with 0xffffffff); the window's list
// then has `out_stride` slots, and bucket_start / bucket_size are written in units of 2^pad_log entries, i.e. as
// positions in the array of pre-reduced points.
__global__ void __launch_bounds__(SORT2_THREADS) k_sort2(const uint2* __restrict__ pairs,
                                                         const uint32_t* __restrict__ part_start, size_t n,
                                                         uint32_t lo_bits, uint32_t hbins, uint32_t* __restrict__ sorted,
                                                         uint32_t* __restrict__ bucket_start,
                                                         uint32_t* __restrict__ bucket_size, uint32_t pad_log,
                                                         size_t out_stride) {
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
    const uint32_t padm = (1u << pad_log) - 1u;
    uint32_t tot = 0;
    for (uint32_t b = b0; b < b1; b++) tot += (cnt[b] + padm) & ~padm;
    // the partition's output region: exact lists start where the input run starts; padded lists leave room for
    // the worst-case padding of every earlier partition of the window
    uint32_t obase = pad_log ? ((lo + p * (padm * lbins + padm + 1u) + padm) & ~padm) : lo;
    uint32_t run = obase + block_exclusive_scan(tot, scan, nullptr);
    // bucket id = (bin << hi_bits) | partition: level 1 splits on the LOW bits of the bucket id, which
    // stay well spread when the digits are skewed towards small values (top window, small scalars)
    const uint32_t hi_bits = 31 - __clz(hbins);
    size_t g0 = ((size_t)w << (lo_bits + hi_bits)) + p;
    for (uint32_t b = b0; b < b1; b++) {
        uint32_t t = cnt[b], tp = (t + padm) & ~padm;
        bucket_start[g0 + ((size_t)b << hi_bits)] = run >> pad_log;
        bucket_size[g0 + ((size_t)b << hi_bits)] = tp >> pad_log;
        cnt[b] = run;
        run += tp;
    }
    __syncthreads();
    uint32_t* out = sorted + (size_t)w * out_stride;
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

// (b) single block: exclusive scan of the block totals in place; counters[0] = number of tasks, the claim
// counter and the split-bucket list lengths restart at zero (one counter block serves every chunk)
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
    if (threadIdx.x == 0) { counters[0] = total; counters[1] = 0; counters[2] = 0; counters[3] = 0; }
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
// merge != 0 (second and later chunks of a streamed MSM): a bucket's accumulator starts from the value the
// earlier chunks left in it instead of infinity — no extra field products.
// DENSE: the tasks run over an array of pre-reduced affine points (k_affine_pairs) instead of an index list into
// the bases: sequential 96-byte reads, entries (0, 0) are holes.
template <class F, bool DENSE>
__global__ void __launch_bounds__(ACC_THREADS, sizeof(F) > 48 ? 2 : ACC_MIN_BLOCKS) k_accumulate(const Affine<F>* __restrict__ bases,
                                                            const uint32_t* __restrict__ sorted, size_t n, uint32_t nb,
                                                            const uint4* __restrict__ tasks,
                                                            uint32_t* __restrict__ counters,
                                                            XYZZ<F>* __restrict__ buckets,
                                                            XYZZ<F>* __restrict__ partials, uint32_t merge) {
    const uint32_t total = counters[0];
    XYZZ<F> acc = XYZZ<F>::infinity();
    uint32_t k = 0, len = 0, task_id = 0;
    uint4 t = make_uint4(0, 0, 0, 0);
    size_t first = 0;                     // slot of the task's first entry in the window-major list / point array
    while (true) {
        if (k == len) {
            if (len) store_pod(t.w ? buckets + t.z : partials + task_id, acc);
            task_id = atomicAdd(&counters[1], 1u);
            if (task_id >= total) break;
            t = __ldg(tasks + task_id);
            k = 0;
            len = t.y;
            first = (size_t)(t.z / nb) * n + t.x;
            if (merge && t.w) acc = load_pod(buckets + t.z);
            else acc = XYZZ<F>::infinity();
        }
        if (DENSE) {
            Affine<F> p = load_pod_ro(bases + first + k);
            k++;
            if (p.x.is_zero() && p.y.is_zero()) continue;
            xyzz_madd_fast(acc, p.x, p.y);
        } else {
            uint32_t e = __ldg(sorted + first + k);
            k++;
            Affine<F> p = load_pod_ro(bases + (e & ~msm::DIGIT_NEG));
            if (e & msm::DIGIT_NEG) p.y = neg(p.y);
            xyzz_madd_fast(acc, p.x, p.y);
        }
    }
}

// ---- affine pre-reduction --------------------------------------------------------------------------------
// Bucket accumulation sits on the IMAD.WIDE pipe at ~0.85 of its ceiling, so the only way to make it faster is
// fewer field products per point.  A mixed XYZZ addition costs 10; an affine chord addition costs 3 once the
// inverse of x1 - x0 is known, and Montgomery's trick (ff/src/fields/mod.rs:597-660) turns one inversion into 3
// products per element of a batch.  The sorted lists are laid out with every bucket's run padded to a multiple
// of 4 (k_sort2, pad_log = 2), so two rounds of *pairwise* affine additions over plain position arithmetic —
// entries (2i, 2i+1) never straddle a bucket — shrink every bucket's list 4x before the XYZZ accumulation:
// 3.5 + 1.75 + 2.5 = 7.75 products per original entry instead of 10.
// One lane handles AFF_B pairs per batch (prefix products in local memory), the 32 lanes of a warp share ONE
// inversion through a shuffle product scan, and that inversion is the binary-Euclid one: shifts and adds on the
// ALU pipe, overlapping the other warps' multiplier work.  (0, 0) is not on the curve and encodes "no point"
// (hole or P + (-P)); P + P takes the tangent.  The group element of every bucket is unchanged.
#ifndef MSM_AFF_B
#define MSM_AFF_B 256
#endif
constexpr int AFF_B = MSM_AFF_B;   // pairs per lane per shared inversion (prefix products: AFF_B field elements of local memory)
constexpr int AFF_THREADS = 128;
#ifndef AFF_MIN_BLOCKS
#define AFF_MIN_BLOCKS 3
#endif

template <class F>
DEV bool aff_none(const Affine<F>& p) { return p.x.is_zero() && p.y.is_zero(); }

// x coordinate only (first half of the affine point): all the forward pass needs unless the two x agree
template <class F>
DEV F load_x_ro(const Affine<F>* p) {
    static_assert(sizeof(F) % 16 == 0, "size");
    F r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(F) / 16); i++) {
        uint4 t = __ldg(q + i);
        w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
    return r;
}

// kind: 0 nothing, 1 first only, 2 second only, 3 chord, 4 tangent; d = the denominator to invert (1 if none).
// FULL = false (forward pass): only d is produced, from the x coordinates alone; the y coordinates (a sign flips
// only y) are fetched just when the x agree.  FULL = true (backward pass): a and b are complete.
template <class F, bool GATHER, bool FULL>
DEV int aff_load_pair(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ idx, size_t i, size_t npairs,
                      Affine<F>& a, Affine<F>& b, F& d) {
    d = F::one();
    if (i >= npairs) return 0;
    bool va, vb;
    const Affine<F>*pa, *pb;
    bool na = false, nb = false;
    if (GATHER) {
        uint2 e = __ldg(reinterpret_cast<const uint2*>(idx) + i);
        va = e.x != 0xffffffffu;
        vb = e.y != 0xffffffffu;
        pa = pts + (e.x & ~msm::DIGIT_NEG);
        pb = pts + (e.y & ~msm::DIGIT_NEG);
        na = (e.x & msm::DIGIT_NEG) != 0;
        nb = (e.y & msm::DIGIT_NEG) != 0;
    } else {
        pa = pts + 2 * i;
        pb = pts + 2 * i + 1;
        va = vb = true;                          // decided from the loaded values below
    }
    if (FULL || !GATHER) {
        // dense inputs carry their validity in the value ((0, 0) = none), so they are always read whole
        if (va) { a = load_pod_ro(pa); if (na) a.y = neg(a.y); }
        if (vb) { b = load_pod_ro(pb); if (nb) b.y = neg(b.y); }
        if (!GATHER) { va = !aff_none(a); vb = !aff_none(b); }
    } else {
        if (va) a.x = load_x_ro(pa);
        if (vb) b.x = load_x_ro(pb);
    }
    if (!va) return vb ? 2 : 0;
    if (!vb) return 1;
    F dx = sub(b.x, a.x);
    if (!dx.is_zero()) { d = dx; return 3; }
    if (!(FULL || !GATHER)) {                    // rare: same x, the y decide between tangent and infinity
        a = load_pod_ro(pa); if (na) a.y = neg(a.y);
        b = load_pod_ro(pb); if (nb) b.y = neg(b.y);
    }
    if (a.y == b.y && !a.y.is_zero()) { d = dbl(a.y); return 4; }
    return 0;                                   // P + (-P) (or a 2-torsion point doubled): the point at infinity
}

// `batch` (<= AFF_B) pairs per lane share one inversion: long batches amortise it, short ones keep every warp of
// the machine busy on small inputs (chosen by the host from the pair count)
template <class F, bool GATHER>
__global__ void __launch_bounds__(AFF_THREADS, sizeof(F) > 48 ? 1 : AFF_MIN_BLOCKS) k_affine_pairs(const Affine<F>* __restrict__ pts,
                                                              const uint32_t* __restrict__ idx, size_t npairs,
                                                              Affine<F>* __restrict__ out, int batch) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    F pre[AFF_B];                               // exclusive prefix products of this lane's denominators
    for (size_t base = warp * (32 * (size_t)batch); base < npairs; base += nwarps * (32 * (size_t)batch)) {
        F run = F::one();
        for (int k = 0; k < batch; k++) {
            Affine<F> a, b;
            F d;
            aff_load_pair<F, GATHER, false>(pts, idx, base + (size_t)k * 32 + lane, npairs, a, b, d);
            pre[k] = run;
            run = mul(run, d);
        }
        // 1 / (this lane's product) from ONE inversion per warp: inclusive prefix and suffix product scans
        F pfx = run, sfx = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            F up = warp_shfl_up(pfx, off), dn = warp_shfl_down(sfx, off);
            if (lane >= (uint32_t)off) pfx = mul(pfx, up);
            if (lane + off < 32) sfx = mul(sfx, dn);
        }
        F total = warp_bcast(pfx, 31);
        F before = warp_shfl_up(pfx, 1), after = warp_shfl_down(sfx, 1);
        F inv_run = inv_euclid(total);          // same data on every lane: no divergence
        if (lane > 0) inv_run = mul(inv_run, before);
        if (lane < 31) inv_run = mul(inv_run, after);
        for (int k = batch - 1; k >= 0; k--) {
            const size_t i = base + (size_t)k * 32 + lane;
            Affine<F> a, b, r;
            F d;
            int kind = aff_load_pair<F, GATHER, true>(pts, idx, i, npairs, a, b, d);
            F dinv = mul(inv_run, pre[k]);
            inv_run = mul(inv_run, d);
            if (kind >= 3) {
                F lam = kind == 3 ? mul(sub(b.y, a.y), dinv) : mul(add(dbl(sqr(a.x)), sqr(a.x)), dinv);
                F x3 = sub(sub(sqr(lam), a.x), kind == 3 ? b.x : a.x);
                r.y = sub(mul(lam, sub(a.x, x3)), a.y);
                r.x = x3;
            } else if (kind == 1) {
                r = a;
            } else if (kind == 2) {
                r = b;
            } else {
                r.x = F::zero();
                r.y = F::zero();
            }
            if (i < npairs) store_pod(out + i, r);
        }
    }
}

// buckets split into 2..SMALL_MULTI_MAX tasks: one thread joins them
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_finalize_small(const uint32_t* __restrict__ small_list,
                                                                const uint32_t* __restrict__ counters,
                                                                const uint32_t* __restrict__ bucket_size,
                                                                const uint32_t* __restrict__ task_start,
                                                                uint32_t task_len, const XYZZ<F>* __restrict__ partials,
                                                                XYZZ<F>* __restrict__ buckets, uint32_t merge) {
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
        if (merge) {
            XYZZ<F> q = load_pod(buckets + g);
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
                                                              XYZZ<F>* __restrict__ buckets, uint32_t merge) {
    extern __shared__ uint4 sm_raw[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(sm_raw);
    uint32_t count = counters[3];
    for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
        uint32_t g = big_list[i];
        uint32_t nt = (bucket_size[g] + task_len - 1) / task_len, ts = task_start[g];
        XYZZ<F> acc = XYZZ<F>::infinity();
        if (merge && threadIdx.x == 0) acc = load_pod(buckets + g);
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

// The same with one WARP per chunk (ec.cuh, warp-cooperative group law): for small MSMs, where the chunk count
// cannot fill the machine anyway and the running-sum chain is pure latency (5 product latencies per addition
// instead of 14, 3 per doubling instead of 9).
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_bucket_reduce_warp(const XYZZ<F>* __restrict__ buckets, uint32_t nwin,
                                                                    uint32_t nb, uint32_t m, uint32_t T,
                                                                    XYZZ<F>* __restrict__ chunk_res) {
    uint32_t id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (id >= nwin * T) return;                    // whole warps leave together
    uint32_t w = id / T, t = id % T;
    const XYZZ<F>* b = buckets + (size_t)w * nb + (size_t)t * m;
    XYZZ<F> running = XYZZ<F>::infinity(), acc = XYZZ<F>::infinity();
    for (uint32_t j = m; j-- > 0;) {
        XYZZ<F> q = load_pod(b + j);
        xyzz_add_warp(running, q);
        xyzz_add_warp(acc, running);
    }
    if (t) {
        XYZZ<F> s = xyzz_mul_small_warp(running, (uint64_t)t * m);
        xyzz_add_warp(acc, s);
    }
    if ((threadIdx.x & 31) == 0) store_pod(chunk_res + id, acc);
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

// Σ_w 2^(c·w)·S_w, high to low (variable_base.rs:92-105): one warp, every lane holds the accumulator and
// the independent products of each doubling / addition run on different lanes (ec.cuh, warp-cooperative
// group law): 3 product latencies per doubling instead of 9.
template <class F>
__global__ void __launch_bounds__(32) k_horner(const XYZZ<F>* __restrict__ window_sum, uint32_t nwin, uint32_t c,
                                               XYZZ<F>* __restrict__ out) {
    XYZZ<F> acc = load_pod(window_sum + nwin - 1);
    for (uint32_t w = nwin - 1; w-- > 0;) {
        for (uint32_t k = 0; k < c; k++) xyzz_dbl_warp(acc);
        XYZZ<F> q = load_pod(window_sum + w);
        xyzz_add_warp(acc, q);
    }
    if (threadIdx.x == 0) store_pod(out, acc);
}

// ---- result emission ----------------------------------------------------------------------------------
// mode 0: affine x|y (2 F) followed by one 32-bit infinity flag;  mode 1: Jacobian x|y|z (3 F).
// One warp: the additions are lane-parallel, the inversion is the binary-Euclid one (same data on every
// lane, so no divergence), ~20x shorter than the Fermat ladder on a single dependent chain.
template <class F>
__global__ void __launch_bounds__(32) k_emit(const XYZZ<F>* __restrict__ in, uint32_t count, uint32_t mode,
                                             uint32_t* __restrict__ out) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t i = 0; i < count; i++) {
        XYZZ<F> q = load_pod(in + i);
        xyzz_add_warp(acc, q);
    }
    if (threadIdx.x) return;
    constexpr int N = sizeof(F) / 4;
    if (mode == 0) {
        F x, y;
        bool finite = xyzz_to_affine_tail(acc, x, y);
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

// The same round as two kernels.  The product (forward) pass is memory-bound — two random x gathers and one product
// per pair — and the addition (backward) pass multiplier-bound; fused in one kernel the first costs a third of the
// round.  Split, the forward kernel of segment j+1 runs on a second stream underneath the backward kernel of segment
// j (MsmJob::affine_round): its prefix products and per-lane totals travel through global memory.
constexpr int AFF_FWD_THREADS = 64;

template <class F, bool GATHER>
__global__ void __launch_bounds__(AFF_FWD_THREADS) k_affine_fwd(const Affine<F>* __restrict__ pts,
                                                                const uint32_t* __restrict__ idx, size_t npairs, int batch,
                                                                size_t wb_lo, size_t wb_hi, F* __restrict__ pre,
                                                                F* __restrict__ totals) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t wb = wb_lo + warp; wb < wb_hi; wb += nwarps) {
        const size_t base = wb * (32 * (size_t)batch);
        F run = F::one();
        for (int k = 0; k < batch; k++) {
            Affine<F> a, b;
            F d;
            const size_t i = base + (size_t)k * 32 + lane;
            aff_load_pair<F, GATHER, false>(pts, idx, i, npairs, a, b, d);
            store_pod(pre + i, run);
            run = mul(run, d);
        }
        store_pod(totals + wb * 32 + lane, run);
    }
}

template <class F, bool GATHER>
__global__ void __launch_bounds__(AFF_THREADS, sizeof(F) > 48 ? 1 : AFF_MIN_BLOCKS) k_affine_bwd(
    const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ idx, size_t npairs, int batch, size_t wb_lo, size_t wb_hi,
    const F* __restrict__ pre, const F* __restrict__ totals, Affine<F>* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t wb = wb_lo + warp; wb < wb_hi; wb += nwarps) {
        const size_t base = wb * (32 * (size_t)batch);
        F run = load_pod(totals + wb * 32 + lane);
        F pfx = run, sfx = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            F up = warp_shfl_up(pfx, off), dn = warp_shfl_down(sfx, off);
            if (lane >= (uint32_t)off) pfx = mul(pfx, up);
            if (lane + off < 32) sfx = mul(sfx, dn);
        }
        F total = warp_bcast(pfx, 31);
        F before = warp_shfl_up(pfx, 1), after = warp_shfl_down(sfx, 1);
        F inv_run = inv_euclid(total);
        if (lane > 0) inv_run = mul(inv_run, before);
        if (lane < 31) inv_run = mul(inv_run, after);
        for (int k = batch - 1; k >= 0; k--) {
            const size_t i = base + (size_t)k * 32 + lane;
            Affine<F> a, b, r;
            F d;
            int kind = aff_load_pair<F, GATHER, true>(pts, idx, i, npairs, a, b, d);
            F dinv = mul(inv_run, load_pod(pre + i));
            inv_run = mul(inv_run, d);
            if (kind >= 3) {
                F lam = kind == 3 ? mul(sub(b.y, a.y), dinv) : mul(add(dbl(sqr(a.x)), sqr(a.x)), dinv);
                F x3 = sub(sub(sqr(lam), a.x), kind == 3 ? b.x : a.x);
                r.y = sub(mul(lam, sub(a.x, x3)), a.y);
                r.x = x3;
            } else if (kind == 1) {
                r = a;
            } else if (kind == 2) {
                r = b;
            } else {
                r.x = F::zero();
                r.y = F::zero();
            }
            if (i < npairs) store_pod(out + i, r);
        }
    }
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

// One MSM in flight: begin() sizes the scratch for chunks of <= cap points, chunk() sorts and accumulates
// one point range into the shared buckets, finish() reduces them to one XYZZ point.
template <class F>
struct MsmJob {
    Plan p;
    cudaStream_t s = nullptr;
    const DeviceInfo* dev = nullptr;
    bool use_table = false;
    TableRef tbl;
    size_t done = 0;             // points consumed so far
    uint32_t nchunks = 0;
    Scratch s_digits, s_sorted, s_pairs, s_hist, s_part, s_bstart, s_bsize, s_tstart, s_bsums, s_tasks, s_small, s_big,
        s_cnt, s_buckets, s_partials, s_chunk, s_wpart, s_wsum;
    uint32_t *digits, *sorted, *hist, *part_start, *bstart, *bsize, *tstart, *bsums, *small_list, *big_list, *counters;
    uint2* pairs;
    uint4* tasks;
    XYZZ<F>*buckets, *partials, *chunk_res, *wpart, *wsum;
    Scratch s_q1, s_q2, s_pre, s_tot;
    Affine<F>*q1 = nullptr, *q2 = nullptr;       // affine pre-reduction: pair sums, sums of four
    F *aff_pre = nullptr, *aff_tot = nullptr;    // split rounds: prefix products per pair, products per lane
    cudaStream_t s2 = nullptr;                   // producer stream of the split rounds
    std::vector<cudaEvent_t> events;
    bool aff_split = false;
    ~MsmJob() {
        for (cudaEvent_t e : events) cudaEventDestroy(e);      // released once the recorded work has completed
    }
    int32_t new_event(cudaEvent_t* e) {
        MPC_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        events.push_back(*e);
        return MPC_CUDA_OK;
    }
    size_t nbuckets = 0;
    int acc_blocks = 1, acc_blocks_d = 1, aff_blocks_g = 1, aff_blocks_d = 1;

    int32_t begin(size_t n, size_t cap, cudaStream_t stream, const TableRef* t) {
        s = stream;
        dev = current_device_info();
        use_table = t && t->table;
        if (use_table) tbl = *t;
        MPC_ARG_CHECK(n < ((size_t)1 << 31) && cap >= 1 && cap <= n);
        p = make_plan(n, cap, dev->sm_count, use_table ? tbl.c : 0);
        nbuckets = (size_t)p.snwin * p.nb;
        MPC_ARG_CHECK(p.snp < ((size_t)1 << 31) && (!use_table || (size_t)p.nwin * tbl.stride < ((size_t)1 << 31)));
        MPC_ARG_CHECK(nbuckets < ((size_t)1 << 31) && (size_t)p.nwin * cap / p.task_len + nbuckets < ((size_t)1 << 32));
        const uint32_t hbins = 1u << p.hi_bits;
        MPC_TRY(s_digits.alloc(&digits, (size_t)p.nwin * cap, s));
        MPC_TRY(s_sorted.alloc(&sorted, p.aff ? (size_t)p.snwin * p.snp : (size_t)p.nwin * cap, s));
        if (p.aff) {
            MPC_TRY(s_q1.alloc(&q1, (size_t)p.snwin * p.snp / 2, s));
            if (p.aff > 1) MPC_TRY(s_q2.alloc(&q2, (size_t)p.snwin * p.snp / 4, s));
            MPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&aff_blocks_g, k_affine_pairs<F, true>, AFF_THREADS, 0));
            MPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&aff_blocks_d, k_affine_pairs<F, false>, AFF_THREADS, 0));
            if (aff_blocks_g < 1) aff_blocks_g = 1;
            if (aff_blocks_d < 1) aff_blocks_d = 1;
            // opt-in two-kernel pipeline: needs the prefix products in global memory. Measured SLOWER than the fused
            // kernel (2^24 points, c = 22: 83.9 vs 78.5 ms; 2^22: 28.1 vs 25.0 ms - the prefix products' round trip
            // through HBM costs more than the overlap wins), so the fused kernel stays the default.
            const size_t pairs0 = (size_t)p.snwin * p.snp / 2;
            aff_split = g_opt_msm_affine_split.load(std::memory_order_relaxed) == 1;
            if (aff_split) {
                const size_t span = 32 * (size_t)AFF_B;
                MPC_TRY(s_pre.alloc(&aff_pre, (pairs0 + span - 1) / span * span + span, s));
                MPC_TRY(s_tot.alloc(&aff_tot, pairs0 / 16 + 64, s));        // >= 32 lanes per batch of >= 16 pairs per lane
                MPC_TRY(aux_stream(&s2));
            }
        }
        MPC_TRY(s_pairs.alloc(&pairs, (size_t)p.nwin * cap, s));
        MPC_TRY(s_hist.alloc(&hist, (size_t)p.snwin * p.chunks * hbins, s));
        MPC_TRY(s_part.alloc(&part_start, (size_t)p.snwin * (hbins + 1), s));
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
        MPC_TRY(s_chunk.alloc(&chunk_res, (size_t)p.snwin * p.red_t, s));
        MPC_TRY(s_wpart.alloc(&wpart, (size_t)p.snwin * p.sum_parts, s));
        MPC_TRY(s_wsum.alloc(&wsum, p.snwin, s));
        MPC_CUDA_TRY(cudaMemsetAsync(buckets, 0, nbuckets * sizeof(XYZZ<F>), s));      // all-zero = infinity
        MPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_blocks_d, k_accumulate<F, true>, ACC_THREADS, 0));
        MPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_blocks, k_accumulate<F, false>, ACC_THREADS, 0));
        if (acc_blocks < 1) acc_blocks = 1;
        if (acc_blocks_d < 1) acc_blocks_d = 1;
        MPC_TRY(allow_smem(k_finalize_big<F>, ACC_THREADS * sizeof(XYZZ<F>)));
        MPC_TRY(allow_smem(k_window_sum<F>, ACC_THREADS * sizeof(XYZZ<F>)));
        return MPC_CUDA_OK;
    }

    // bases / inf / scalars: device pointers to this chunk's `len` points (bases unused with a table)
    int32_t chunk(const Affine<F>* bases, const uint8_t* inf, const Fr* scalars, size_t len) {
        if (len == 0) return MPC_CUDA_OK;
        MPC_ARG_CHECK(len <= p.cap && done + len <= p.n);
        const uint32_t hbins = 1u << p.hi_bits, lbins = 1u << p.lo_bits;
        const size_t sn = use_table ? len * p.nwin : len;
        const size_t chunk_len = (sn + p.chunks - 1) / p.chunks;
        const uint32_t merge = nchunks ? 1u : 0u;
        // pre-reduction rounds of THIS chunk: the plan's (sized for the longest chunk), fewer for a short chunk
        // whose buckets are too thin for the padding and the shared inversions to pay (same rule as the plan's)
        uint32_t aff = p.aff;
        if (aff && p.aff_auto) {
            const size_t entries = len * (size_t)p.nwin;
            const uint32_t by_size = entries < 8 * nbuckets ? 0 : entries >= ((size_t)3 << 25) ? 2 : entries >= ((size_t)3 << 24) ? 1 : 0;
            if (by_size < aff) aff = by_size;
        }
        profile_begin("msm_sort", s);
        // digits[w*len + i]: with a table the flat array IS one window of nwin*len entries
        k_digits<<<grid_for(len, 256, 8), 256, 0, s>>>(scalars, inf, len, p.c, p.nwin, digits);
        MPC_KERNEL_CHECK();
        dim3 grid1(p.chunks, p.snwin);
        k_hist1<<<grid1, SORT_THREADS, hbins * sizeof(uint32_t), s>>>(digits, sn, p.lo_bits, hbins, chunk_len, hist);
        MPC_KERNEL_CHECK();
        k_scan1<<<p.snwin, 1024, 0, s>>>(hist, p.chunks, hbins, part_start);
        MPC_KERNEL_CHECK();
        k_scatter1<<<grid1, SORT_THREADS, hbins * sizeof(uint32_t), s>>>(digits, sn, p.lo_bits, hbins, chunk_len, hist,
                                                                        pairs, len, use_table ? tbl.stride : 0,
                                                                        use_table ? tbl.offset + done : 0);
        MPC_KERNEL_CHECK();
        if (aff) MPC_CUDA_TRY(cudaMemsetAsync(sorted, 0xff, (size_t)p.snwin * p.snp * sizeof(uint32_t), s));   // holes
        k_sort2<<<dim3(hbins, p.snwin), SORT2_THREADS, (lbins + SORT2_THREADS) * sizeof(uint32_t), s>>>(
            pairs, part_start, sn, p.lo_bits, hbins, sorted, bstart, bsize, aff, aff ? p.snp : sn);
        MPC_KERNEL_CHECK();
        k_task_sums<<<p.scan_blocks, 1024, 0, s>>>(bsize, nbuckets, p.task_len, bsums);
        MPC_KERNEL_CHECK();
        k_task_offsets<<<1, 1024, 0, s>>>(bsums, p.scan_blocks, counters);
        MPC_KERNEL_CHECK();
        k_build_tasks<<<p.scan_blocks, 1024, 0, s>>>(bstart, bsize, bsums, nbuckets, p.task_len, tasks, tstart,
                                                     small_list, big_list, counters);
        MPC_KERNEL_CHECK();
        profile_end("msm_sort", s);

        profile_begin("msm_accumulate", s);
        const Affine<F>* points = use_table ? (const Affine<F>*)tbl.table : bases;
        // grids sized by the work of THIS chunk (an upper bound of its task count), not by the machine: the MSMs of
        // one proof and of the other parties run side by side on their own streams, and a full-machine grid of
        // mostly idle CTAs in front of them would serialise the small ones
        const size_t task_bound = ((sn * p.snwin) >> aff) / p.task_len + nbuckets + 1;
        const unsigned acc_grid = (unsigned)std::min<size_t>((size_t)dev->sm_count * (aff ? acc_blocks_d : acc_blocks),
                                                             (task_bound + ACC_THREADS - 1) / ACC_THREADS);
        const unsigned fin_small = (unsigned)std::min<size_t>((size_t)dev->sm_count * 2, (nbuckets + ACC_THREADS - 1) / ACC_THREADS);
        const unsigned fin_big = (unsigned)std::min<size_t>((size_t)dev->sm_count, nbuckets);
        if (aff) {
            // pairwise affine additions inside every bucket's padded run, then XYZZ accumulation of what is left
            const size_t slots = (size_t)p.snwin * p.snp;
            MPC_TRY((affine_round<true>(points, sorted, slots / 2, q1, aff_blocks_g)));
            if (aff > 1) MPC_TRY((affine_round<false>(q1, nullptr, slots / 4, q2, aff_blocks_d)));
            k_accumulate<F, true><<<acc_grid, ACC_THREADS, 0, s>>>(
                aff > 1 ? q2 : q1, nullptr, p.snp >> aff, p.nb, tasks, counters, buckets, partials, merge);
        } else {
            k_accumulate<F, false><<<acc_grid, ACC_THREADS, 0, s>>>(points, sorted, sn, p.nb, tasks, counters, buckets, partials,
                                                                    merge);
        }
        MPC_KERNEL_CHECK();
        profile_end("msm_accumulate", s);

        profile_begin("msm_reduce", s);
        k_finalize_small<F><<<fin_small, ACC_THREADS, 0, s>>>(small_list, counters, bsize, tstart, p.task_len, partials, buckets,
                                                              merge);
        MPC_KERNEL_CHECK();
        k_finalize_big<F><<<fin_big, ACC_THREADS, ACC_THREADS * sizeof(XYZZ<F>), s>>>(
            big_list, counters, bsize, tstart, p.task_len, partials, buckets, merge);
        MPC_KERNEL_CHECK();
        profile_end("msm_reduce", s);
        done += len;
        nchunks++;
        return MPC_CUDA_OK;
    }

    // one pre-reduction round over `npairs` pairs: the fused kernel, or forward / backward kernels pipelined over
    // two streams in segments of whole warp batches
    template <bool GATHER>
    int32_t affine_round(const Affine<F>* pts, const uint32_t* idx, size_t npairs, Affine<F>* out, int blocks) {
        // pairs per lane per shared inversion: at least ~4 batches for every resident warp
        const size_t lanes = (size_t)dev->sm_count * blocks * AFF_THREADS;
        size_t b = npairs / (lanes * 4);
        const int batch = (int)(b < 16 ? 16 : b > (size_t)AFF_B ? (size_t)AFF_B : b);
        const size_t nwb = (npairs + 32 * (size_t)batch - 1) / (32 * (size_t)batch);       // warp batches
        const int segments = !aff_split ? 0 : nwb >= 4096 ? 8 : nwb >= 8 ? 4 : 1;
        if (segments == 0) {
            k_affine_pairs<F, GATHER><<<dev->sm_count * blocks, AFF_THREADS, 0, s>>>(pts, idx, npairs, out, batch);
            MPC_KERNEL_CHECK();
            return MPC_CUDA_OK;
        }
        // the producer stream may start once everything enqueued so far on s (sort, previous round's readers of
        // aff_pre / aff_tot) is done
        cudaEvent_t ready;
        MPC_TRY(new_event(&ready));
        MPC_CUDA_TRY(cudaEventRecord(ready, s));
        MPC_CUDA_TRY(cudaStreamWaitEvent(s2, ready, 0));
        const unsigned fwd_grid = (unsigned)std::min<size_t>((size_t)dev->sm_count * 4, (nwb * 32 + AFF_FWD_THREADS - 1) / AFF_FWD_THREADS);
        const unsigned bwd_grid = (unsigned)std::min<size_t>((size_t)dev->sm_count * blocks, (nwb * 32 + AFF_THREADS - 1) / AFF_THREADS);
        for (int c = 0; c < segments; c++) {
            const size_t lo = nwb * c / segments, hi = nwb * (c + 1) / segments;
            if (lo == hi) continue;
            k_affine_fwd<F, GATHER><<<fwd_grid, AFF_FWD_THREADS, 0, s2>>>(pts, idx, npairs, batch, lo, hi, aff_pre, aff_tot);
            MPC_KERNEL_CHECK();
            cudaEvent_t done;
            MPC_TRY(new_event(&done));
            MPC_CUDA_TRY(cudaEventRecord(done, s2));
            MPC_CUDA_TRY(cudaStreamWaitEvent(s, done, 0));
            k_affine_bwd<F, GATHER><<<bwd_grid, AFF_THREADS, 0, s>>>(pts, idx, npairs, batch, lo, hi, aff_pre, aff_tot, out);
            MPC_KERNEL_CHECK();
        }
        return MPC_CUDA_OK;
    }

    int32_t finish(XYZZ<F>* result) {
        profile_begin("msm_reduce", s);
        const size_t tree_smem = ACC_THREADS * sizeof(XYZZ<F>);
        uint32_t red_threads = p.snwin * p.red_t;
        const int64_t warp_opt = g_opt_msm_reduce_warp_max.load(std::memory_order_relaxed);
        // one warp per chunk (32x the threads) only while that stays a sliver of the machine: the MSMs of one proof and of
        // the other parties run concurrently.  3-party prove, 2^13 / 2^15 SPDZ: 6.26 / 13.3 ms with a limit of 128 chunks,
        // 6.19 / 12.8 at 512, 5.8 / 16.8 at 8192
        if (red_threads <= (uint32_t)(warp_opt > 0 ? warp_opt : 512)) {
            k_bucket_reduce_warp<F><<<(red_threads * 32 + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, s>>>(
                buckets, p.snwin, p.nb, p.red_m, p.red_t, chunk_res);
        } else {
            k_bucket_reduce<F><<<(red_threads + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, s>>>(buckets, p.snwin, p.nb,
                                                                                                    p.red_m, p.red_t, chunk_res);
        }
        MPC_KERNEL_CHECK();
        if (p.sum_parts == 1) {
            k_window_sum<F><<<dim3(1, p.snwin), ACC_THREADS, tree_smem, s>>>(chunk_res, p.red_t, p.red_t, wsum);
            MPC_KERNEL_CHECK();
        } else {
            k_window_sum<F><<<dim3(p.sum_parts, p.snwin), ACC_THREADS, tree_smem, s>>>(chunk_res, p.red_t, 1024, wpart);
            MPC_KERNEL_CHECK();
            k_window_sum<F><<<dim3(1, p.snwin), ACC_THREADS, tree_smem, s>>>(wpart, p.sum_parts, p.sum_parts, wsum);
            MPC_KERNEL_CHECK();
        }
        k_horner<F><<<1, 32, 0, s>>>(wsum, p.snwin, p.c, result);
        MPC_KERNEL_CHECK();
        profile_end("msm_reduce", s);
        return MPC_CUDA_OK;
    }
};

// result (one XYZZ on the device) = Σ scalars[i] * bases[i], everything already resident
template <class F>
int32_t msm_run(const Affine<F>* bases, const uint8_t* inf, const Fr* scalars, size_t n, XYZZ<F>* result,
                cudaStream_t s, const TableRef* tbl = nullptr) {
    if (n == 0) {
        MPC_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(XYZZ<F>), s));
        return MPC_CUDA_OK;
    }
    ProfileScope prof_total("msm_total", s);
    MsmJob<F> job;
    MPC_TRY(job.begin(n, n, s, tbl));
    MPC_TRY(job.chunk(bases, inf, scalars, n));
    return job.finish(result);
}

// ---- registered base vectors --------------------------------------------------------------------------
template <class F>
int32_t register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(handle && (n == 0 || bases_xy));
    BaseRef v = std::make_shared<BaseVec>();
    v->n = n;
    v->g2 = g2;
    v->dev_index = current_device_index();
    v->cuda_device = current_device_info()->cuda_device;
    MPC_CUDA_TRY(cudaMalloc(&v->bases, (n ? n : 1) * sizeof(Affine<F>)));
    MPC_CUDA_TRY(cudaMemcpyAsync(v->bases, bases_xy, n * sizeof(Affine<F>), cudaMemcpyHostToDevice, s));
    if (inf) {
        bool any = false;
        for (size_t i = 0; i < n && !any; i++) any = inf[i] != 0;
        if (any) {
            MPC_CUDA_TRY(cudaMalloc((void**)&v->inf, n));
            MPC_CUDA_TRY(cudaMemcpyAsync(v->inf, inf, n, cudaMemcpyHostToDevice, s));
        }
    }
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    *handle = registry_add(v);
    return MPC_CUDA_OK;
}

// the vector cut into `parts` point ranges, range k resident on device k mod n_dev of the init list
template <class F>
int32_t register_bases_sharded(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint32_t parts, uint64_t* handle,
                               bool g2) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(handle && (n == 0 || bases_xy) && parts >= 1 && parts <= 64);
    const int n_dev = device_list_size();
    BaseRef parent = std::make_shared<BaseVec>();
    parent->n = n;
    parent->g2 = g2;
    parent->owned = false;
    for (uint32_t k = 0; k < parts; k++) {
        size_t lo = n / parts * k + std::min<size_t>(k, n % parts);
        size_t len = n / parts + (k < n % parts ? 1 : 0);
        DeviceScope scope((int)(k % n_dev));      // more parts than devices: round robin (single-GPU tests)
        MPC_TRY(scope.rc);
        uint64_t h = 0;
        MPC_TRY(register_bases<F>(bases_xy ? bases_xy + lo * (sizeof(Affine<F>) / 8) : nullptr, inf ? inf + lo : nullptr, len,
                                  &h, g2));
        BaseSnap snap;
        MPC_TRY(registry_find(h, g2, 0, 0, &snap));
        parent->parts.push_back(snap.ref);
        parent->lo.push_back(lo);
        MPC_TRY(mpc_cuda_msm_release_bases(h));      // the parent keeps the only reference
    }
    parent->lo.push_back(n);
    *handle = registry_add(parent);
    return MPC_CUDA_OK;
}

inline TableRef table_of(const BaseSnap& v, size_t offset) {
    TableRef t;
    t.table = v.table;
    t.c = v.table_c;
    t.stride = v.n;
    t.offset = offset;
    return t;
}

// ---- window table T[w][i] = 2^(c*w) * P_i in affine form -------------------------------------------------
// One thread owns PRE_K consecutive bases.  Pass 1 walks each base's doubling chain in Jacobian coordinates
// (2M + 5S per doubling), leaving X, Y in the table slot and Z in scratch, and multiplies the Z's into a
// running prefix product kept in scratch; ONE field inversion per thread (Montgomery's trick,
// ff/src/fields/mod.rs:597-660 batch_inversion) then normalises all PRE_K * (nwin - 1) entries on the way
// back.  The cost that remains is the 253 doublings per base.
constexpr int PRE_K = 4;

template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_precompute(const Affine<F>* __restrict__ bases,
                                                            const uint8_t* __restrict__ inf, size_t n, size_t first,
                                                            size_t count, uint32_t c, uint32_t nwin,
                                                            Affine<F>* __restrict__ table,
                                                            F* __restrict__ zs /* [nwin-1][count] */,
                                                            F* __restrict__ prefix /* [nwin-1][count] */) {
    // this launch covers bases [first, first + count); the scratch is indexed relative to `first`
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i0 = first + t * PRE_K;
    if (i0 >= first + count) return;
    size_t i1 = i0 + PRE_K < first + count ? i0 + PRE_K : first + count;
    F run = F::one();
    for (size_t i = i0; i < i1; i++) {
        Affine<F> pt = load_pod_ro(bases + i);
        store_pod(table + i, pt);
        if (inf && inf[i]) continue;               // never referenced: its digits are zero
        Jac<F> acc;
        acc.x = pt.x; acc.y = pt.y; acc.z = F::one();
        for (uint32_t w = 1; w < nwin; w++) {
            for (uint32_t k = 0; k < c; k++) jac_dbl(acc);
            Affine<F> raw;
            raw.x = acc.x; raw.y = acc.y;
            size_t slot = (size_t)(w - 1) * count + (i - first);
            store_pod(table + (size_t)w * n + i, raw);
            store_pod(zs + slot, acc.z);
            store_pod(prefix + slot, run);          // product of every Z before this one
            run = mul(run, acc.z);                  // Z != 0: the points have odd prime order
        }
    }
    F invrun = inv(run);
    for (size_t i = i1; i-- > i0;) {
        if (inf && inf[i]) continue;
        for (uint32_t w = nwin - 1; w >= 1; w--) {
            size_t slot = (size_t)(w - 1) * count + (i - first);
            F z = load_pod(zs + slot), pre = load_pod(prefix + slot);
            F zi = mul(invrun, pre);                // 1 / Z of this entry
            invrun = mul(invrun, z);
            F zi2 = sqr(zi);
            Affine<F> e = load_pod(table + (size_t)w * n + i);
            e.x = mul(e.x, zi2);
            e.y = mul(e.y, mul(zi2, zi));
            store_pod(table + (size_t)w * n + i, e);
        }
    }
}

template <class F>
int32_t precompute_one(const BaseSnap& v, uint64_t handle, uint32_t window_bits, cudaStream_t s) {
    uint32_t c = window_bits;
    if (c == 0) {
        uint32_t l = log2_ceil(v.n ? v.n : 1);
        // small vectors are latency-bound; what counts for a proof is several of them running side by side (4 MSMs x
        // 3 parties): tools/tune_msm.py + bench.py extra.prove with PK_BITS_G1/G2, domain 2^13: c = 12 for both groups
        // gives the shortest 3-party prove (10.1 ms; c = 10: 11.9 ms), single MSMs 1.1 ms (G1) / 3.4 ms (G2)
        const uint32_t small_c = 12;
        // 2^24 with the affine pre-reduction: c = 22 -> 78.3 ms, c = 23 -> 80.0 ms (both 12 windows; half the buckets)
        c = l >= 22 ? 22 : l >= 20 ? 20 : l >= 18 ? 17 : l >= 16 ? 16 : l >= 11 ? small_c : l > 7 ? l - 3 : 4;
        if (msm::SCALAR_BITS % c == 1) c--;       // a one-bit top window would put n/2 entries in one bucket
    }
    uint32_t nwin = msm::num_windows(c);
    MPC_ARG_CHECK((size_t)nwin * v.n < ((size_t)1 << 31));
    void* table = nullptr;
    MPC_CUDA_TRY(cudaMalloc(&table, (size_t)nwin * (v.n ? v.n : 1) * sizeof(Affine<F>)));
    if (v.n) {
        // slices of 2^20 bases: the scratch (2 x (nwin - 1) field elements per base) stays ~1 GB however long the CRS
        const size_t slice = std::min<size_t>(v.n, (size_t)1 << 20);
        Scratch sz, sp;
        F *zs, *prefix;
        if (sz.alloc(&zs, (size_t)(nwin - 1) * slice, s) != MPC_CUDA_OK || sp.alloc(&prefix, (size_t)(nwin - 1) * slice, s) != MPC_CUDA_OK) {
            cudaFree(table);
            return MPC_CUDA_ERR_CUDA;
        }
        ProfileScope prof("msm_precompute", s);
        for (size_t first = 0; first < v.n; first += slice) {
            size_t count = std::min(slice, v.n - first);
            size_t threads = (count + PRE_K - 1) / PRE_K;
            k_precompute<F><<<(unsigned)((threads + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, s>>>(
                (const Affine<F>*)v.bases, v.inf, v.n, first, count, c, nwin, (Affine<F>*)table, zs, prefix);
            MPC_KERNEL_CHECK();
        }
    }
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    if (handle) {
        int32_t rc = registry_set_table(handle, table, c);
        if (rc != MPC_CUDA_OK) cudaFree(table);
        return rc;
    }
    // a part of a sharded vector: reachable only through its parent, which the caller holds
    BaseVec& bv = *v.ref;
    if (bv.table) bv.retired.push_back(bv.table);
    bv.table = table;
    bv.table_c = c;
    return MPC_CUDA_OK;
}

template <class F>
int32_t precompute(uint64_t handle, uint32_t window_bits, bool g2) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(window_bits == 0 || (window_bits >= 3 && window_bits <= MAX_WINDOW_BITS));
    BaseSnap v;
    MPC_TRY(registry_find(handle, g2, 0, 0, &v));
    if (v.ref->parts.empty()) return precompute_one<F>(v, handle, window_bits, s);
    // every part picks its window width for its own length unless the caller fixed one
    for (size_t k = 0; k < v.ref->parts.size(); k++) {
        DeviceScope scope(v.ref->parts[k]->dev_index);
        MPC_TRY(scope.rc);
        MPC_TRY(precompute_one<F>(snapshot_of(v.ref->parts[k]), 0, window_bits, scope.s));
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

// Host-buffer MSM (AffineMsm::msm as the Rust shim calls it): the scalars and the bases cross PCIe inside
// the call.  Large inputs are streamed as point-range chunks on a second stream so the copy of chunk j+1
// overlaps the sort and bucket accumulation of chunk j; all chunks add into one bucket set.
struct CopyLane {
    static constexpr int MAX_CHUNKS = 16;
    cudaStream_t s = nullptr;
    cudaEvent_t ev[MAX_CHUNKS + 1] = {};
    ~CopyLane() {
        if (s) cudaStreamSynchronize(s);             // error paths: no copy may outlive the buffers it writes
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
        if (s) cudaStreamDestroy(s);
    }
};

inline uint32_t host_chunks_for(size_t n) {
    int64_t k = g_opt_msm_host_chunks.load(std::memory_order_relaxed);
    if (k <= 0) k = n >= ((size_t)1 << 23) ? 6 : n >= ((size_t)1 << 21) ? 4 : n >= ((size_t)1 << 19) ? 2 : 1;
    if (k > 16) k = 16;
    if ((size_t)k > n) k = (int64_t)(n ? n : 1);
    return (uint32_t)k;
}

template <class F>
int32_t msm_host(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                 uint8_t* out_inf) {
    constexpr int N = sizeof(F) / 4;
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out_xy && out_inf && (n == 0 || (bases_xy && scalars)));
    Scratch sb, si, ss, s_res, s_out;
    Affine<F>* db;
    uint8_t* di = nullptr;
    Fr* dsc;
    XYZZ<F>* res;
    uint32_t* out;
    MPC_TRY(sb.alloc(&db, n, s));
    MPC_TRY(ss.alloc(&dsc, n, s));
    if (inf && n) MPC_TRY(si.alloc(&di, n, s));
    MPC_TRY(s_res.alloc(&res, 1, s));
    MPC_TRY(s_out.alloc(&out, 3 * N + 4, s));
    CopyLane lane;                                   // destroyed after the final synchronisation below
    if (n == 0) {
        MPC_CUDA_TRY(cudaMemsetAsync(res, 0, sizeof(XYZZ<F>), s));
    } else {
        // Chunk sizes double towards the middle and halve again (1 2 4 4 2 1): when the kernels are the slower side
        // (one GPU per host) only the first, small copy is exposed; when the copies are (eight GPUs sharing the host's
        // memory and PCIe switches) only the last, small chunk's kernels are; the long middle chunks keep the buckets
        // dense enough for the pre-reduction
        const uint32_t K = host_chunks_for(n);
        size_t bound[CopyLane::MAX_CHUNKS + 1];
        uint64_t wsum = 0, wacc = 0;
        auto weight = [K](uint32_t k) { return (uint64_t)1 << std::min(k, K - 1 - k); };
        for (uint32_t k = 0; k < K; k++) wsum += weight(k);
        bound[0] = 0;
        for (uint32_t k = 1; k <= K; k++) {
            wacc += weight(k - 1);
            bound[k] = k == K ? n : (size_t)((unsigned __int128)n * wacc / wsum);
        }
        size_t cap = 1;
        for (uint32_t k = 0; k < K; k++) cap = std::max(cap, bound[k + 1] - bound[k]);
        MPC_CUDA_TRY(cudaStreamCreateWithFlags(&lane.s, cudaStreamNonBlocking));
        // the buffers were allocated stream-ordered on s: the copy stream may touch them only after that point
        MPC_CUDA_TRY(cudaEventCreateWithFlags(&lane.ev[CopyLane::MAX_CHUNKS], cudaEventDisableTiming));
        MPC_CUDA_TRY(cudaEventRecord(lane.ev[CopyLane::MAX_CHUNKS], s));
        MPC_CUDA_TRY(cudaStreamWaitEvent(lane.s, lane.ev[CopyLane::MAX_CHUNKS], 0));
        ProfileScope prof_total("msm_total", s);
        MsmJob<F> job;
        MPC_TRY(job.begin(n, cap, s, nullptr));
        for (uint32_t k = 0; k < K; k++) {
            const size_t lo = bound[k], len = bound[k + 1] - lo;
            if (len) {
                MPC_CUDA_TRY(cudaMemcpyAsync(dsc + lo, scalars + lo * 4, len * sizeof(Fr), cudaMemcpyHostToDevice, lane.s));
                if (di) MPC_CUDA_TRY(cudaMemcpyAsync(di + lo, inf + lo, len, cudaMemcpyHostToDevice, lane.s));
                MPC_CUDA_TRY(cudaMemcpyAsync(db + lo, bases_xy + lo * (sizeof(Affine<F>) / 8), len * sizeof(Affine<F>),
                                             cudaMemcpyHostToDevice, lane.s));
            }
            MPC_CUDA_TRY(cudaEventCreateWithFlags(&lane.ev[k], cudaEventDisableTiming));
            MPC_CUDA_TRY(cudaEventRecord(lane.ev[k], lane.s));
        }
        for (uint32_t k = 0; k < K; k++) {
            const size_t lo = bound[k], len = bound[k + 1] - lo;
            MPC_CUDA_TRY(cudaStreamWaitEvent(s, lane.ev[k], 0));
            if (len) MPC_TRY(job.chunk(db + lo, di ? di + lo : nullptr, dsc + lo, len));
        }
        MPC_TRY(job.finish(res));
    }
    k_emit<F><<<1, 32, 0, s>>>(res, 1, 0, out);
    MPC_KERNEL_CHECK();
    uint32_t host[2 * N + 1];
    MPC_CUDA_TRY(cudaMemcpyAsync(host, out, sizeof(host), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_xy, host, 2 * N * sizeof(uint32_t));
    *out_inf = (uint8_t)host[2 * N];
    return MPC_CUDA_OK;
}

struct EventHolder {
    cudaEvent_t e = nullptr;
    EventHolder() {}
    EventHolder(const EventHolder&) = delete;
    EventHolder& operator=(const EventHolder&) = delete;
    ~EventHolder() { if (e) cudaEventDestroy(e); }
    int32_t create() { MPC_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return MPC_CUDA_OK; }
};

// one part of a sharded MSM, on the calling thread's CURRENT device (the caller holds a DeviceScope):
// scalars [lo, hi) of the part -> Jacobian partial -> peer copy into slot `dst` on the home device
template <class F>
int32_t msm_shard_part(const BaseSnap& part, size_t first, size_t len, const uint64_t* scalars_host, const Fr* scalars_dev,
                       uint32_t* dst, int home_cuda, cudaEvent_t home_ready, cudaEvent_t done, cudaStream_t s) {
    constexpr int N = sizeof(F) / 4;
    Scratch ssc, sjac;
    Fr* dsc = nullptr;
    uint32_t* jac;
    MPC_TRY(sjac.alloc(&jac, 3 * N, s));
    const Fr* sc = scalars_dev;
    if (!sc) {
        MPC_TRY(ssc.alloc(&dsc, len, s));
        MPC_CUDA_TRY(cudaMemcpyAsync(dsc, scalars_host, len * sizeof(Fr), cudaMemcpyHostToDevice, s));
        sc = dsc;
    }
    TableRef tbl = table_of(part, first);
    MPC_TRY(msm_emit<F>((const Affine<F>*)part.bases + first, part.inf ? part.inf + first : nullptr, sc, len, 1, jac, nullptr,
                        nullptr, s, &tbl));
    MPC_CUDA_TRY(cudaStreamWaitEvent(s, home_ready, 0));
    MPC_CUDA_TRY(cudaMemcpyPeerAsync(dst, home_cuda, jac, part.ref->cuda_device, 3 * N * sizeof(uint32_t), s));
    MPC_CUDA_TRY(cudaEventRecord(done, s));
    return MPC_CUDA_OK;
}

// Σ over the parts of a sharded vector: part k runs the whole pipeline on device k and leaves a Jacobian
// partial; the partials travel to the calling thread's device over NVLink (peer copies) where they are added
// and normalised (no collective exists for elliptic-curve addition: gather + add).
// scalars: host pointer (scalars_dev == nullptr), or one device pointer per part holding that part's slice
// of [offset, offset + n).
template <class F>
int32_t msm_sharded(const BaseSnap& v, size_t offset, const uint64_t* scalars_host, const uint64_t* const* scalars_dev,
                    size_t n, uint64_t* out_xy, uint8_t* out_inf) {
    constexpr int N = sizeof(F) / 4;
    cudaStream_t s0;
    MPC_TRY(enter(&s0));
    MPC_TRY(enable_peer_access());
    const BaseVec& parent = *v.ref;
    const size_t P = parent.parts.size();
    const int home_cuda = current_device_info()->cuda_device;
    Scratch s_gather, s_pts, s_out;
    uint32_t *gather, *out;
    XYZZ<F>* pts;
    MPC_TRY(s_gather.alloc(&gather, P * 3 * N, s0));
    MPC_TRY(s_pts.alloc(&pts, P, s0));
    MPC_TRY(s_out.alloc(&out, 3 * N + 4, s0));
    MPC_CUDA_TRY(cudaMemsetAsync(gather, 0, P * 3 * N * sizeof(uint32_t), s0));       // z = 0: infinity
    EventHolder home_ready;
    MPC_TRY(home_ready.create());
    MPC_CUDA_TRY(cudaEventRecord(home_ready.e, s0));
    std::vector<EventHolder> done(P);
    for (size_t k = 0; k < P; k++) {
        size_t plo = parent.lo[k], phi = parent.lo[k + 1];
        size_t lo = std::max(plo, offset), hi = std::min(phi, offset + n);      // this part's share of the range
        if (lo >= hi) continue;
        BaseSnap part = snapshot_of(parent.parts[k]);
        DeviceScope scope(part.ref->dev_index);
        MPC_TRY(scope.rc);
        MPC_TRY(done[k].create());                 // an event records only on streams of the device it was created on
        MPC_TRY(msm_shard_part<F>(part, lo - plo, hi - lo, scalars_host ? scalars_host + (lo - offset) * 4 : nullptr,
                                  scalars_dev ? (const Fr*)scalars_dev[k] : nullptr, gather + k * 3 * N, home_cuda,
                                  home_ready.e, done[k].e, scope.s));
    }
    for (size_t k = 0; k < P; k++)
        if (done[k].e) MPC_CUDA_TRY(cudaStreamWaitEvent(s0, done[k].e, 0));
    k_jac_to_xyzz<F><<<(unsigned)((P + 127) / 128), 128, 0, s0>>>((const Jac<F>*)gather, (uint32_t)P, pts);
    MPC_KERNEL_CHECK();
    k_emit<F><<<1, 32, 0, s0>>>(pts, (uint32_t)P, 0, out);
    MPC_KERNEL_CHECK();
    uint32_t host[2 * N + 1];
    MPC_CUDA_TRY(cudaMemcpyAsync(host, out, sizeof(host), cudaMemcpyDeviceToHost, s0));
    MPC_CUDA_TRY(cudaStreamSynchronize(s0));
    // the parts' scratch was freed stream-ordered on their own streams; nothing else to wait for
    memcpy(out_xy, host, 2 * N * sizeof(uint32_t));
    *out_inf = (uint8_t)host[2 * N];
    return MPC_CUDA_OK;
}

}  // namespace
