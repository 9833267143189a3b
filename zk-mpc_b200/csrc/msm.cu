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

constexpr uint32_t MAX_WINDOW_BITS = 16;       // 2^15 histogram bins x 4 B = 128 KB of shared memory
constexpr int ACC_THREADS = 128;
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
    size_t n;
    uint32_t c, nwin, nb;        // window bits, windows, buckets per window = 2^(c-1)
    uint32_t chunks;             // sort chunks per window
    size_t chunk_len;
    uint32_t task_len;           // max points per accumulate task
    size_t max_tasks;
    uint32_t red_m, red_t;       // bucket-reduce: red_t chunks of red_m buckets per window
};

uint32_t log2_ceil(size_t x) {
    uint32_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

Plan make_plan(size_t n, int sm_count) {
    Plan p;
    p.n = n;
    int64_t c = g_opt_msm_window_bits;
    if (c <= 0) {
        // madds = n * windows, bucket work ~ windows * 2^(c-1): c ~ log2(n) - 4 balances them on this part
        int l = (int)log2_ceil(n);
        c = l - 4;
    }
    if (c < 3) c = 3;
    if (c > MAX_WINDOW_BITS) c = MAX_WINDOW_BITS;
    p.c = (uint32_t)c;
    p.nwin = msm::num_windows(p.c);
    p.nb = 1u << (p.c - 1);
    uint32_t want = (uint32_t)((2 * sm_count + p.nwin - 1) / p.nwin);
    size_t by_len = (n + 4095) / 4096;
    p.chunks = (uint32_t)(by_len < want ? by_len : want);
    if (p.chunks < 1) p.chunks = 1;
    p.chunk_len = (n + p.chunks - 1) / p.chunks;
    int64_t tl = g_opt_msm_task_len;
    if (tl <= 0) {
        // enough tasks to balance ~4 rounds over the resident threads, but not so short that joins dominate
        size_t entries = n * (size_t)p.nwin;
        size_t resident = (size_t)sm_count * 384;
        tl = (int64_t)(entries / (resident * 8));
        if (tl < 32) tl = 32;
        if (tl > 256) tl = 256;
    }
    p.task_len = (uint32_t)tl;
    p.max_tasks = n * (size_t)p.nwin / p.task_len + (size_t)p.nwin * p.nb + 1;
    p.red_m = p.nb < 32 ? p.nb : 32;
    p.red_t = p.nb / p.red_m;
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

__global__ void __launch_bounds__(SORT_THREADS) k_hist(const uint32_t* __restrict__ digits, size_t n, uint32_t nb,
                                                       size_t chunk_len, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t sm[];
    uint32_t w = blockIdx.y, ch = blockIdx.x;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) sm[b] = 0;
    __syncthreads();
    size_t lo = (size_t)ch * chunk_len, hi = lo + chunk_len < n ? lo + chunk_len : n;
    const uint32_t* d = digits + (size_t)w * n;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        uint32_t m = d[i] & ~msm::DIGIT_NEG;
        if (m) atomicAdd(&sm[m - 1], 1u);
    }
    __syncthreads();
    uint32_t* out = hist + ((size_t)w * gridDim.x + ch) * nb;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) out[b] = sm[b];
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

// one block per window: bucket sizes / start offsets, and hist[w][chunk][b] -> first slot of that chunk
__global__ void __launch_bounds__(1024) k_scan_window(uint32_t* __restrict__ hist, uint32_t chunks, uint32_t nb,
                                                      uint32_t* __restrict__ bucket_start,
                                                      uint32_t* __restrict__ bucket_size) {
    __shared__ uint32_t sm[1024];
    uint32_t w = blockIdx.x;
    uint32_t per = (nb + blockDim.x - 1) / blockDim.x;
    uint32_t b0 = threadIdx.x * per, b1 = b0 + per < nb ? b0 + per : nb;
    uint32_t* h = hist + (size_t)w * chunks * nb;
    uint32_t tot = 0;
    for (uint32_t b = b0; b < b1; b++) {
        uint32_t s = 0;
        for (uint32_t ch = 0; ch < chunks; ch++) s += h[(size_t)ch * nb + b];
        bucket_size[(size_t)w * nb + b] = s;
        tot += s;
    }
    uint32_t run = block_exclusive_scan(tot, sm, nullptr);
    for (uint32_t b = b0; b < b1; b++) {
        bucket_start[(size_t)w * nb + b] = run;
        for (uint32_t ch = 0; ch < chunks; ch++) {
            uint32_t t = h[(size_t)ch * nb + b];
            h[(size_t)ch * nb + b] = run;
            run += t;
        }
    }
}

// single block: exclusive scan of ceil(size / task_len) over all buckets; counters[0] = total tasks
__global__ void __launch_bounds__(1024) k_task_scan(const uint32_t* __restrict__ bucket_size, size_t nbuckets,
                                                    uint32_t task_len, uint32_t* __restrict__ task_start,
                                                    uint32_t* __restrict__ counters) {
    __shared__ uint32_t sm[1024];
    size_t per = (nbuckets + blockDim.x - 1) / blockDim.x;
    size_t g0 = threadIdx.x * per, g1 = g0 + per < nbuckets ? g0 + per : nbuckets;
    uint32_t tot = 0;
    for (size_t g = g0; g < g1; g++) tot += (bucket_size[g] + task_len - 1) / task_len;
    uint32_t total;
    uint32_t run = block_exclusive_scan(tot, sm, &total);
    for (size_t g = g0; g < g1; g++) {
        task_start[g] = run;
        run += (bucket_size[g] + task_len - 1) / task_len;
    }
    if (threadIdx.x == 0) counters[0] = total;
}

// task = (first slot in the window's sorted list, length, bucket, bucket-has-a-single-task)
__global__ void __launch_bounds__(256) k_build_tasks(const uint32_t* __restrict__ bucket_start,
                                                     const uint32_t* __restrict__ bucket_size,
                                                     const uint32_t* __restrict__ task_start, size_t nbuckets,
                                                     uint32_t task_len, uint4* __restrict__ tasks,
                                                     uint32_t* __restrict__ small_list, uint32_t* __restrict__ big_list,
                                                     uint32_t* __restrict__ counters) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nbuckets) return;
    uint32_t size = bucket_size[g];
    if (size == 0) return;
    uint32_t nt = (size + task_len - 1) / task_len, ts = task_start[g], pos = bucket_start[g];
    for (uint32_t k = 0; k < nt; k++) {
        uint32_t len = size - k * task_len < task_len ? size - k * task_len : task_len;
        tasks[ts + k] = make_uint4(pos + k * task_len, len, (uint32_t)g, nt == 1 ? 1u : 0u);
    }
    if (nt > 1) {
        if (nt <= SMALL_MULTI_MAX) small_list[atomicAdd(&counters[2], 1u)] = (uint32_t)g;
        else big_list[atomicAdd(&counters[3], 1u)] = (uint32_t)g;
    }
}

__global__ void __launch_bounds__(SORT_THREADS) k_scatter(const uint32_t* __restrict__ digits, size_t n, uint32_t nb,
                                                          size_t chunk_len, const uint32_t* __restrict__ cursors,
                                                          uint32_t* __restrict__ sorted) {
    extern __shared__ uint32_t sm[];
    uint32_t w = blockIdx.y, ch = blockIdx.x;
    const uint32_t* cur = cursors + ((size_t)w * gridDim.x + ch) * nb;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) sm[b] = cur[b];
    __syncthreads();
    size_t lo = (size_t)ch * chunk_len, hi = lo + chunk_len < n ? lo + chunk_len : n;
    const uint32_t* d = digits + (size_t)w * n;
    uint32_t* out = sorted + (size_t)w * n;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        uint32_t e = d[i], m = e & ~msm::DIGIT_NEG;
        if (m) {
            uint32_t pos = atomicAdd(&sm[m - 1], 1u);
            out[pos] = (uint32_t)i | (e & msm::DIGIT_NEG);
        }
    }
}

// ---- the hot kernel -----------------------------------------------------------------------------------
// Every thread repeatedly claims a task (a run of <= task_len sorted entries of one bucket) and adds the
// referenced affine bases into an XYZZ accumulator.  The claim is folded into the point loop, so the
// lanes of a warp stay converged on the mixed addition whatever the task lengths are.
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_accumulate(const Affine<F>* __restrict__ bases,
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

// one block per window: sum of its T chunk results
template <class F>
__global__ void __launch_bounds__(ACC_THREADS) k_window_sum(const XYZZ<F>* __restrict__ chunk_res, uint32_t T,
                                                            XYZZ<F>* __restrict__ window_sum) {
    extern __shared__ uint4 sm_raw[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(sm_raw);
    uint32_t w = blockIdx.x;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t k = threadIdx.x; k < T; k += blockDim.x) {
        XYZZ<F> q = load_pod(chunk_res + (size_t)w * T + k);
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
        store_pod(window_sum + w, a);
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

// result (one XYZZ on the device) = Σ scalars[i] * bases[i]
template <class F>
int32_t msm_run(const Affine<F>* bases, const uint8_t* inf, const Fr* scalars, size_t n, XYZZ<F>* result,
                cudaStream_t s) {
    if (n == 0) {
        MPC_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(XYZZ<F>), s));
        return MPC_CUDA_OK;
    }
    MPC_ARG_CHECK(n < ((size_t)1 << 31));
    const DeviceInfo* dev = current_device_info();
    Plan p = make_plan(n, dev->sm_count);
    size_t nbuckets = (size_t)p.nwin * p.nb;
    MPC_ARG_CHECK(nbuckets < ((size_t)1 << 31) && (size_t)p.nwin * n / p.task_len + nbuckets < ((size_t)1 << 32));

    Scratch s_digits, s_sorted, s_hist, s_bstart, s_bsize, s_tstart, s_tasks, s_small, s_big, s_cnt, s_buckets,
        s_partials, s_chunk, s_wsum;
    uint32_t *digits, *sorted, *hist, *bstart, *bsize, *tstart, *small_list, *big_list, *counters;
    uint4* tasks;
    XYZZ<F>*buckets, *partials, *chunk_res, *wsum;
    MPC_TRY(s_digits.alloc(&digits, (size_t)p.nwin * n, s));
    MPC_TRY(s_sorted.alloc(&sorted, (size_t)p.nwin * n, s));
    MPC_TRY(s_hist.alloc(&hist, (size_t)p.nwin * p.chunks * p.nb, s));
    MPC_TRY(s_bstart.alloc(&bstart, nbuckets, s));
    MPC_TRY(s_bsize.alloc(&bsize, nbuckets, s));
    MPC_TRY(s_tstart.alloc(&tstart, nbuckets, s));
    MPC_TRY(s_tasks.alloc(&tasks, p.max_tasks, s));
    MPC_TRY(s_small.alloc(&small_list, nbuckets, s));
    MPC_TRY(s_big.alloc(&big_list, nbuckets, s));
    MPC_TRY(s_cnt.alloc(&counters, 8, s));
    MPC_TRY(s_buckets.alloc(&buckets, nbuckets, s));
    MPC_TRY(s_partials.alloc(&partials, p.max_tasks, s));
    MPC_TRY(s_chunk.alloc(&chunk_res, (size_t)p.nwin * p.red_t, s));
    MPC_TRY(s_wsum.alloc(&wsum, p.nwin, s));

    ProfileScope prof_total("msm_total", s);
    profile_begin("msm_sort", s);
    MPC_CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(uint32_t), s));
    MPC_CUDA_TRY(cudaMemsetAsync(buckets, 0, nbuckets * sizeof(XYZZ<F>), s));      // all-zero = infinity

    k_digits<<<grid_for(n, 256, 8), 256, 0, s>>>(scalars, inf, n, p.c, p.nwin, digits);
    MPC_KERNEL_CHECK();

    size_t sort_smem = (size_t)p.nb * sizeof(uint32_t);
    MPC_TRY(allow_smem(k_hist, sort_smem));
    MPC_TRY(allow_smem(k_scatter, sort_smem));
    dim3 sort_grid(p.chunks, p.nwin);
    k_hist<<<sort_grid, SORT_THREADS, sort_smem, s>>>(digits, n, p.nb, p.chunk_len, hist);
    MPC_KERNEL_CHECK();
    k_scan_window<<<p.nwin, 1024, 0, s>>>(hist, p.chunks, p.nb, bstart, bsize);
    MPC_KERNEL_CHECK();
    k_task_scan<<<1, 1024, 0, s>>>(bsize, nbuckets, p.task_len, tstart, counters);
    MPC_KERNEL_CHECK();
    k_build_tasks<<<(unsigned)((nbuckets + 255) / 256), 256, 0, s>>>(bstart, bsize, tstart, nbuckets, p.task_len, tasks,
                                                                    small_list, big_list, counters);
    MPC_KERNEL_CHECK();
    k_scatter<<<sort_grid, SORT_THREADS, sort_smem, s>>>(digits, n, p.nb, p.chunk_len, hist, sorted);
    MPC_KERNEL_CHECK();

    profile_end("msm_sort", s);

    int acc_blocks = 0;
    MPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_blocks, k_accumulate<F>, ACC_THREADS, 0));
    if (acc_blocks < 1) acc_blocks = 1;
    profile_begin("msm_accumulate", s);
    k_accumulate<F><<<dev->sm_count * acc_blocks, ACC_THREADS, 0, s>>>(bases, sorted, n, p.nb, tasks, counters, buckets,
                                                                      partials);
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

    uint32_t red_threads = p.nwin * p.red_t;
    k_bucket_reduce<F><<<(red_threads + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, s>>>(buckets, p.nwin, p.nb,
                                                                                            p.red_m, p.red_t, chunk_res);
    MPC_KERNEL_CHECK();
    k_window_sum<F><<<p.nwin, ACC_THREADS, tree_smem, s>>>(chunk_res, p.red_t, wsum);
    MPC_KERNEL_CHECK();
    k_horner<F><<<1, 32, 0, s>>>(wsum, p.nwin, p.c, result);
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

// run + emit; scalars already on the device.  host_out: affine limbs + flag copied back and synchronised.
template <class F>
int32_t msm_emit(const Affine<F>* bases, const uint8_t* inf, const Fr* scalars_dev, size_t n, uint32_t mode,
                 uint32_t* out_dev, uint64_t* host_xy, uint8_t* host_inf, cudaStream_t s) {
    constexpr int N = sizeof(F) / 4;
    Scratch s_res, s_out;
    XYZZ<F>* res;
    MPC_TRY(s_res.alloc(&res, 1, s));
    MPC_TRY(msm_run<F>(bases, inf, scalars_dev, n, res, s));
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
    return msm_emit<F>((const Affine<F>*)v.bases + offset, v.inf ? v.inf + offset : nullptr, dsc, n, 0, nullptr, out_xy,
                       out_inf, s);
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
    if (v.owned) {
        int cur = 0;
        MPC_CUDA_TRY(cudaGetDevice(&cur));
        MPC_CUDA_TRY(cudaSetDevice(v.cuda_device));
        cudaFree(v.bases);
        if (v.inf) cudaFree(v.inf);
        MPC_CUDA_TRY(cudaSetDevice(cur));
    }
    return MPC_CUDA_OK;
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
    return msm_emit<Fq>((const Affine<Fq>*)v.bases + offset, v.inf ? v.inf + offset : nullptr,
                        (const Fr*)scalars_mont_dev, n, 1, (uint32_t*)out_jac_dev, nullptr, nullptr,
                        pick_stream(stream, s));
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
