// Share NTT over Fr: in-order radix-2 transforms of Radix2EvaluationDomain, computed as multi-stage
// shared-memory passes (decimation in frequency) with the bit reversal fused into the last pass.
//
// Replaces  fft_in_place / ifft_in_place / coset_ifft_in_place   arkworks/algebra/poly/src/domain/radix2/mod.rs:99-114
//           coset_fft_in_place                                    arkworks/algebra/poly/src/domain/mod.rs:138-141
//           io_helper / oi_helper / derange / roots_of_unity      arkworks/algebra/poly/src/domain/radix2/fft.rs:76-78,185-307
//           distribute_powers_and_mul_by_const                    arkworks/algebra/poly/src/domain/mod.rs:98-105
//           divide_by_vanishing_poly_on_coset_in_place            arkworks/algebra/poly/src/domain/mod.rs:183-190
// The outputs are exact field elements of the uniquely defined transforms
//   fft: out[i] = Σ a[j] ω^(ij);  ifft: out[i] = n⁻¹ Σ a[j] ω^(-ij);  coset variants pre/post-scale by 22^(±i),
// so any correct evaluation order is bit-identical to the reference's io_helper/oi_helper loops.
//
// Layout of one pass over stages [s0, s0+deg):  element index i = hi·(n>>s0) + j·T + lo with
// T = n >> (s0+deg); a CTA owns a tile of 2^deg rows (j) x C adjacent columns (lo), i.e. 128-byte
// contiguous runs in HBM, keeps it in shared memory as 8 limb planes (bank-conflict-free for
// consecutive elements), and runs `deg` butterfly stages on it.  The last pass (T = 1) gathers C
// sub-transforms whose bit-reversed destinations are adjacent, so its stores are 128-byte runs too,
// and fuses the n⁻¹ / coset scaling.  The index algebra is modelled and tested in tests/ntt_model.py.
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace mpc;

namespace mpc {
std::atomic<int64_t> g_opt_ntt_occupancy{0};    // 1: the pass kernels compiled for one more CTA per SM
std::atomic<int64_t> g_opt_ntt_generic{0};
std::atomic<int64_t> g_opt_ntt_graph{0};        // sharded NTT replay through CUDA graphs: 1 = on (0 / 2 = off)      // 1: always run the generic pass kernel (A/B measurements, tests)
}

namespace {

constexpr int NTT_MAX_DEG = 8;
constexpr int NTT_LOG_C = 2;               // 4 adjacent 32-byte elements = one 128-byte line
constexpr int COSET_LO_BITS = 12;
constexpr int MAX_LOG_N = 30;

// ---- small kernels that build the per-device constant tables (all arithmetic stays on the GPU) -----
// consts[l] = ω_l (primitive 2^l-th root: TWO_ADIC_ROOT squared 47-l times, ff/src/fields/mod.rs:363-378)
// consts[48+l] = (2^l)⁻¹ = size_inv,  consts[96+l] = (22^(2^l) - 1)⁻¹ (vanishing polynomial on the coset)
__global__ void k_domain_consts(Fr root47, Fr two_inv, Fr gen, Fr* __restrict__ consts) {
    uint32_t l = threadIdx.x;
    if (l > 47) return;
    Fr w = root47;
    for (uint32_t i = l; i < 47; i++) w = sqr(w);
    consts[l] = w;
    Fr ni = Fr::one();
    for (uint32_t i = 0; i < l; i++) ni = mul(ni, two_inv);
    consts[48 + l] = ni;
    Fr z = gen;
    for (uint32_t i = 0; i < l; i++) z = sqr(z);
    consts[96 + l] = inv(sub(z, Fr::one()));
}

// out[i] = (base^(2^pre))^i
__global__ void __launch_bounds__(256) k_powers(const Fr* __restrict__ base_ptr, uint32_t pre, size_t count,
                                                Fr* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr b = load_fe_ro(base_ptr);
    for (uint32_t k = 0; k < pre; k++) b = sqr(b);
    Fr acc = Fr::one();
    bool started = false;
    for (int bit = 63 - __clzll((unsigned long long)(i | 1)); bit >= 0; bit--) {
        if (started) acc = sqr(acc);
        if ((i >> bit) & 1) {
            acc = started ? mul(acc, b) : b;
            started = true;
        }
    }
    store_fe(out + i, acc);
}

__global__ void __launch_bounds__(256) k_mul_const_dev(Fr* __restrict__ data, const Fr* __restrict__ c, size_t n) {
    Fr k = load_fe_ro(c);
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        store_fe(data + i, mul(load_fe(data + i), k));
}

// ---- the pass kernel ----------------------------------------------------------------------------------
struct PassArgs {
    const Fr* src;
    Fr* dst;
    const Fr* tw[NTT_MAX_DEG];     // tw[r] = forward table of domain l = log_n - s0 - r  (ω_l^i, i < 2^(l-1))
    const Fr* pre_lo;              // coset_fft (first pass): x *= pre_lo[i & 4095] * pre_hi[i >> 12]
    const Fr* pre_hi;
    const Fr* post_lo;             // coset_ifft (last pass): x *= scale * post_lo[d & 4095] * post_hi[d >> 12]
    const Fr* post_hi;
    const Fr* scale;               // ifft / coset_ifft (last pass): n⁻¹
    uint32_t log_n, s0, deg, lc;   // lc = log2(columns per tile)
    uint32_t last, inverse;
};

// Tile element e lives at word swz(e) of each of the 8 limb planes.  The XOR swizzle makes every access pattern
// of the kernel hit 32 distinct banks per warp: a warp touches elements whose index varies in 5 bit positions
// {0,1} + three of {2..9} (contiguous runs of the load/store phases, the strided item sets of the radix-4 rounds
// with half = 4, 2, 1, the radix-2 tail, the bit-reversed rows of the last store); bank bits 2..4 are b2^b5^b7,
// b3^b5^b6^b8 and b4^b6^b9, whose restriction to each of those triples is invertible over GF(2).  swz is linear
// over GF(2), so for an item whose four elements differ only in bits that are zero in e0,
// swz(e0 + k*es) = swz(e0) ^ swz(k*es): one XOR per element.
DEV uint32_t swz(uint32_t e) {
    return e ^ (((e >> 5) & 1u) * 0xCu) ^ (((e >> 6) & 1u) * 0x18u) ^ (((e >> 7) & 1u) << 2) ^ (((e >> 8) & 1u) << 3) ^
           (((e >> 9) & 1u) << 4);
}
// p = physical (swizzled) word index
DEV Fr sm_load(const uint32_t* sm, uint32_t E, uint32_t p) {
    Fr r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = sm[l * E + p];
    return r;
}
DEV void sm_store(uint32_t* sm, uint32_t E, uint32_t p, const Fr& a) {
#pragma unroll
    for (int l = 0; l < 8; l++) sm[l * E + p] = a.v[l];
}
DEV uint32_t bitrev(uint32_t x, uint32_t bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

// one DIF butterfly: u' = u + v, v' = (u - v) * w^k with w = ω_l, read from the forward table of domain l;
// inverse transforms use ω^-k = -ω^(2^(l-1) - k) (the sign is folded into the subtraction order)
DEV void bfly(Fr& u, Fr& v, const Fr* __restrict__ tw, uint32_t k, uint32_t l, bool inverse) {
    bool swap = inverse && k != 0;
    if (swap) k = (1u << (l - 1)) - k;
    Fr d = swap ? sub(v, u) : sub(u, v);
    u = add(u, v);
    v = l > 1 ? mul(d, load_fe_ro(tw + k)) : d;       // l == 1: the only twiddle is 1
}

// The same butterfly on lazily reduced values (fp.cuh: everything in the tile stays in [0, 2p); one conditional
// subtraction per addition, none per product): the pass kernel's version.  INV is a compile-time flag so the
// forward kernels carry no operand selects.
// ONE: the caller knows k == 0 for every lane (twiddle 1): no product, no table read.
template <bool INV, bool ONE = false>
DEV void bfly_lazy(Fr& u, Fr& v, const Fr* __restrict__ tw, uint32_t k, uint32_t l) {
    Fr d;
    if (ONE) {
        d = sub_lazy(u, v);
        u = add_lazy(u, v);
        fp_detail::cond_sub_2p<consts::FrParams>(d.v);
        v = d;
        return;
    }
    if (INV) {
        const bool swap = k != 0;
        if (swap) k = (1u << (l - 1)) - k;
        d = swap ? sub_lazy(v, u) : sub_lazy(u, v);
    } else {
        d = sub_lazy(u, v);
    }
    u = add_lazy(u, v);
    if (l > 1) {
        v = mul_lazy(d, load_fe_ro(tw + k));
    } else {                                          // l == 1: the only twiddle is 1
        fp_detail::cond_sub_2p<consts::FrParams>(d.v);
        v = d;
    }
}

// Every thread owns four rows of the tile and runs TWO stages on them in registers between
// shared-memory exchanges (radix-4 step = 4 products, half the shared traffic and barriers of radix-2);
// an odd stage count ends with one radix-2 stage.
// DEG > 0: the tile shape (2^DEG rows x 4 columns, the shape of every pass of a transform of 2^12 or more
// elements) is a compile-time constant, so the rounds unroll and the index algebra folds into immediates;
// DEG = 0 is the generic kernel for the small shapes.
constexpr int NTT_THREADS = 256;
// OCC = 1: compiled for one more resident CTA per SM (64 registers instead of 80)
template <int DEG, int LAST, bool INV, int OCC>
__global__ void __launch_bounds__(DEG ? (1 << DEG) : NTT_THREADS,
                                  (DEG == 8 ? 3 : DEG == 7 ? 6 : DEG == 6 ? 12 : 3) + (OCC ? (DEG == 8 ? 1 : DEG == 7 ? 2 : DEG == 6 ? 4 : 1) : 0))
    k_ntt_pass(PassArgs a) {
    extern __shared__ uint32_t sm[];
    const uint32_t t = threadIdx.x;
    const uint32_t deg = DEG ? DEG : a.deg, lc = DEG ? NTT_LOG_C : a.lc;
    const bool last = DEG ? LAST != 0 : a.last != 0;
    const uint32_t rows = 1u << deg, C = 1u << lc, E = rows << lc;
    const uint32_t log_t = a.log_n - a.s0 - deg;            // T = 2^log_t: distance of the tile's rows
    const uint32_t T = 1u << log_t;
    const size_t n = (size_t)1 << a.log_n;
    const uint32_t tile = blockIdx.x;
    const Fr* src = a.src + (size_t)blockIdx.y * n;
    Fr* dst = a.dst + (size_t)blockIdx.y * n;

    // tile origin: non-last: i = origin + j*T + c;  last: i = (hb + (bitrev(c) << (s0-lc))) * rows + j
    uint32_t origin = 0, lo0 = 0, hb = 0;
    if (!last) {
        uint32_t log_tiles_per_hi = log_t - lc;
        uint32_t hi = tile >> log_tiles_per_hi;
        lo0 = (tile & ((1u << log_tiles_per_hi) - 1)) << lc;
        origin = (hi << (a.log_n - a.s0)) + lo0;
    } else {
        hb = bitrev(tile << lc, a.s0);
    }

    // ---- load: tile slot e = (row j, column c); consecutive threads take consecutive slots
    for (uint32_t e = t; e < E; e += blockDim.x) {
        const uint32_t j = e >> lc, c = e & (C - 1);
        uint32_t i;
        if (!last) i = origin + (j << log_t) + c;
        else i = ((hb + (bitrev(c, lc) << (a.s0 - lc))) << deg) + j;     // 4 runs of 8 x 32 B per warp
        Fr x = load_fe(src + i);
        if (a.pre_lo) {
            x = mul(x, load_fe_ro(a.pre_lo + (i & ((1u << COSET_LO_BITS) - 1))));
            if (a.log_n > COSET_LO_BITS) x = mul(x, load_fe_ro(a.pre_hi + (i >> COSET_LO_BITS)));
        }
        sm_store(sm, E, swz(e), x);
    }
    __syncthreads();

    // ---- radix-4 rounds: stages r (half = 2h) and r+1 (half = h) on rows base + {0, h, 2h, 3h}
    uint32_t r = 0;
#pragma unroll
    for (; r + 2 <= deg; r += 2) {
        const uint32_t hbits = deg - 2 - r, h = 1u << hbits;
        const uint32_t l0 = a.log_n - a.s0 - r;
        const uint32_t es = h << lc, s1 = swz(es), s2 = swz(2 * es), s3 = s1 ^ s2;
        for (uint32_t item = t; item < E / 4; item += blockDim.x) {
            const uint32_t c = item & (C - 1), q = item >> lc;
            const uint32_t lo = last ? 0 : lo0 + c;
            const uint32_t qq = q & (h - 1);
            const uint32_t p0 = swz(((((q >> hbits) << (hbits + 2)) | qq) << lc) + c);
            Fr x0 = sm_load(sm, E, p0), x1 = sm_load(sm, E, p0 ^ s1), x2 = sm_load(sm, E, p0 ^ s2),
               x3 = sm_load(sm, E, p0 ^ s3);
            const uint32_t k = (qq << log_t) + lo;
            // last round of the last pass: qq = lo = 0 in every lane, so the first butterfly's twiddle is 1
            if (last && hbits == 0) bfly_lazy<INV, true>(x0, x2, a.tw[r], 0, l0);
            else bfly_lazy<INV>(x0, x2, a.tw[r], k, l0);
            bfly_lazy<INV>(x1, x3, a.tw[r], k + (h << log_t), l0);
            bfly_lazy<INV>(x0, x1, a.tw[r + 1], k, l0 - 1);
            bfly_lazy<INV>(x2, x3, a.tw[r + 1], k, l0 - 1);
            sm_store(sm, E, p0, x0);
            sm_store(sm, E, p0 ^ s1, x1);
            sm_store(sm, E, p0 ^ s2, x2);
            sm_store(sm, E, p0 ^ s3, x3);
        }
        __syncthreads();
    }
    // ---- odd stage count: last stage of the pass (half = 1) pairs rows (2b, 2b+1)
    if (r < deg) {
        const uint32_t l0 = a.log_n - a.s0 - r;
        for (uint32_t item = t; item < E / 2; item += blockDim.x) {
            const uint32_t c = item & (C - 1), bq = item >> lc;
            const uint32_t lo = last ? 0 : lo0 + c;
            const uint32_t p0 = swz(((2 * bq) << lc) + c), p1 = p0 ^ swz(C);
            Fr x0 = sm_load(sm, E, p0), x1 = sm_load(sm, E, p1);
            bfly_lazy<INV>(x0, x1, a.tw[r], lo, l0);
            sm_store(sm, E, p0, x0);
            sm_store(sm, E, p1, x1);
        }
        __syncthreads();
    }

    // ---- store
    for (uint32_t e = t; e < E; e += blockDim.x) {
        uint32_t j = e >> lc, cc = e & (C - 1);
        if (!last) {
            store_fe(dst + origin + (j << log_t) + cc, sm_load(sm, E, swz(e)));
        } else {
            // row j of the store is destination block j: source row bitrev(j)
            Fr x = sm_load(sm, E, swz((bitrev(j, deg) << lc) + cc));
            uint32_t d = (j << a.s0) + (tile << lc) + cc;
            // back to the canonical range: a strict product does it (input < 2p), else two conditional subtractions
            if (a.scale) x = mul(x, load_fe_ro(a.scale));
            else if (!a.post_lo) x = reduce_full(x);
            if (a.post_lo) {
                x = mul(x, load_fe_ro(a.post_lo + (d & ((1u << COSET_LO_BITS) - 1))));
                if (a.log_n > COSET_LO_BITS) x = mul(x, load_fe_ro(a.post_hi + (d >> COSET_LO_BITS)));
            }
            store_fe(dst + d, x);
        }
    }
}

template <int DEG, int LAST, bool INV, int OCC>
int32_t launch_pass_occ(const PassArgs& a, dim3 grid, uint32_t threads, size_t smem, cudaStream_t s) {
    static bool configured[64] = {};                 // per device of the init list; benign race (idempotent call)
    int dev = current_device_index();
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        // 16-32 KB tiles: let several CTAs share an SM (the default carveout fits one)
        MPC_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass<DEG, LAST, INV, OCC>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        configured[dev] = true;
    }
    k_ntt_pass<DEG, LAST, INV, OCC><<<grid, threads, smem, s>>>(a);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}
template <int DEG, int LAST>
int32_t launch_pass(const PassArgs& a, dim3 grid, uint32_t threads, size_t smem, cudaStream_t s) {
    // measured on B200 (tools/time_ntt.py): the 64-register build wins for the 2^8-row tiles (4 CTAs of 256 threads
    // per SM instead of 3: 3.52 -> 3.34 ms at 2^24) and loses for the smaller ones; option ntt_occupancy: 1 = all
    // shapes, 2 = none
    const int64_t occ = g_opt_ntt_occupancy.load(std::memory_order_relaxed);
    if (DEG && (occ == 1 || (occ == 0 && DEG == 8)))
        return a.inverse ? launch_pass_occ<DEG, LAST, true, DEG ? 1 : 0>(a, grid, threads, smem, s)
                         : launch_pass_occ<DEG, LAST, false, DEG ? 1 : 0>(a, grid, threads, smem, s);
    return a.inverse ? launch_pass_occ<DEG, LAST, true, 0>(a, grid, threads, smem, s)
                     : launch_pass_occ<DEG, LAST, false, 0>(a, grid, threads, smem, s);
}

// n == 1: ifft / coset_ifft multiply by 1⁻¹ = 1, everything is the identity
// ---- per-device domain cache ---------------------------------------------------------------------------
struct DomainCache {
    bool ready = false;
    Fr* consts = nullptr;                  // 144 entries, see k_domain_consts
    Fr* tw[MAX_LOG_N + 1] = {};            // forward twiddle tables by domain log-size
    Fr *g_lo = nullptr, *gi_lo = nullptr;  // 22^i, 22^-i, i < 4096
    Fr *g_hi = nullptr, *gi_hi = nullptr;  // 22^(±4096 h), h < 2^hi_bits
    Fr* gi_lo_scaled[MAX_LOG_N + 1] = {};  // n^-1 * 22^-i, i < 4096, per domain size: coset_ifft folds its 1/n here
    Fr* gen_pair = nullptr;                // [22, 22⁻¹]
    uint32_t hi_bits = 0;
};
std::mutex g_ntt_mu;
DomainCache g_cache[64];

Fr fr_from_limbs(const uint32_t* l) {
    Fr r;
    memcpy(r.v, l, sizeof(r.v));
    return r;
}

// Make sure the tables a transform of size 2^log_n needs exist on the current device and return a COPY of the
// cache entry taken under the lock.  Published tables are never freed: when a larger coset table replaces a
// smaller one, the old arrays stay allocated (retired) because another party thread may already hold their
// pointers for a launch it has not issued yet (LocalTestNet runs all parties on one device).
int32_t ensure_tables(int dev_index, uint32_t log_n, bool coset, cudaStream_t s, DomainCache* out) {
    std::lock_guard<std::mutex> lk(g_ntt_mu);
    DomainCache& d = g_cache[dev_index];
    bool dirty = false;
    if (!d.ready) {
        MPC_CUDA_TRY(cudaMalloc((void**)&d.consts, 144 * sizeof(Fr)));
        MPC_CUDA_TRY(cudaMalloc((void**)&d.gen_pair, 2 * sizeof(Fr)));
        k_domain_consts<<<1, 64, 0, s>>>(fr_from_limbs(consts::FR_TWO_ADIC_ROOT), fr_from_limbs(consts::FR_TWO_INV),
                                        fr_from_limbs(consts::FR_GENERATOR), d.consts);
        MPC_KERNEL_CHECK();
        Fr pair[2] = {fr_from_limbs(consts::FR_GENERATOR), fr_from_limbs(consts::FR_GENERATOR_INV)};
        MPC_CUDA_TRY(cudaMemcpyAsync(d.gen_pair, pair, sizeof(pair), cudaMemcpyHostToDevice, s));
        MPC_CUDA_TRY(cudaStreamSynchronize(s));       // `pair` is a stack buffer
        d.ready = true;
    }
    for (uint32_t l = 2; l <= log_n; l++) {
        if (d.tw[l]) continue;
        size_t count = (size_t)1 << (l - 1);
        Fr* t = nullptr;
        MPC_CUDA_TRY(cudaMalloc((void**)&t, count * sizeof(Fr)));
        k_powers<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d.consts + l, 0, count, t);
        MPC_KERNEL_CHECK();
        MPC_CUDA_TRY(cudaStreamSynchronize(s));       // publish only complete tables
        d.tw[l] = t;
    }
    if (coset) {
        const size_t lo_n = (size_t)1 << COSET_LO_BITS;
        if (!d.g_lo) {
            Fr *a = nullptr, *b = nullptr;
            MPC_CUDA_TRY(cudaMalloc((void**)&a, lo_n * sizeof(Fr)));
            MPC_CUDA_TRY(cudaMalloc((void**)&b, lo_n * sizeof(Fr)));
            k_powers<<<(unsigned)(lo_n / 256), 256, 0, s>>>(d.gen_pair, 0, lo_n, a);
            MPC_KERNEL_CHECK();
            k_powers<<<(unsigned)(lo_n / 256), 256, 0, s>>>(d.gen_pair + 1, 0, lo_n, b);
            MPC_KERNEL_CHECK();
            MPC_CUDA_TRY(cudaStreamSynchronize(s));
            d.g_lo = a;
            d.gi_lo = b;
        }
        if (log_n >= 1 && !d.gi_lo_scaled[log_n]) {
            Fr* a = nullptr;
            MPC_CUDA_TRY(cudaMalloc((void**)&a, lo_n * sizeof(Fr)));
            MPC_CUDA_TRY(cudaMemcpyAsync(a, d.gi_lo, lo_n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
            k_mul_const_dev<<<(unsigned)(lo_n / 256), 256, 0, s>>>(a, d.consts + 48 + log_n, lo_n);
            MPC_KERNEL_CHECK();
            MPC_CUDA_TRY(cudaStreamSynchronize(s));
            d.gi_lo_scaled[log_n] = a;
        }
        uint32_t need = log_n > COSET_LO_BITS ? log_n - COSET_LO_BITS : 0;
        if (need && (!d.g_hi || need > d.hi_bits)) {
            size_t hi_n = (size_t)1 << need;
            Fr *a = nullptr, *b = nullptr;
            MPC_CUDA_TRY(cudaMalloc((void**)&a, hi_n * sizeof(Fr)));
            MPC_CUDA_TRY(cudaMalloc((void**)&b, hi_n * sizeof(Fr)));
            k_powers<<<(unsigned)((hi_n + 255) / 256), 256, 0, s>>>(d.gen_pair, COSET_LO_BITS, hi_n, a);
            MPC_KERNEL_CHECK();
            k_powers<<<(unsigned)((hi_n + 255) / 256), 256, 0, s>>>(d.gen_pair + 1, COSET_LO_BITS, hi_n, b);
            MPC_KERNEL_CHECK();
            MPC_CUDA_TRY(cudaStreamSynchronize(s));
            // the superseded (smaller) tables stay allocated: total retired size < the live tables
            d.g_hi = a;
            d.gi_hi = b;
            d.hi_bits = need;
        }
    }
    (void)dirty;
    *out = d;
    return MPC_CUDA_OK;
}

// tmp_override: caller-owned ping-pong buffer of n * batch elements (the captured multi-GPU path must not allocate)
int32_t ntt_dev(Fr* data, uint32_t log_n, uint32_t kind, uint32_t batch, cudaStream_t s, Fr* tmp_override = nullptr) {
    MPC_ARG_CHECK(kind <= MPC_CUDA_NTT_COSET_IFFT);
    MPC_ARG_CHECK(log_n <= MAX_LOG_N && log_n <= (uint32_t)consts::FR_TWO_ADICITY);
    if (batch == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(data != nullptr);
    if (log_n == 0) return MPC_CUDA_OK;      // size-1 domain: every kind is the identity (size_inv = 1, g^0 = 1)
    const bool inverse = kind == MPC_CUDA_NTT_IFFT || kind == MPC_CUDA_NTT_COSET_IFFT;
    const bool coset = kind >= MPC_CUDA_NTT_COSET_FFT;
    DomainCache dc;
    MPC_TRY(ensure_tables(current_device_index(), log_n, coset, s, &dc));
    const DomainCache* d = &dc;

    ProfileScope prof("ntt", s);
    const size_t n = (size_t)1 << log_n;
    uint32_t npass = (log_n + NTT_MAX_DEG - 1) / NTT_MAX_DEG;
    uint32_t base = log_n / npass, extra = log_n % npass;
    Scratch stmp;
    Fr* tmp = tmp_override;
    if (npass > 1 && !tmp) MPC_TRY(stmp.alloc(&tmp, n * batch, s));

    uint32_t s0 = 0;
    for (uint32_t p = 0; p < npass; p++) {
        uint32_t deg = base + (p < extra ? 1 : 0);
        bool first = p == 0, last = p + 1 == npass;
        PassArgs a;
        memset(&a, 0, sizeof(a));
        a.src = first ? data : tmp;
        a.dst = last ? data : tmp;
        a.log_n = log_n; a.s0 = s0; a.deg = deg;
        a.last = last; a.inverse = inverse;
        for (uint32_t r = 0; r < deg; r++) a.tw[r] = d->tw[log_n - s0 - r];      // l = 1 entry is nullptr (unused)
        size_t T = n >> (s0 + deg);
        size_t tiles;
        if (!last) {
            uint32_t lc = NTT_LOG_C;
            while (((size_t)1 << lc) > T) lc--;
            a.lc = lc;
            tiles = ((size_t)1 << s0) * (T >> lc);
        } else {
            a.lc = s0 < (uint32_t)NTT_LOG_C ? s0 : NTT_LOG_C;
            tiles = ((size_t)1 << s0) >> a.lc;
        }
        if (first && kind == MPC_CUDA_NTT_COSET_FFT) { a.pre_lo = d->g_lo; a.pre_hi = d->g_hi; }
        if (last && kind == MPC_CUDA_NTT_IFFT) a.scale = d->consts + 48 + log_n;
        if (last && kind == MPC_CUDA_NTT_COSET_IFFT) { a.post_lo = d->gi_lo_scaled[log_n]; a.post_hi = d->gi_hi; }   // 1/n folded in
        uint32_t E = (1u << deg) << a.lc;
        MPC_ARG_CHECK(tiles < ((size_t)1 << 31) && batch < 65536);
        dim3 grid((unsigned)tiles, batch);
        uint32_t threads = E / 4 < 32 ? 32 : (E / 4 > NTT_THREADS ? NTT_THREADS : E / 4);
        const size_t smem = (size_t)E * sizeof(Fr);
        const bool shaped = a.lc == (uint32_t)NTT_LOG_C && g_opt_ntt_generic.load(std::memory_order_relaxed) == 0;
        int32_t rc;
        if (shaped && deg == 8) rc = last ? launch_pass<8, 1>(a, grid, 256, smem, s) : launch_pass<8, 0>(a, grid, 256, smem, s);
        else if (shaped && deg == 7) rc = last ? launch_pass<7, 1>(a, grid, 128, smem, s) : launch_pass<7, 0>(a, grid, 128, smem, s);
        else if (shaped && deg == 6) rc = last ? launch_pass<6, 1>(a, grid, 64, smem, s) : launch_pass<6, 0>(a, grid, 64, smem, s);
        else rc = launch_pass<0, 0>(a, grid, threads, smem, s);
        MPC_TRY(rc);
        s0 += deg;
    }
    return MPC_CUDA_OK;
}


// ---- multi-GPU: the stages that cross device boundaries -------------------------------------------------
// With the vector block-distributed over g = 2^k devices (device q holds indices [q n/g, (q+1) n/g)), the
// first k DIF stages pair elements n/2, n/4, ..., n/g apart, i.e. equal local offsets on different devices;
// afterwards every block is an independent size-n/g transform (local k_ntt_pass).  One thread runs the k
// cross stages for one local offset l in registers, reading the g values D[q][l] through per-block base
// pointers.  Those pointers are either g planes of a gathered buffer (mpc_cuda_ntt_cross_stage_dev: the
// caller did the all-to-all, e.g. NCCL across processes) or the blocks themselves on their home devices
// (mpc_cuda_ntt_fr_sharded_dev: one process, NVLink peer loads and stores inside the kernel — the exchange is
// fused into the butterflies and no separate transpose pass exists).
// kind fft/coset_fft: (22^j scaling,) DIF stages 0..k-1.  kind ifft/coset_ifft: the exact inverse (stages
// k-1..0 of t = hi w^-1, lo' = lo + t, hi' = lo - t, then g^-1 and, for the coset, 22^-i).  Output layout of the
// forward transform: device r, local m holds X[m g + bitrev_k(r)] (the usual transposed order of a four-step
// NTT); the inverse consumes that layout.
constexpr int MAX_LOG_G = 3;

struct CrossArgs {
    Fr* blk[1 << MAX_LOG_G];       // element (q, l) lives at blk[q][l - base]
    const Fr* tw[MAX_LOG_G];       // tw[s] = forward table of domain log_n - s
    const Fr *g_lo, *g_hi;         // 22^(+-i) two-level tables (coset kinds), or nullptr
    const Fr* scale;               // g^-1 (inverse kinds)
    size_t l0, len, base;          // local offsets [l0, l0 + len); blk pointers are biased by `base`
    uint32_t log_n, log_g, inverse;
};

template <int LOG_G>
__global__ void __launch_bounds__(128) k_ntt_cross(CrossArgs a) {
    constexpr int G = 1 << LOG_G;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.len) return;
    const size_t n = (size_t)1 << a.log_n, blk = n >> LOG_G, l = a.l0 + t;
    const bool inverse = a.inverse != 0;
    Fr x[G];
#pragma unroll
    for (int q = 0; q < G; q++) x[q] = load_fe(a.blk[q] + (l - a.base));      // all (peer) loads in flight first
#pragma unroll
    for (int q = 0; q < G; q++) {
        if (!inverse && a.g_lo) {
            size_t j = (size_t)q * blk + l;
            x[q] = mul(x[q], load_fe_ro(a.g_lo + (j & ((1u << COSET_LO_BITS) - 1))));
            if (a.log_n > COSET_LO_BITS) x[q] = mul(x[q], load_fe_ro(a.g_hi + (j >> COSET_LO_BITS)));
        }
    }
    if (!inverse) {
#pragma unroll
        for (int s = 0; s < LOG_G; s++) {
            const int dist = G >> (s + 1);
            const size_t half = n >> (s + 1);
#pragma unroll
            for (int q = 0; q < G; q++) {
                if (q & dist) continue;
                size_t k = ((size_t)q * blk + l) & (half - 1);
                bfly(x[q], x[q + dist], a.tw[s], k, a.log_n - s, false);
            }
        }
    } else {
#pragma unroll
        for (int s = LOG_G - 1; s >= 0; s--) {
            const int dist = G >> (s + 1);
            const size_t half = n >> (s + 1);
            const uint32_t lg = a.log_n - s;
#pragma unroll
            for (int q = 0; q < G; q++) {
                if (q & dist) continue;
                size_t k = ((size_t)q * blk + l) & (half - 1);
                // t = hi * w^-k with w^-k = -w^(2^(lg-1) - k) for k != 0
                Fr tv = x[q + dist];
                if (lg > 1) tv = mul(tv, load_fe_ro(a.tw[s] + (k ? ((size_t)1 << (lg - 1)) - k : 0)));
                if (k) { x[q + dist] = add(x[q], tv); x[q] = sub(x[q], tv); }
                else { x[q + dist] = sub(x[q], tv); x[q] = add(x[q], tv); }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < G; q++) {
        if (inverse) {
            x[q] = mul(x[q], load_fe_ro(a.scale));
            if (a.g_lo) {
                size_t i = (size_t)q * blk + l;
                x[q] = mul(x[q], load_fe_ro(a.g_lo + (i & ((1u << COSET_LO_BITS) - 1))));
                if (a.log_n > COSET_LO_BITS) x[q] = mul(x[q], load_fe_ro(a.g_hi + (i >> COSET_LO_BITS)));
            }
        }
        store_fe(a.blk[q] + (l - a.base), x[q]);
    }
}

// blk[q]: where element (q, l0) of the call lives; `base` = the local offset blk[q][0] corresponds to
int32_t ntt_cross_launch(Fr* const* blk, size_t base, uint32_t log_n, uint32_t log_g, size_t l0, size_t len, uint32_t kind,
                         cudaStream_t s) {
    MPC_ARG_CHECK(kind <= MPC_CUDA_NTT_COSET_IFFT && log_g >= 1 && log_g <= (uint32_t)MAX_LOG_G);
    MPC_ARG_CHECK(log_n <= MAX_LOG_N && log_n > log_g && l0 + len <= ((size_t)1 << (log_n - log_g)));
    if (len == 0) return MPC_CUDA_OK;
    const bool inverse = kind == MPC_CUDA_NTT_IFFT || kind == MPC_CUDA_NTT_COSET_IFFT;
    const bool coset = kind >= MPC_CUDA_NTT_COSET_FFT;
    DomainCache dc;
    MPC_TRY(ensure_tables(current_device_index(), log_n, coset, s, &dc));
    CrossArgs a;
    memset(&a, 0, sizeof(a));
    for (uint32_t q = 0; q < (1u << log_g); q++) {
        MPC_ARG_CHECK(blk[q] != nullptr);
        a.blk[q] = blk[q];
    }
    for (uint32_t st = 0; st < log_g; st++) a.tw[st] = dc.tw[log_n - st];
    if (coset) { a.g_lo = inverse ? dc.gi_lo : dc.g_lo; a.g_hi = inverse ? dc.gi_hi : dc.g_hi; }
    a.scale = dc.consts + 48 + log_g;
    a.l0 = l0; a.len = len; a.base = base; a.log_n = log_n; a.log_g = log_g; a.inverse = inverse;
    unsigned blocks = (unsigned)((len + 127) / 128);
    switch (log_g) {
        case 1: k_ntt_cross<1><<<blocks, 128, 0, s>>>(a); break;
        case 2: k_ntt_cross<2><<<blocks, 128, 0, s>>>(a); break;
        default: k_ntt_cross<3><<<blocks, 128, 0, s>>>(a); break;
    }
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t ntt_cross_dev(Fr* data, uint32_t log_n, uint32_t log_g, size_t l0, size_t len, uint32_t kind, cudaStream_t s) {
    MPC_ARG_CHECK(log_g >= 1 && log_g <= (uint32_t)MAX_LOG_G);
    if (len == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(data != nullptr);
    Fr* blk[1 << MAX_LOG_G];
    for (uint32_t q = 0; q < (1u << log_g); q++) blk[q] = data + (size_t)q * len;
    return ntt_cross_launch(blk, l0, log_n, log_g, l0, len, kind, s);
}

// ---- one process, g devices: the whole sharded transform -------------------------------------------------
// Cross-device barrier: every stream waits until all streams reached this point.  Two phases through the first
// device's stream (2g waits instead of g^2), events cached per host thread and device (creating and destroying
// them per call costs more than the barrier).  Reusing an event is safe: a later record can only be issued after
// the host enqueued every wait on the earlier one, and a wait captures the record that precedes it.
struct EvCache {
    cudaEvent_t arrive[64] = {}, release = nullptr;
    int release_dev = -1;
};
thread_local EvCache t_ev;

// mode: BAR_FULL everyone waits for everyone; BAR_JOIN only the first stream waits for the others (end of a
// captured graph); BAR_FORK only the others wait for the first stream (start of a captured graph)
enum { BAR_FULL = 0, BAR_JOIN = 1, BAR_FORK = 2 };
int32_t cross_barrier(const int* dev, cudaStream_t* st, int g, int mode = BAR_FULL) {
    bool distinct = false;
    for (int q = 1; q < g; q++) distinct = distinct || st[q] != st[0];
    if (!distinct) return MPC_CUDA_OK;               // one stream: already ordered
    if (mode != BAR_FORK) {
        for (int q = 1; q < g; q++) {
            DeviceScope scope(dev[q]);
            MPC_TRY(scope.rc);
            cudaEvent_t& e = t_ev.arrive[dev[q]];
            if (!e) MPC_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            MPC_CUDA_TRY(cudaEventRecord(e, st[q]));
        }
    }
    {
        DeviceScope scope(dev[0]);
        MPC_TRY(scope.rc);
        if (mode != BAR_FORK)
            for (int q = 1; q < g; q++) MPC_CUDA_TRY(cudaStreamWaitEvent(st[0], t_ev.arrive[dev[q]], 0));
        if (mode == BAR_JOIN) return MPC_CUDA_OK;
        if (t_ev.release && t_ev.release_dev != dev[0]) { cudaEventDestroy(t_ev.release); t_ev.release = nullptr; }
        if (!t_ev.release) {
            MPC_CUDA_TRY(cudaEventCreateWithFlags(&t_ev.release, cudaEventDisableTiming));
            t_ev.release_dev = dev[0];
        }
        MPC_CUDA_TRY(cudaEventRecord(t_ev.release, st[0]));
    }
    for (int q = 1; q < g; q++) MPC_CUDA_TRY(cudaStreamWaitEvent(st[q], t_ev.release, 0));
    return MPC_CUDA_OK;
}

// the launches of one sharded transform.  captured = true: the streams are being captured into a CUDA graph from
// the first stream: fork at the start, join at the end, no allocation (tmp[q] = ping-pong buffer on device q).
int32_t ntt_sharded_body(Fr* const* blocks, const int* dev, cudaStream_t* st, int g, uint32_t log_n, uint32_t log_g,
                         uint32_t kind, Fr* const* tmp, bool captured) {
    const bool inverse = kind == MPC_CUDA_NTT_IFFT || kind == MPC_CUDA_NTT_COSET_IFFT;
    const size_t m = (size_t)1 << (log_n - log_g), slice = m >> log_g;
    if (captured) MPC_TRY(cross_barrier(dev, st, g, BAR_FORK));
    if (inverse) {
        for (int q = 0; q < g; q++) {
            DeviceScope scope(dev[q]);
            MPC_TRY(scope.rc);
            MPC_TRY(ntt_dev(blocks[q], log_n - log_g, MPC_CUDA_NTT_IFFT, 1, st[q], tmp ? tmp[q] : nullptr));
        }
    }
    if (!captured || inverse) MPC_TRY(cross_barrier(dev, st, g));      // every block is ready to be read by every device
    for (int r = 0; r < g; r++) {
        DeviceScope scope(dev[r]);
        MPC_TRY(scope.rc);
        MPC_TRY(ntt_cross_launch(blocks, 0, log_n, log_g, (size_t)r * slice, slice, kind, st[r]));
    }
    // every block received the stores of every device
    MPC_TRY(cross_barrier(dev, st, g, inverse && captured ? BAR_JOIN : BAR_FULL));
    if (!inverse) {
        for (int q = 0; q < g; q++) {
            DeviceScope scope(dev[q]);
            MPC_TRY(scope.rc);
            MPC_TRY(ntt_dev(blocks[q], log_n - log_g, MPC_CUDA_NTT_FFT, 1, st[q], tmp ? tmp[q] : nullptr));
        }
        // the first device's stream now orders after the whole job
        MPC_TRY(cross_barrier(dev, st, g, captured ? BAR_JOIN : BAR_FULL));
    }
    return MPC_CUDA_OK;
}

// A sharded transform is ~170 driver calls from one host thread (8 devices x (cross stage + 3 passes) + the event
// barriers).  Repeated transforms of the same blocks (witness_map runs seven per proof over the same buffers) can
// be replayed from a CUDA graph captured across the devices' streams (option ntt_graph = 1): first call direct
// (builds tables, sets attributes), second call captured, later calls one cudaGraphLaunch between two small real
// barriers that order it against the other streams' earlier / later work.  Off by default: on 8 B200s the replay
// (0.680 ms at 2^24) is no faster than the direct launches (0.686 ms) — the GPUs, not the host, are the limit.
struct ShardKey {
    Fr* blocks[1 << MAX_LOG_G];
    int dev[1 << MAX_LOG_G];
    uint32_t log_n, log_g, kind;
    bool operator==(const ShardKey& o) const { return memcmp(this, &o, sizeof(ShardKey)) == 0; }
};
struct ShardGraph {
    ShardKey key;
    uint32_t uses = 0;
    bool failed = false;
    cudaGraphExec_t exec = nullptr;
    Fr* tmp[1 << MAX_LOG_G] = {};
};
thread_local std::vector<ShardGraph> t_shard_graphs;

void shard_graph_release(ShardGraph& e) {
    if (e.exec) cudaGraphExecDestroy(e.exec);
    for (int q = 0; q < (1 << MAX_LOG_G); q++) {
        if (e.tmp[q]) {
            DeviceScope scope(e.key.dev[q]);
            cudaFree(e.tmp[q]);
        }
    }
}

int32_t ntt_sharded_dev(Fr* const* blocks, const int32_t* dev_index, uint32_t log_n, uint32_t log_g, uint32_t kind) {
    MPC_ARG_CHECK(blocks && kind <= MPC_CUDA_NTT_COSET_IFFT && log_g <= (uint32_t)MAX_LOG_G && log_n <= MAX_LOG_N);
    const int g = 1 << log_g;
    if (log_g == 0) {
        DeviceScope scope(dev_index ? dev_index[0] : current_device_index());
        MPC_TRY(scope.rc);
        return ntt_dev(blocks[0], log_n, kind, 1, scope.s);
    }
    // every device runs log_g cross stages on a slice of m / g local offsets, so the block must hold >= g elements
    MPC_ARG_CHECK(log_n >= 2 * log_g);
    int dev[1 << MAX_LOG_G];
    cudaStream_t st[1 << MAX_LOG_G];
    bool distinct = false;
    for (int q = 0; q < g; q++) {
        dev[q] = dev_index ? dev_index[q] : q;
        MPC_ARG_CHECK(dev[q] >= 0 && dev[q] < device_list_size() && blocks[q]);
        if (dev[q] != dev[0]) distinct = true;
    }
    if (distinct) MPC_TRY(enable_peer_access());
    for (int q = 0; q < g; q++) {
        DeviceScope scope(dev[q]);
        MPC_TRY(scope.rc);
        st[q] = scope.s;
    }
    const int64_t gopt = g_opt_ntt_graph.load(std::memory_order_relaxed);
    // measured on 2 and 8 B200s (tools/bench_sharded_ntt.py): direct launches and graph replay take the same time (the
    // transform is bound by the GPUs, not by the ~170 driver calls), so replay is opt-in
    const bool want_graph = gopt == 1 && !g_opt_profile.load(std::memory_order_relaxed);
    (void)distinct;
    if (!want_graph) return ntt_sharded_body(blocks, dev, st, g, log_n, log_g, kind, nullptr, false);

    ShardKey key;
    memset(&key, 0, sizeof(key));
    for (int q = 0; q < g; q++) { key.blocks[q] = blocks[q]; key.dev[q] = dev[q]; }
    key.log_n = log_n; key.log_g = log_g; key.kind = kind;
    ShardGraph* e = nullptr;
    for (ShardGraph& c : t_shard_graphs)
        if (c.key == key) e = &c;
    if (!e) {
        if (t_shard_graphs.size() >= 32) {           // bounded cache: drop the oldest entry
            shard_graph_release(t_shard_graphs.front());
            t_shard_graphs.erase(t_shard_graphs.begin());
        }
        t_shard_graphs.emplace_back();
        e = &t_shard_graphs.back();
        e->key = key;
    }
    e->uses++;
    if (e->failed || e->uses == 1) return ntt_sharded_body(blocks, dev, st, g, log_n, log_g, kind, nullptr, false);
    if (!e->exec) {
        // capture: ping-pong buffers first (no allocation inside the graph), then the launches
        const size_t m = (size_t)1 << (log_n - log_g);
        bool ok = true;
        if (log_n - log_g > (uint32_t)NTT_MAX_DEG) {
            for (int q = 0; q < g && ok; q++) {
                DeviceScope scope(dev[q]);
                ok = scope.rc == MPC_CUDA_OK && cudaMalloc((void**)&e->tmp[q], m * sizeof(Fr)) == cudaSuccess;
            }
        }
        cudaGraph_t graph = nullptr;
        {
            DeviceScope scope(dev[0]);
            ok = ok && scope.rc == MPC_CUDA_OK && cudaStreamBeginCapture(st[0], cudaStreamCaptureModeRelaxed) == cudaSuccess;
        }
        if (ok) {
            int32_t rc = ntt_sharded_body(blocks, dev, st, g, log_n, log_g, kind, e->tmp[0] ? e->tmp : nullptr, true);
            DeviceScope scope(dev[0]);
            cudaError_t ce = cudaStreamEndCapture(st[0], &graph);
            ok = rc == MPC_CUDA_OK && ce == cudaSuccess && graph != nullptr;
            if (ok) ok = cudaGraphInstantiate(&e->exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
        }
        if (!ok) {                                   // graphs are an optimisation: fall back to direct launches for good
            cudaGetLastError();
            e->failed = true;
            e->exec = nullptr;
            return ntt_sharded_body(blocks, dev, st, g, log_n, log_g, kind, nullptr, false);
        }
    }
    // the graph runs on the first stream: order it after the other streams' earlier work and them after it
    MPC_TRY(cross_barrier(dev, st, g, BAR_JOIN));
    {
        DeviceScope scope(dev[0]);
        MPC_TRY(scope.rc);
        MPC_CUDA_TRY(cudaGraphLaunch(e->exec, st[0]));
        count_launch();
    }
    return cross_barrier(dev, st, g, BAR_FORK);
}

// natural block order <-> the transposed order the forward kinds produce: out[r][mm] = X[mm g + bitrev(r)].
// to_transposed: device r gathers its strided elements from every natural block (32-byte peer reads);
// from_transposed: device q gathers natural index q m + j from owner bitrev(j mod g), local j / g.
struct ReorderArgs {
    const Fr* src[1 << MAX_LOG_G];
    Fr* dst;
    size_t m;
    uint32_t log_g, me, to_transposed;
};

__global__ void __launch_bounds__(256) k_ntt_reorder(ReorderArgs a) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.m) return;
    const uint32_t g = 1u << a.log_g;
    if (a.to_transposed) {
        size_t nat = i * g + bitrev(a.me, a.log_g);                   // natural index of out[me][i]
        store_fe(a.dst + i, load_fe(a.src[nat / a.m] + nat % a.m));
    } else {
        size_t nat = (size_t)a.me * a.m + i;                           // natural index of out[me][i]
        store_fe(a.dst + i, load_fe(a.src[bitrev((uint32_t)(nat % g), a.log_g)] + nat / g));
    }
}

int32_t ntt_reorder_sharded(Fr* const* in, Fr* const* out, const int32_t* dev_index, uint32_t log_n, uint32_t log_g,
                            uint32_t to_transposed) {
    MPC_ARG_CHECK(in && out && log_g >= 1 && log_g <= (uint32_t)MAX_LOG_G && log_n >= log_g && log_n <= MAX_LOG_N);
    const int g = 1 << log_g;
    int dev[1 << MAX_LOG_G];
    cudaStream_t st[1 << MAX_LOG_G];
    bool distinct = false;
    for (int q = 0; q < g; q++) {
        dev[q] = dev_index ? dev_index[q] : q;
        MPC_ARG_CHECK(dev[q] >= 0 && dev[q] < device_list_size() && in[q] && out[q] && in[q] != out[q]);
        if (dev[q] != dev[0]) distinct = true;
    }
    if (distinct) MPC_TRY(enable_peer_access());
    for (int q = 0; q < g; q++) {
        DeviceScope scope(dev[q]);
        MPC_TRY(scope.rc);
        st[q] = scope.s;
    }
    MPC_TRY(cross_barrier(dev, st, g));
    const size_t m = (size_t)1 << (log_n - log_g);
    for (int q = 0; q < g; q++) {
        DeviceScope scope(dev[q]);
        MPC_TRY(scope.rc);
        ReorderArgs a;
        for (int r = 0; r < g; r++) a.src[r] = in[r];
        a.dst = out[q];
        a.m = m; a.log_g = log_g; a.me = (uint32_t)q; a.to_transposed = to_transposed;
        k_ntt_reorder<<<(unsigned)((m + 255) / 256), 256, 0, st[q]>>>(a);
        MPC_KERNEL_CHECK();
    }
    return cross_barrier(dev, st, g);
}

}  // namespace

extern "C" {

int32_t mpc_cuda_ntt_fr_dev(uint64_t* data, uint32_t log_n, uint32_t kind, uint32_t batch, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    return ntt_dev((Fr*)data, log_n, kind, batch, pick_stream(stream, s));
}

int32_t mpc_cuda_ntt_cross_stage_dev(uint64_t* data, uint32_t log_n, uint32_t log_g, size_t slice_offset,
                                     size_t slice_len, uint32_t kind, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    return ntt_cross_dev((Fr*)data, log_n, log_g, slice_offset, slice_len, kind, pick_stream(stream, s));
}

int32_t mpc_cuda_ntt_fr_sharded_dev(uint64_t* const* blocks, const int32_t* dev_index, uint32_t log_n, uint32_t log_g,
                                    uint32_t kind) {
    MPC_TRY(enter(nullptr));
    return ntt_sharded_dev((Fr* const*)blocks, dev_index, log_n, log_g, kind);
}

int32_t mpc_cuda_ntt_reorder_sharded_dev(uint64_t* const* in, uint64_t* const* out, const int32_t* dev_index,
                                         uint32_t log_n, uint32_t log_g, uint32_t to_transposed) {
    MPC_TRY(enter(nullptr));
    return ntt_reorder_sharded((Fr* const*)in, (Fr* const*)out, dev_index, log_n, log_g, to_transposed);
}

// host vector, in-order in and out like mpc_cuda_ntt_fr, computed on 2^log_g devices: block q goes to device q
// over its own PCIe link, the transform runs sharded, the blocks come back in natural order
int32_t mpc_cuda_ntt_fr_sharded(uint64_t* data, uint32_t log_n, uint32_t kind, uint32_t log_g) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(data && kind <= MPC_CUDA_NTT_COSET_IFFT && log_n <= MAX_LOG_N && log_g <= (uint32_t)MAX_LOG_G);
    if (log_g == 0) return mpc_cuda_ntt_fr(data, log_n, kind, 1);
    MPC_ARG_CHECK(log_n >= 2 * log_g && (1 << log_g) <= device_list_size());
    const int g = 1 << log_g;
    const size_t m = (size_t)1 << (log_n - log_g);
    const bool inverse = kind == MPC_CUDA_NTT_IFFT || kind == MPC_CUDA_NTT_COSET_IFFT;
    MPC_TRY(enable_peer_access());
    Scratch sa[1 << MAX_LOG_G], sb[1 << MAX_LOG_G];
    Fr *a[1 << MAX_LOG_G], *b[1 << MAX_LOG_G];
    cudaStream_t st[1 << MAX_LOG_G];
    Fr* host = (Fr*)data;
    for (int q = 0; q < g; q++) {
        DeviceScope scope(q);
        MPC_TRY(scope.rc);
        st[q] = scope.s;
        MPC_TRY(sa[q].alloc(&a[q], m, st[q]));
        MPC_TRY(sb[q].alloc(&b[q], m, st[q]));
        MPC_CUDA_TRY(cudaMemcpyAsync(a[q], host + (size_t)q * m, m * sizeof(Fr), cudaMemcpyHostToDevice, st[q]));
    }
    Fr** work = a;
    if (inverse) {          // the inverse kinds consume the transposed order
        MPC_TRY(ntt_reorder_sharded(a, b, nullptr, log_n, log_g, 1));
        work = b;
    }
    MPC_TRY(ntt_sharded_dev(work, nullptr, log_n, log_g, kind));
    if (!inverse) {         // the forward kinds produce it
        MPC_TRY(ntt_reorder_sharded(a, b, nullptr, log_n, log_g, 0));
        work = b;
    }
    for (int q = 0; q < g; q++) {
        DeviceScope scope(q);
        MPC_TRY(scope.rc);
        MPC_CUDA_TRY(cudaMemcpyAsync(host + (size_t)q * m, work[q], m * sizeof(Fr), cudaMemcpyDeviceToHost, st[q]));
    }
    for (int q = 0; q < g; q++) {
        DeviceScope scope(q);
        MPC_TRY(scope.rc);
        MPC_CUDA_TRY(cudaStreamSynchronize(st[q]));
    }
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_ntt_fr(uint64_t* data, uint32_t log_n, uint32_t kind, uint32_t batch) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(log_n <= MAX_LOG_N);
    if (batch == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(data != nullptr);
    size_t count = ((size_t)1 << log_n) * batch;
    Scratch sd;
    Fr* d;
    MPC_TRY(sd.alloc(&d, count, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(d, data, count * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_TRY(ntt_dev(d, log_n, kind, batch, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(data, d, count * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_divide_by_vanishing_on_coset_dev(uint64_t* data, uint32_t log_n, void* stream) {
    cudaStream_t s0;
    MPC_TRY(enter(&s0));
    cudaStream_t s = pick_stream(stream, s0);
    MPC_ARG_CHECK(data != nullptr && log_n <= MAX_LOG_N);
    DomainCache dc;
    MPC_TRY(ensure_tables(current_device_index(), 0, false, s, &dc));
    size_t n = (size_t)1 << log_n;
    k_mul_const_dev<<<grid_for(n, 256, 8), 256, 0, s>>>((Fr*)data, dc.consts + 96 + log_n, n);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_divide_by_vanishing_on_coset(uint64_t* data, uint32_t log_n) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(data != nullptr && log_n <= MAX_LOG_N);
    size_t n = (size_t)1 << log_n;
    Scratch sd;
    Fr* d;
    MPC_TRY(sd.alloc(&d, n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(d, data, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_TRY(mpc_cuda_divide_by_vanishing_on_coset_dev((uint64_t*)d, log_n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(data, d, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

}  // extern "C"
