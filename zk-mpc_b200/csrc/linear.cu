// The linear steps either side of the NTT / Beaver / MSM kernels, all on local share
// values (every one of them is linear in the shares, so a party's output is its share of the result):
//
//   f2  public sparse matrix x share vector   evaluate_constraint, src/groth16.rs:205-234 (used at :263-270,
//       289-293); Marlin's inner products arkworks/marlin/src/ahp/prover.rs:258-278
//   f3  wire bytes of a Vec<Fr> as MpcSerNet::broadcast sends them (mpc-algebra/src/channel.rs:12-28):
//       u64 LE length + 32 LE bytes of the canonical integer per element
//       (arkworks/algebra/serialize/src/lib.rs:263-272, ff/src/fields/macros.rs:1-110), fused with the Beaver
//       mask on the way out and with the open-sum on the way in (share/additive.rs:125-131)
//   f4  p(x) / (x - z) and p(z) for a public point z: univariate_div_qr on shares
//       (mpc-algebra/src/wire/field.rs:1007-1065 -> share/additive.rs:154-162 ->
//       poly/src/polynomial/univariate/mod.rs:133-172) as KZG10::open uses it
//       (poly-commit/src/kzg10/mod.rs:241-258), evaluation = dense.rs:71-75 (Horner)
#include <mutex>
#include <unordered_map>

#include "common.cuh"

using namespace mpc;

namespace {

constexpr int LIN_THREADS = 256;

// ---------------------------------------------------------------------------------------------- f2: CSR x vector
struct Csr {
    uint64_t* row_ptr = nullptr;      // rows + 1
    uint32_t* col = nullptr;          // nnz
    Fr* coeff = nullptr;              // nnz, Montgomery
    size_t rows = 0, cols = 0, nnz = 0;
    int cuda_device = 0;
};
std::mutex g_csr_mu;
std::unordered_map<uint64_t, Csr> g_csr;
uint64_t g_csr_next = 1;

DEV bool is_mont_one(const Fr& c) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < Fr::N; i++) acc |= c.v[i] ^ consts::FrParams::one(i);
    return acc == 0;
}

// one thread per (row, plane): rows of an R1CS matrix hold a handful of terms, most with coefficient 1
// (added without a product, as the reference does)
__global__ void __launch_bounds__(LIN_THREADS) k_spmv_rows(const uint64_t* __restrict__ row_ptr,
                                                          const uint32_t* __restrict__ col,
                                                          const Fr* __restrict__ coeff, size_t rows,
                                                          const Fr* __restrict__ x, size_t x_stride, Fr* __restrict__ out,
                                                          size_t out_stride, uint32_t planes) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= rows * planes) return;
    size_t r = id % rows, p = id / rows;
    const Fr* xv = x + p * x_stride;
    Fr sum = Fr::zero();
    for (uint64_t k = row_ptr[r], e = row_ptr[r + 1]; k < e; k++) {
        Fr c = load_fe_ro(coeff + k), v = load_fe_ro(xv + col[k]);
        sum = add(sum, is_mont_one(c) ? v : mul(v, c));
    }
    store_fe(out + p * out_stride + r, sum);
}

// one warp per (row, plane) for matrices with long rows: lanes stride over the terms, shuffle-tree sum
__global__ void __launch_bounds__(LIN_THREADS) k_spmv_warp(const uint64_t* __restrict__ row_ptr,
                                                          const uint32_t* __restrict__ col,
                                                          const Fr* __restrict__ coeff, size_t rows,
                                                          const Fr* __restrict__ x, size_t x_stride, Fr* __restrict__ out,
                                                          size_t out_stride, uint32_t planes) {
    size_t id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (id >= rows * planes) return;              // whole warps leave together
    size_t r = id % rows, p = id / rows;
    const Fr* xv = x + p * x_stride;
    Fr sum = Fr::zero();
    for (uint64_t k = row_ptr[r] + lane, e = row_ptr[r + 1]; k < e; k += 32) {
        Fr c = load_fe_ro(coeff + k), v = load_fe_ro(xv + col[k]);
        sum = add(sum, is_mont_one(c) ? v : mul(v, c));
    }
    for (int off = 16; off > 0; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < Fr::N; i++) o.v[i] = __shfl_down_sync(0xffffffffu, sum.v[i], off);
        sum = add(sum, o);
    }
    if (lane == 0) store_fe(out + p * out_stride + r, sum);
}

int32_t csr_find(uint64_t handle, Csr* out) {
    std::lock_guard<std::mutex> lk(g_csr_mu);
    auto it = g_csr.find(handle);
    if (it == g_csr.end()) {
        set_error("unknown CSR matrix handle %llu", (unsigned long long)handle);
        return MPC_CUDA_ERR_HANDLE;
    }
    if (it->second.cuda_device != current_device_info()->cuda_device) {
        set_error("CSR matrix %llu lives on CUDA device %d, calling thread uses %d", (unsigned long long)handle,
                  it->second.cuda_device, current_device_info()->cuda_device);
        return MPC_CUDA_ERR_HANDLE;
    }
    *out = it->second;
    return MPC_CUDA_OK;
}

int32_t spmv_launch(const Csr& m, const Fr* x, size_t x_stride, uint32_t planes, Fr* out, size_t out_stride, cudaStream_t s) {
    if (m.rows == 0 || planes == 0) return MPC_CUDA_OK;
    size_t work = m.rows * planes;
    if (m.nnz > 64 * m.rows) {
        k_spmv_warp<<<(unsigned)((work * 32 + LIN_THREADS - 1) / LIN_THREADS), LIN_THREADS, 0, s>>>(
            m.row_ptr, m.col, m.coeff, m.rows, x, x_stride, out, out_stride, planes);
    } else {
        k_spmv_rows<<<(unsigned)((work + LIN_THREADS - 1) / LIN_THREADS), LIN_THREADS, 0, s>>>(
            m.row_ptr, m.col, m.coeff, m.rows, x, x_stride, out, out_stride, planes);
    }
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------- f3: wire bytes
// The payload is 8 header bytes followed by 32-byte elements, so elements are 8- but not 16-byte aligned:
// 64-bit accesses.
DEV void store_canonical(uint8_t* dst, const Fr& canon) {
    uint2* q = reinterpret_cast<uint2*>(dst);
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = make_uint2(canon.v[2 * i], canon.v[2 * i + 1]);
}
DEV Fr load_canonical(const uint8_t* src) {
    Fr r;
    const uint2* q = reinterpret_cast<const uint2*>(src);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint2 t = q[i];
        r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y;
    }
    return r;
}
DEV bool below_modulus(const Fr& a) {
    ptx::sub_cc(a.v[0], consts::FrParams::mod(0));
#pragma unroll
    for (int i = 1; i < Fr::N; i++) ptx::subc_cc(a.v[i], consts::FrParams::mod(i));
    return ptx::subc(0, 0) != 0;
}

// out = [n as u64 LE | canonical(s[i] (+ x[i]))]
template <bool MASK>
__global__ void __launch_bounds__(LIN_THREADS) k_serialize(const Fr* __restrict__ s, const Fr* __restrict__ x, size_t n,
                                                          uint8_t* __restrict__ out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (first == 0) *reinterpret_cast<uint2*>(out) = make_uint2((uint32_t)n, (uint32_t)((uint64_t)n >> 32));
    for (size_t i = first; i < n; i += stride) {
        Fr v = load_fe_ro(s + i);
        if (MASK) v = add(v, load_fe_ro(x + i));
        store_canonical(out + 8 + 32 * i, from_mont(v));
    }
}

// out[i] = to_mont(sum over the P payloads of element i); flags[0] = 1 + smallest index of an element >= r,
// flags[1] = 1 if a length prefix differs from n
__global__ void __launch_bounds__(LIN_THREADS) k_deserialize_sum(const uint8_t* __restrict__ in, uint32_t P, size_t n,
                                                                Fr* __restrict__ out, unsigned long long* __restrict__ flags) {
    const size_t payload = 8 + 32 * n;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (first < P) {
        uint2 h = *reinterpret_cast<const uint2*>(in + first * payload);
        if ((((uint64_t)h.y << 32) | h.x) != (uint64_t)n) flags[1] = 1;
    }
    for (size_t i = first; i < n; i += stride) {
        Fr acc = Fr::zero();
        for (uint32_t p = 0; p < P; p++) {
            Fr v = load_canonical(in + p * payload + 8 + 32 * i);
            if (!below_modulus(v)) atomicMin(flags, (unsigned long long)i + 1);
            acc = add(acc, v);                     // canonical integers mod r add like field elements
        }
        store_fe(out + i, to_mont(acc));
    }
}

// ---------------------------------------------------------------------------------------------- f4: p / (x - z)
// S_i = sum_{j >= i} p_j z^(j-i) satisfies S_i = p_i + z S_(i+1); quotient q_(i-1) = S_i (i >= 1), remainder
// p(z) = S_0.  A three-level blocked scan of that recurrence: threads own L consecutive coefficients, CTAs 256
// chunks, one CTA the block heads.
constexpr int DIV_CHUNKS = 256;

DEV Fr fr_pow(Fr b, uint64_t e) {
    Fr acc = Fr::one();
    while (e) {
        if (e & 1) acc = mul(acc, b);
        b = sqr(b);
        e >>= 1;
    }
    return acc;
}

// t[c] = sum_{j in chunk c} p_j z^(j - lo_c)
__global__ void __launch_bounds__(LIN_THREADS) k_div_chunks(const Fr* __restrict__ p, size_t n, uint32_t L, Fr z,
                                                           Fr* __restrict__ t, size_t chunks) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chunks) return;
    size_t lo = c * L, hi = lo + L < n ? lo + L : n;
    Fr acc = Fr::zero();
    for (size_t j = hi; j-- > lo;) acc = add(mul(acc, z), load_fe_ro(p + j));
    store_fe(t + c, acc);
}

// in-place suffix recurrence inside each CTA's DIV_CHUNKS entries with multiplier m (carry-in 0 from the right):
// u[c] = sum_{d >= 0, c + d in block} m^d v[c + d]; heads[b] = u[first entry of block b]
__global__ void __launch_bounds__(DIV_CHUNKS) k_div_scan(Fr* __restrict__ v, size_t count, const Fr* __restrict__ m,
                                                        Fr* __restrict__ heads) {
    __shared__ uint32_t sm[DIV_CHUNKS * Fr::N];
    Fr* buf = reinterpret_cast<Fr*>(sm);
    const uint32_t t = threadIdx.x;
    size_t c = (size_t)blockIdx.x * DIV_CHUNKS + t;
    Fr u = c < count ? load_fe(v + c) : Fr::zero();
    Fr mp = load_fe_ro(m);                         // m^(2^k)
    for (uint32_t off = 1; off < DIV_CHUNKS; off <<= 1) {
        buf[t] = u;
        __syncthreads();
        if (t + off < DIV_CHUNKS) u = add(u, mul(mp, buf[t + off]));
        __syncthreads();
        mp = sqr(mp);
    }
    if (c < count) store_fe(v + c, u);
    if (t == 0 && heads) store_fe(heads + blockIdx.x, u);
}

// add the carry of everything right of the block: u[c] += m^(entries from c to the block end) * upper[b + 1],
// upper[] holding the finished values of the block heads
__global__ void __launch_bounds__(DIV_CHUNKS) k_div_fix(Fr* __restrict__ v, size_t count, const Fr* __restrict__ m,
                                                        const Fr* __restrict__ upper, size_t nblocks) {
    size_t b = blockIdx.x, c = b * DIV_CHUNKS + threadIdx.x;
    if (c >= count || b + 1 >= nblocks) return;
    Fr carry = load_fe(upper + b + 1);
    Fr scale = fr_pow(load_fe_ro(m), (uint64_t)(DIV_CHUNKS - threadIdx.x));
    store_fe(v + c, add(load_fe(v + c), mul(scale, carry)));
}

// chunk c restarts its Horner loop from the finished S of the next chunk and emits q and, for chunk 0, p(z)
__global__ void __launch_bounds__(LIN_THREADS) k_div_emit(const Fr* __restrict__ p, size_t n, uint32_t L, Fr z,
                                                         const Fr* __restrict__ t, size_t chunks, Fr* __restrict__ q,
                                                         Fr* __restrict__ rem) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chunks) return;
    size_t lo = c * L, hi = lo + L < n ? lo + L : n;
    Fr acc = c + 1 < chunks ? load_fe(t + c + 1) : Fr::zero();
    for (size_t j = hi; j-- > lo;) {
        acc = add(mul(acc, z), load_fe_ro(p + j));
        if (j >= 1) { if (q) store_fe(q + j - 1, acc); }
        else store_fe(rem, acc);
    }
}

// single thread: out[0] = z^L (multiplier between chunks), out[k] = out[k-1]^DIV_CHUNKS (between level-k blocks)
__global__ void k_div_powers(Fr z, uint32_t L, uint32_t levels, Fr* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    Fr a = fr_pow(z, L);
    store_fe(out, a);
    for (uint32_t k = 1; k < levels; k++) {
        a = fr_pow(a, DIV_CHUNKS);
        store_fe(out + k, a);
    }
}

constexpr int DIV_MAX_LEVELS = 4;

// q (n - 1 elements, may be null) and rem (1 element) on the device; q must not alias p
int32_t poly_div_linear_dev(const Fr* p, size_t n, const uint64_t* z_host, Fr* q, Fr* rem, cudaStream_t s) {
    MPC_ARG_CHECK(n >= 1 && n <= ((size_t)1 << 30) && p && z_host && rem && (const Fr*)q != p);
    Fr z;
    memcpy(z.v, z_host, sizeof(Fr));
    const uint32_t L = 16;
    size_t count[DIV_MAX_LEVELS];
    uint32_t levels = 0;
    for (size_t c = (n + L - 1) / L;; c = (c + DIV_CHUNKS - 1) / DIV_CHUNKS) {
        MPC_ARG_CHECK(levels < DIV_MAX_LEVELS);
        count[levels++] = c;
        if (c <= DIV_CHUNKS) break;
    }
    size_t total = 0;
    for (uint32_t k = 0; k < levels; k++) total += count[k];
    Scratch sv, sm;
    Fr *v, *mult;
    MPC_TRY(sv.alloc(&v, total, s));
    MPC_TRY(sm.alloc(&mult, DIV_MAX_LEVELS, s));
    Fr* lvl[DIV_MAX_LEVELS];
    lvl[0] = v;
    for (uint32_t k = 1; k < levels; k++) lvl[k] = lvl[k - 1] + count[k - 1];
    k_div_powers<<<1, 32, 0, s>>>(z, L, levels, mult);
    MPC_KERNEL_CHECK();
    k_div_chunks<<<(unsigned)((count[0] + LIN_THREADS - 1) / LIN_THREADS), LIN_THREADS, 0, s>>>(p, n, L, z, lvl[0], count[0]);
    MPC_KERNEL_CHECK();
    for (uint32_t k = 0; k < levels; k++) {
        unsigned blocks = (unsigned)((count[k] + DIV_CHUNKS - 1) / DIV_CHUNKS);
        k_div_scan<<<blocks, DIV_CHUNKS, 0, s>>>(lvl[k], count[k], mult + k, k + 1 < levels ? lvl[k + 1] : nullptr);
        MPC_KERNEL_CHECK();
    }
    for (uint32_t k = levels - 1; k-- > 0;) {
        unsigned blocks = (unsigned)((count[k] + DIV_CHUNKS - 1) / DIV_CHUNKS);
        k_div_fix<<<blocks, DIV_CHUNKS, 0, s>>>(lvl[k], count[k], mult + k, lvl[k + 1], blocks);
        MPC_KERNEL_CHECK();
    }
    k_div_emit<<<(unsigned)((count[0] + LIN_THREADS - 1) / LIN_THREADS), LIN_THREADS, 0, s>>>(p, n, L, z, lvl[0], count[0], q, rem);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

// ---------------------------------------------------------------------- f4: division by / product with x^m - 1
// p = q (x^m - 1) + r on the local values of p's n coefficients (Marlin's divide_by_vanishing_poly on shares:
// poly/src/polynomial/univariate/dense.rs:166-173 -> mod.rs:133-143 -> share/additive.rs:154-162).  With the
// coefficients laid out as K = ceil(n / m) rows of m columns, q[k][c] is the sum of column c strictly below row k
// and r[c] = p[0][c] + q[0][c]: a suffix scan down every column.  Rows are cut into `chunks` runs of R rows so a
// short divisor (the public-input domain) still fills the device: run sums, a scan over the runs, then the emit.
constexpr size_t VAN_TARGET_THREADS = (size_t)1 << 18;
constexpr size_t VAN_MAX_CHUNKS = 4096;

__global__ void __launch_bounds__(LIN_THREADS) k_van_sums(const Fr* __restrict__ p, size_t n, size_t m, size_t K, size_t R,
                                                         size_t chunks, Fr* __restrict__ sums) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * m) return;
    size_t chunk = t / m, col = t % m;
    size_t lo = chunk * R, hi = lo + R < K ? lo + R : K;
    Fr acc = Fr::zero();
    for (size_t row = lo; row < hi; row++) {
        size_t idx = row * m + col;
        if (idx < n) acc = add(acc, load_fe_ro(p + idx));
    }
    store_fe(sums + t, acc);
}

// one thread per column: sums[chunk][col] <- sum of the runs below it
__global__ void __launch_bounds__(LIN_THREADS) k_van_scan(Fr* __restrict__ sums, size_t m, size_t chunks) {
    size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m) return;
    Fr carry = Fr::zero();
    for (size_t chunk = chunks; chunk-- > 0;) {
        Fr v = load_fe(sums + chunk * m + col);
        store_fe(sums + chunk * m + col, carry);
        carry = add(carry, v);
    }
}

__global__ void __launch_bounds__(LIN_THREADS) k_van_emit(const Fr* __restrict__ p, size_t n, size_t m, size_t K, size_t R,
                                                         size_t chunks, const Fr* __restrict__ sums, Fr* __restrict__ q,
                                                         Fr* __restrict__ rem) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * m) return;
    size_t chunk = t / m, col = t % m;
    size_t lo = chunk * R, hi = lo + R < K ? lo + R : K;
    Fr carry = sums ? load_fe(sums + t) : Fr::zero();
    for (size_t row = hi; row-- > lo;) {
        size_t idx = row * m + col;
        Fr v = idx < n ? load_fe_ro(p + idx) : Fr::zero();
        if (q && idx + m < n) store_fe(q + idx, carry);
        if (row == 0) store_fe(rem + col, add(v, carry));
        carry = add(carry, v);
    }
}

// out (n + m elements) = p * (x^m - 1): the coefficients shifted up by m minus themselves (dense.rs:155-162)
__global__ void __launch_bounds__(LIN_THREADS) k_van_mul(const Fr* __restrict__ p, size_t n, size_t m, Fr* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + m) return;
    Fr hi = i >= m ? load_fe_ro(p + i - m) : Fr::zero();
    Fr lo = i < n ? load_fe_ro(p + i) : Fr::zero();
    store_fe(out + i, sub(hi, lo));
}

// q (n - m elements when n > m, may be null) and rem (m elements) on the device; neither may alias p
int32_t poly_div_vanishing_dev(const Fr* p, size_t n, size_t m, Fr* q, Fr* rem, cudaStream_t s) {
    MPC_ARG_CHECK(n >= 1 && n <= ((size_t)1 << 30) && m >= 1 && m <= ((size_t)1 << 30) && p && rem && (const Fr*)q != p &&
                  (const Fr*)rem != p);
    const size_t K = (n + m - 1) / m;
    size_t chunks = VAN_TARGET_THREADS / m;
    if (chunks > VAN_MAX_CHUNKS) chunks = VAN_MAX_CHUNKS;
    if (chunks > K) chunks = K;
    if (chunks < 1) chunks = 1;
    const size_t R = (K + chunks - 1) / chunks;
    chunks = (K + R - 1) / R;
    const unsigned grid = (unsigned)((chunks * m + LIN_THREADS - 1) / LIN_THREADS);
    Scratch ss;
    Fr* sums = nullptr;
    if (chunks > 1) {
        MPC_TRY(ss.alloc(&sums, chunks * m, s));
        k_van_sums<<<grid, LIN_THREADS, 0, s>>>(p, n, m, K, R, chunks, sums);
        MPC_KERNEL_CHECK();
        k_van_scan<<<(unsigned)((m + LIN_THREADS - 1) / LIN_THREADS), LIN_THREADS, 0, s>>>(sums, m, chunks);
        MPC_KERNEL_CHECK();
    }
    k_van_emit<<<grid, LIN_THREADS, 0, s>>>(p, n, m, K, R, chunks, sums, n > m ? q : nullptr, rem);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t poly_mul_vanishing_dev(const Fr* p, size_t n, size_t m, Fr* out, cudaStream_t s) {
    MPC_ARG_CHECK(n >= 1 && n <= ((size_t)1 << 30) && m >= 1 && m <= ((size_t)1 << 30) && p && out && (const Fr*)out != p);
    k_van_mul<<<(unsigned)((n + m + LIN_THREADS - 1) / LIN_THREADS), LIN_THREADS, 0, s>>>(p, n, m, out);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

}  // namespace

extern "C" {

// ---- f2 --------------------------------------------------------------------------------------------------
int32_t mpc_cuda_csr_register(const uint64_t* row_ptr, const uint32_t* col, const uint64_t* coeff_mont, size_t rows,
                              size_t cols, uint64_t* handle) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(row_ptr && handle && cols < ((size_t)1 << 32));
    const size_t nnz = (size_t)row_ptr[rows];
    MPC_ARG_CHECK(row_ptr[0] == 0 && (nnz == 0 || (col && coeff_mont)));
    for (size_t r = 0; r < rows; r++) MPC_ARG_CHECK(row_ptr[r] <= row_ptr[r + 1]);
    for (size_t k = 0; k < nnz; k++) MPC_ARG_CHECK(col[k] < cols);
    Csr m;
    m.rows = rows; m.cols = cols; m.nnz = nnz;
    m.cuda_device = current_device_info()->cuda_device;
    MPC_CUDA_TRY(cudaMalloc((void**)&m.row_ptr, (rows + 1) * sizeof(uint64_t)));
    MPC_CUDA_TRY(cudaMalloc((void**)&m.col, (nnz ? nnz : 1) * sizeof(uint32_t)));
    MPC_CUDA_TRY(cudaMalloc((void**)&m.coeff, (nnz ? nnz : 1) * sizeof(Fr)));
    MPC_CUDA_TRY(cudaMemcpyAsync(m.row_ptr, row_ptr, (rows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(m.col, col, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(m.coeff, coeff_mont, nnz * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> lk(g_csr_mu);
    *handle = g_csr_next++;
    g_csr[*handle] = m;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_csr_release(uint64_t handle) {
    MPC_TRY(enter(nullptr));
    Csr m;
    {
        std::lock_guard<std::mutex> lk(g_csr_mu);
        auto it = g_csr.find(handle);
        if (it == g_csr.end()) {
            set_error("unknown CSR matrix handle %llu", (unsigned long long)handle);
            return MPC_CUDA_ERR_HANDLE;
        }
        m = it->second;
        g_csr.erase(it);
    }
    int cur = 0;
    MPC_CUDA_TRY(cudaGetDevice(&cur));
    MPC_CUDA_TRY(cudaSetDevice(m.cuda_device));
    cudaFree(m.row_ptr);
    cudaFree(m.col);
    cudaFree(m.coeff);
    MPC_CUDA_TRY(cudaSetDevice(cur));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_csr_dims(uint64_t handle, size_t* rows, size_t* cols, size_t* nnz) {
    MPC_TRY(enter(nullptr));
    Csr m;
    MPC_TRY(csr_find(handle, &m));
    if (rows) *rows = m.rows;
    if (cols) *cols = m.cols;
    if (nnz) *nnz = m.nnz;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_csr_spmv_dev(uint64_t handle, const uint64_t* x_dev, size_t x_stride, uint32_t planes, uint64_t* out_dev,
                              size_t out_stride, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    Csr m;
    MPC_TRY(csr_find(handle, &m));
    MPC_ARG_CHECK(planes >= 1 && x_stride >= m.cols && out_stride >= m.rows && x_dev && out_dev);
    return spmv_launch(m, (const Fr*)x_dev, x_stride, planes, (Fr*)out_dev, out_stride, pick_stream(stream, s));
}

int32_t mpc_cuda_csr_spmv(uint64_t handle, const uint64_t* x, uint32_t planes, uint64_t* out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    Csr m;
    MPC_TRY(csr_find(handle, &m));
    MPC_ARG_CHECK(planes >= 1 && x && out);
    Scratch sx, so;
    Fr *dx, *dout;
    MPC_TRY(sx.alloc(&dx, m.cols * planes, s));
    MPC_TRY(so.alloc(&dout, m.rows * planes, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dx, x, m.cols * planes * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_TRY(spmv_launch(m, dx, m.cols, planes, dout, m.rows, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(out, dout, m.rows * planes * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

// ---- f3 --------------------------------------------------------------------------------------------------
int32_t mpc_cuda_beaver_mask_serialize_dev(const uint64_t* s_, const uint64_t* x, size_t n, uint8_t* out, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out && (n == 0 || s_) && ((uintptr_t)out & 7) == 0);
    cudaStream_t st = pick_stream(stream, s);
    int g = grid_for(n ? n : 1, LIN_THREADS, 8);
    if (x) k_serialize<true><<<g, LIN_THREADS, 0, st>>>((const Fr*)s_, (const Fr*)x, n, out);
    else k_serialize<false><<<g, LIN_THREADS, 0, st>>>((const Fr*)s_, nullptr, n, out);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

static int32_t serialize_host(const uint64_t* s_, const uint64_t* x, size_t n, uint8_t* out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(out && (n == 0 || s_));
    Scratch sa, sb, so;
    Fr *da, *db = nullptr;
    uint8_t* dout;
    MPC_TRY(sa.alloc(&da, n, s));
    MPC_TRY(so.alloc(&dout, 8 + 32 * n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(da, s_, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    if (x) {
        MPC_TRY(sb.alloc(&db, n, s));
        MPC_CUDA_TRY(cudaMemcpyAsync(db, x, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    }
    MPC_TRY(mpc_cuda_beaver_mask_serialize_dev((const uint64_t*)da, (const uint64_t*)db, n, dout, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(out, dout, 8 + 32 * n, cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_fr_serialize(const uint64_t* vals_mont, size_t n, uint8_t* out) { return serialize_host(vals_mont, nullptr, n, out); }

int32_t mpc_cuda_beaver_mask_serialize(const uint64_t* s_, const uint64_t* x, size_t n, uint8_t* out) {
    MPC_ARG_CHECK(n == 0 || x);
    return serialize_host(s_, x, n, out);
}

int32_t mpc_cuda_open_sum_deserialize_dev(const uint8_t* payloads, uint32_t P, size_t n, uint64_t* out, uint64_t* flags_dev,
                                          void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(P >= 1 && payloads && flags_dev && (n == 0 || out) && ((uintptr_t)payloads & 7) == 0);
    cudaStream_t st = pick_stream(stream, s);
    const unsigned long long init[2] = {~0ull, 0ull};
    MPC_CUDA_TRY(cudaMemcpyAsync(flags_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    size_t work = n > P ? n : P;
    k_deserialize_sum<<<grid_for(work, LIN_THREADS, 8), LIN_THREADS, 0, st>>>(payloads, P, n, (Fr*)out,
                                                                              (unsigned long long*)flags_dev);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_open_sum_deserialize(const uint8_t* payloads, uint32_t P, size_t n, uint64_t* out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(P >= 1 && payloads && (n == 0 || out));
    const size_t bytes = (size_t)P * (8 + 32 * n);
    Scratch si, so, sf;
    uint8_t* din;
    Fr* dout;
    uint64_t* dflags;
    MPC_TRY(si.alloc(&din, bytes, s));
    MPC_TRY(so.alloc(&dout, n, s));
    MPC_TRY(sf.alloc(&dflags, 2, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(din, payloads, bytes, cudaMemcpyHostToDevice, s));
    MPC_TRY(mpc_cuda_open_sum_deserialize_dev(din, P, n, (uint64_t*)dout, dflags, s));
    uint64_t flags[2];
    MPC_CUDA_TRY(cudaMemcpyAsync(flags, dflags, sizeof(flags), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(out, dout, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    if (flags[1]) {
        set_error("deserialize: a payload's length prefix differs from %zu", n);
        return MPC_CUDA_ERR_ARG;
    }
    if (flags[0] != ~0ull) {
        set_error("deserialize: element %llu is not below the Fr modulus (SerializationError::InvalidData)",
                  (unsigned long long)(flags[0] - 1));
        return MPC_CUDA_ERR_ARG;
    }
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_fr_deserialize(const uint8_t* in, size_t n, uint64_t* vals_mont) {
    return mpc_cuda_open_sum_deserialize(in, 1, n, vals_mont);
}

// ---- f4 --------------------------------------------------------------------------------------------------
int32_t mpc_cuda_poly_div_linear_dev(const uint64_t* coeffs, size_t n, const uint64_t* z_mont_host, uint64_t* q_out,
                                     uint64_t* rem_out, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    return poly_div_linear_dev((const Fr*)coeffs, n, z_mont_host, (Fr*)q_out, (Fr*)rem_out, pick_stream(stream, s));
}

int32_t mpc_cuda_poly_div_linear(const uint64_t* coeffs, size_t n, const uint64_t* z_mont, uint64_t* q_out, uint64_t* rem_out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(coeffs && z_mont && rem_out && n >= 1);
    Scratch sp, sq, sr;
    Fr *dp, *dq = nullptr, *dr;
    MPC_TRY(sp.alloc(&dp, n, s));
    MPC_TRY(sr.alloc(&dr, 1, s));
    if (q_out && n > 1) MPC_TRY(sq.alloc(&dq, n - 1, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dp, coeffs, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_TRY(poly_div_linear_dev(dp, n, z_mont, dq, dr, s));
    if (dq) MPC_CUDA_TRY(cudaMemcpyAsync(q_out, dq, (n - 1) * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(rem_out, dr, sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_poly_div_vanishing_dev(const uint64_t* coeffs, size_t n, size_t m, uint64_t* q_out, uint64_t* rem_out,
                                        void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    return poly_div_vanishing_dev((const Fr*)coeffs, n, m, (Fr*)q_out, (Fr*)rem_out, pick_stream(stream, s));
}

int32_t mpc_cuda_poly_div_vanishing(const uint64_t* coeffs, size_t n, size_t m, uint64_t* q_out, uint64_t* rem_out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(coeffs && rem_out && n >= 1 && m >= 1);
    Scratch sp, sq, sr;
    Fr *dp, *dq = nullptr, *dr;
    MPC_TRY(sp.alloc(&dp, n, s));
    MPC_TRY(sr.alloc(&dr, m, s));
    if (q_out && n > m) MPC_TRY(sq.alloc(&dq, n - m, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dp, coeffs, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_TRY(poly_div_vanishing_dev(dp, n, m, dq, dr, s));
    if (dq) MPC_CUDA_TRY(cudaMemcpyAsync(q_out, dq, (n - m) * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(rem_out, dr, m * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_poly_mul_vanishing_dev(const uint64_t* coeffs, size_t n, size_t m, uint64_t* out, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    return poly_mul_vanishing_dev((const Fr*)coeffs, n, m, (Fr*)out, pick_stream(stream, s));
}

int32_t mpc_cuda_poly_mul_vanishing(const uint64_t* coeffs, size_t n, size_t m, uint64_t* out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(coeffs && out && n >= 1 && m >= 1);
    Scratch sp, so;
    Fr *dp, *dout;
    MPC_TRY(sp.alloc(&dp, n, s));
    MPC_TRY(so.alloc(&dout, n + m, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(dp, coeffs, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_TRY(poly_mul_vanishing_dev(dp, n, m, dout, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(out, dout, (n + m) * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    return MPC_CUDA_OK;
}

}  // extern "C"
