// Elementwise share kernels over Fr: the local halves of Beaver multiplication and batch_open,
// plus the vector helpers the provers use around the NTTs.  All of them are one pass over HBM
// with 128-bit loads/stores, grid = SM count x resident CTAs, grid-stride loop.
//
// Reference semantics (mpc-algebra/src/share):
//   mask      s + x                                   field.rs:108-117, additive.rs:132-135
//   combine   z - y*sx - x*oy (+ sx*oy on the leader) field.rs:118-128, additive.rs:136-152, spdz.rs:202-219
//   open_sum  sum over parties                        additive.rs:125-131, spdz.rs:181-184
//   mac_check mac_share*val - mac                     spdz.rs:185-189
#include "common.cuh"

using namespace mpc;

namespace {

constexpr int EW_THREADS = 256;

__global__ void __launch_bounds__(EW_THREADS) k_mask(const Fr* __restrict__ s, const Fr* __restrict__ x,
                                                     Fr* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr a = load_fe_ro(s + i), b = load_fe_ro(x + i);
        store_fe(out + i, add(a, b));
    }
}

// Algorithmic traffic: additive 6 x 32 B = 192 B/element; SPDZ 10 x 32 B = 320 B per share pair.
template <bool LEADER, bool SPDZ>
__global__ void __launch_bounds__(EW_THREADS) k_combine(const Fr* __restrict__ x, const Fr* __restrict__ y,
                                                        const Fr* __restrict__ z, const Fr* __restrict__ sx,
                                                        const Fr* __restrict__ oy, Fr* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // issue every load before the arithmetic so the 5 (or 8) requests are in flight together
        Fr vx = load_fe_ro(x + i), vy = load_fe_ro(y + i), vz = load_fe_ro(z + i);
        Fr vsx = load_fe_ro(sx + i), voy = load_fe_ro(oy + i);
        Fr mx, my, mz;
        if (SPDZ) {
            mx = load_fe_ro(x + n + i);
            my = load_fe_ro(y + n + i);
            mz = load_fe_ro(z + n + i);
        }
        Fr shift;
        if (LEADER) shift = mul(vsx, voy);
        Fr r = sub(sub(vz, mul(vy, vsx)), mul(vx, voy));
        if (LEADER) r = add(r, shift);
        store_fe(out + i, r);
        if (SPDZ) {
            Fr m = sub(sub(mz, mul(my, vsx)), mul(mx, voy));
            if (LEADER) m = add(m, shift);      // mac_share = 1 on the leader
            store_fe(out + n + i, m);
        }
    }
}

__global__ void __launch_bounds__(EW_THREADS) k_open_sum(const Fr* __restrict__ parts, uint32_t P,
                                                         Fr* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr acc = load_fe_ro(parts + i);
        for (uint32_t p = 1; p < P; p++) acc = add(acc, load_fe_ro(parts + (size_t)p * n + i));
        store_fe(out + i, acc);
    }
}

template <bool LEADER>
__global__ void __launch_bounds__(EW_THREADS) k_mac_check(const Fr* __restrict__ vals, const Fr* __restrict__ macs,
                                                          Fr* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr m = load_fe_ro(macs + i);
        Fr v = LEADER ? load_fe_ro(vals + i) : Fr::zero();
        store_fe(out + i, sub(v, m));
    }
}

// `out` may be `a` (in place): neither is declared restrict and `a` is read through the coherent path
template <int OP>
__global__ void __launch_bounds__(EW_THREADS) k_vec_op(const Fr* a, const Fr* __restrict__ b, Fr c, Fr* out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr va = load_fe(a + i);
        Fr r;
        if (OP == MPC_CUDA_VEC_SUB) r = sub(va, load_fe_ro(b + i));
        else if (OP == MPC_CUDA_VEC_MUL) r = mul(va, load_fe_ro(b + i));
        else if (OP == MPC_CUDA_VEC_MUL_CONST) r = mul(va, c);
        else r = add(va, mul(c, load_fe_ro(b + i)));
        store_fe(out + i, r);
    }
}

int ew_grid(size_t n) { return grid_for(n, EW_THREADS, 8); }

// stage host operands on the device, run `launch`, copy `out_elems` results back, synchronise
struct HostStage {
    cudaStream_t s;
    Scratch bufs[8];
    int used = 0;
    int32_t in(const uint64_t* h, size_t elems, const Fr** d) {
        Fr* p;
        MPC_TRY(bufs[used++].alloc(&p, elems, s));
        MPC_CUDA_TRY(cudaMemcpyAsync(p, h, elems * sizeof(Fr), cudaMemcpyHostToDevice, s));
        *d = p;
        return MPC_CUDA_OK;
    }
    int32_t out(size_t elems, Fr** d) { return bufs[used++].alloc(d, elems, s); }
    int32_t finish(uint64_t* h, const Fr* d, size_t elems) {
        MPC_CUDA_TRY(cudaMemcpyAsync(h, d, elems * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        MPC_CUDA_TRY(cudaStreamSynchronize(s));
        return MPC_CUDA_OK;
    }
};

// elementwise inverse of public values (the denominators of Marlin's r(alpha, .), ahp/mod.rs:357-364); zero stays
// zero as in ark_ff::batch_inversion (ff/src/fields/mod.rs:597-660), whose trick this is: every thread takes INV_BATCH
// strided elements, multiplies the non-zero ones up, inverts the product once (binary Euclid) and peels the
// individual inverses off backwards: 3 products per element instead of an inversion.
constexpr int INV_BATCH = 8;
__global__ void __launch_bounds__(128) k_inverse(const Fr* __restrict__ a, Fr* __restrict__ out, size_t n) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (tid >= n) return;
    Fr x[INV_BATCH], pre[INV_BATCH];
    Fr run = Fr::one();
#pragma unroll
    for (int k = 0; k < INV_BATCH; k++) {
        const size_t i = tid + (size_t)k * stride;
        x[k] = i < n ? load_fe(a + i) : Fr::zero();
        pre[k] = run;
        if (!x[k].is_zero()) run = mul(run, x[k]);
    }
    Fr inv_run = inv_euclid(run);                  // run is a product of non-zero elements (1 if there were none)
#pragma unroll
    for (int k = INV_BATCH - 1; k >= 0; k--) {
        const size_t i = tid + (size_t)k * stride;
        if (i >= n) continue;
        if (x[k].is_zero()) {
            store_fe(out + i, x[k]);
        } else {
            store_fe(out + i, mul(inv_run, pre[k]));
            inv_run = mul(inv_run, x[k]);
        }
    }
}

}  // namespace

extern "C" {

int32_t mpc_cuda_beaver_mask_dev(const uint64_t* s_, const uint64_t* x, uint64_t* out, size_t n, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(s_ && x && out);
    k_mask<<<ew_grid(n), EW_THREADS, 0, pick_stream(stream, s)>>>((const Fr*)s_, (const Fr*)x, (Fr*)out, n);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_beaver_mask(const uint64_t* s_, const uint64_t* x, uint64_t* out, size_t n) {
    HostStage st;
    MPC_TRY(enter(&st.s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(s_ && x && out);
    const Fr *ds, *dx;
    Fr* dout;
    MPC_TRY(st.in(s_, n, &ds));
    MPC_TRY(st.in(x, n, &dx));
    MPC_TRY(st.out(n, &dout));
    MPC_TRY(mpc_cuda_beaver_mask_dev((const uint64_t*)ds, (const uint64_t*)dx, (uint64_t*)dout, n, st.s));
    return st.finish(out, dout, n);
}

int32_t mpc_cuda_beaver_combine_dev(const uint64_t* x, const uint64_t* y, const uint64_t* z, const uint64_t* sx,
                                    const uint64_t* oy, uint64_t* out, size_t n, uint32_t is_leader, uint32_t spdz,
                                    void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(x && y && z && sx && oy && out);
    cudaStream_t st = pick_stream(stream, s);
    int g = ew_grid(n);
    const Fr *fx = (const Fr*)x, *fy = (const Fr*)y, *fz = (const Fr*)z, *fsx = (const Fr*)sx, *foy = (const Fr*)oy;
    Fr* fo = (Fr*)out;
    if (is_leader) {
        if (spdz) k_combine<true, true><<<g, EW_THREADS, 0, st>>>(fx, fy, fz, fsx, foy, fo, n);
        else k_combine<true, false><<<g, EW_THREADS, 0, st>>>(fx, fy, fz, fsx, foy, fo, n);
    } else {
        if (spdz) k_combine<false, true><<<g, EW_THREADS, 0, st>>>(fx, fy, fz, fsx, foy, fo, n);
        else k_combine<false, false><<<g, EW_THREADS, 0, st>>>(fx, fy, fz, fsx, foy, fo, n);
    }
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_beaver_combine(const uint64_t* x, const uint64_t* y, const uint64_t* z, const uint64_t* sx,
                                const uint64_t* oy, uint64_t* out, size_t n, uint32_t is_leader, uint32_t spdz) {
    HostStage st;
    MPC_TRY(enter(&st.s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(x && y && z && sx && oy && out);
    size_t m = spdz ? 2 * n : n;
    const Fr *dx, *dy, *dz, *dsx, *doy;
    Fr* dout;
    MPC_TRY(st.in(x, m, &dx));
    MPC_TRY(st.in(y, m, &dy));
    MPC_TRY(st.in(z, m, &dz));
    MPC_TRY(st.in(sx, n, &dsx));
    MPC_TRY(st.in(oy, n, &doy));
    MPC_TRY(st.out(m, &dout));
    MPC_TRY(mpc_cuda_beaver_combine_dev((const uint64_t*)dx, (const uint64_t*)dy, (const uint64_t*)dz,
                                        (const uint64_t*)dsx, (const uint64_t*)doy, (uint64_t*)dout, n, is_leader,
                                        spdz, st.s));
    return st.finish(out, dout, m);
}

int32_t mpc_cuda_open_sum_dev(const uint64_t* parts, uint32_t P, uint64_t* out, size_t n, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(P >= 1);
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(parts && out);
    k_open_sum<<<ew_grid(n), EW_THREADS, 0, pick_stream(stream, s)>>>((const Fr*)parts, P, (Fr*)out, n);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_open_sum(const uint64_t* parts, uint32_t P, uint64_t* out, size_t n) {
    HostStage st;
    MPC_TRY(enter(&st.s));
    MPC_ARG_CHECK(P >= 1);
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(parts && out);
    const Fr* dp;
    Fr* dout;
    MPC_TRY(st.in(parts, (size_t)P * n, &dp));
    MPC_TRY(st.out(n, &dout));
    MPC_TRY(mpc_cuda_open_sum_dev((const uint64_t*)dp, P, (uint64_t*)dout, n, st.s));
    return st.finish(out, dout, n);
}

int32_t mpc_cuda_spdz_mac_check_dev(const uint64_t* vals, const uint64_t* macs, uint64_t* out, size_t n,
                                    uint32_t is_leader, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(vals && macs && out);
    cudaStream_t st = pick_stream(stream, s);
    if (is_leader) k_mac_check<true><<<ew_grid(n), EW_THREADS, 0, st>>>((const Fr*)vals, (const Fr*)macs, (Fr*)out, n);
    else k_mac_check<false><<<ew_grid(n), EW_THREADS, 0, st>>>((const Fr*)vals, (const Fr*)macs, (Fr*)out, n);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_spdz_mac_check(const uint64_t* vals, const uint64_t* macs, uint64_t* out, size_t n,
                                uint32_t is_leader) {
    HostStage st;
    MPC_TRY(enter(&st.s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(vals && macs && out);
    const Fr *dv, *dm;
    Fr* dout;
    MPC_TRY(st.in(vals, n, &dv));
    MPC_TRY(st.in(macs, n, &dm));
    MPC_TRY(st.out(n, &dout));
    MPC_TRY(mpc_cuda_spdz_mac_check_dev((const uint64_t*)dv, (const uint64_t*)dm, (uint64_t*)dout, n, is_leader, st.s));
    return st.finish(out, dout, n);
}

int32_t mpc_cuda_fr_inverse_dev(const uint64_t* a, uint64_t* out, size_t n, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(a && out);
    const size_t threads = (n + INV_BATCH - 1) / INV_BATCH;
    k_inverse<<<(unsigned)((threads + 127) / 128), 128, 0, pick_stream(stream, s)>>>((const Fr*)a, (Fr*)out, n);
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_vec_op_dev(uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c_host, uint64_t* out,
                            size_t n, void* stream) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(op <= MPC_CUDA_VEC_AXPY);
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(a && out);
    MPC_ARG_CHECK(op == MPC_CUDA_VEC_MUL_CONST || b);
    MPC_ARG_CHECK(op < MPC_CUDA_VEC_MUL_CONST || c_host);
    Fr c = Fr::zero();
    if (c_host) memcpy(c.v, c_host, sizeof(Fr));
    cudaStream_t st = pick_stream(stream, s);
    int g = ew_grid(n);
    const Fr *fa = (const Fr*)a, *fb = (const Fr*)b;
    Fr* fo = (Fr*)out;
    switch (op) {
        case MPC_CUDA_VEC_SUB: k_vec_op<MPC_CUDA_VEC_SUB><<<g, EW_THREADS, 0, st>>>(fa, fb, c, fo, n); break;
        case MPC_CUDA_VEC_MUL: k_vec_op<MPC_CUDA_VEC_MUL><<<g, EW_THREADS, 0, st>>>(fa, fb, c, fo, n); break;
        case MPC_CUDA_VEC_MUL_CONST: k_vec_op<MPC_CUDA_VEC_MUL_CONST><<<g, EW_THREADS, 0, st>>>(fa, fb, c, fo, n); break;
        default: k_vec_op<MPC_CUDA_VEC_AXPY><<<g, EW_THREADS, 0, st>>>(fa, fb, c, fo, n); break;
    }
    MPC_KERNEL_CHECK();
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_vec_op(uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* out, size_t n) {
    HostStage st;
    MPC_TRY(enter(&st.s));
    MPC_ARG_CHECK(op <= MPC_CUDA_VEC_AXPY);
    if (n == 0) return MPC_CUDA_OK;
    MPC_ARG_CHECK(a && out);
    const Fr *da, *db = nullptr;
    Fr* dout;
    MPC_TRY(st.in(a, n, &da));
    if (op != MPC_CUDA_VEC_MUL_CONST) {
        MPC_ARG_CHECK(b);
        MPC_TRY(st.in(b, n, &db));
    }
    MPC_TRY(st.out(n, &dout));
    MPC_TRY(mpc_cuda_vec_op_dev(op, (const uint64_t*)da, (const uint64_t*)db, c, (uint64_t*)dout, n, st.s));
    return st.finish(out, dout, n);
}

}  // extern "C"
