// Fused R1CStoQAP::witness_map on one party's local share values (src/groth16.rs:240-307): the vectors stay
// resident in HBM across the constraint-row evaluation (public CSR matrices x share vector, :263-276,289-293),
// 3 iFFT, 3 coset FFT, the Beaver mask, and — after the two network opens that stay on the host's mpc-net — the
// Beaver combine, the subtraction, the division by the vanishing polynomial and the coset iFFT; h can stay on
// the device as the scalar vector of the h_query MSM (src/groth16.rs:106).  Only the masked values leave the
// device in between.  Additive shares are one plane of n values; SPDZ shares are two planes
// [sh | mac] (mpc-algebra/src/share/spdz.rs:50-53): every linear step runs on both planes, the Beaver combine
// follows spdz.rs:197-219.  Everything is composed from the library's own stream-ordered entry points.
#include <mutex>
#include <unordered_map>

#include "common.cuh"

using namespace mpc;

extern "C" int32_t mpc_cuda_csr_spmv_dev(uint64_t handle, const uint64_t* x_dev, size_t x_stride, uint32_t planes,
                                         uint64_t* out_dev, size_t out_stride, void* stream);
extern "C" int32_t mpc_cuda_csr_dims(uint64_t handle, size_t* rows, size_t* cols, size_t* nnz);

namespace {

struct WitnessState {
    Fr* buf = nullptr;          // [a | b | c | tx | ty | ma | mb], each `planes` x n; later reused for the combine
    size_t n = 0;
    uint32_t log_n = 0, planes = 1;
    int cuda_device = 0;
    bool finished = false;      // finish_dev ran: buf + 6*planes*n holds h
};
std::mutex g_ws_mu;
std::unordered_map<uint64_t, WitnessState> g_ws;
uint64_t g_ws_next = 1;

struct Layout {
    Fr *a, *b, *c, *tx, *ty, *ma, *mb;
    Layout(const WitnessState& w) {
        size_t pn = (size_t)w.planes * w.n;
        a = w.buf; b = a + pn; c = b + pn; tx = c + pn; ty = tx + pn; ma = ty + pn; mb = ma + pn;
    }
};

int32_t alloc_state(WitnessState* w, uint32_t log_n, uint32_t spdz) {
    MPC_ARG_CHECK(log_n <= 28);
    w->n = (size_t)1 << log_n;
    w->log_n = log_n;
    w->planes = spdz ? 2 : 1;
    w->cuda_device = current_device_info()->cuda_device;
    MPC_CUDA_TRY(cudaMalloc((void**)&w->buf, 7 * (size_t)w->planes * w->n * sizeof(Fr)));
    return MPC_CUDA_OK;
}

// a, b, c (evaluations over the domain) are in place on the device: transforms, masks, copy out, publish
int32_t begin_tail(WitnessState& w, const uint64_t* tx, const uint64_t* ty, uint64_t* masked_a, uint64_t* masked_b,
                   uint64_t* state, cudaStream_t s) {
    Layout L(w);
    const size_t pn = (size_t)w.planes * w.n, bytes = pn * sizeof(Fr);
    MPC_CUDA_TRY(cudaMemcpyAsync(L.tx, tx, bytes, cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(L.ty, ty, bytes, cudaMemcpyHostToDevice, s));
    // ifft then coset_fft of a, b, c (every plane): 3 x planes equal-size transforms per call
    MPC_TRY(mpc_cuda_ntt_fr_dev((uint64_t*)L.a, w.log_n, MPC_CUDA_NTT_IFFT, 3 * w.planes, s));
    MPC_TRY(mpc_cuda_ntt_fr_dev((uint64_t*)L.a, w.log_n, MPC_CUDA_NTT_COSET_FFT, 3 * w.planes, s));
    // Beaver masks of the batch product a' * b' (share/field.rs:108-117): one launch over [a | b] + [tx | ty]
    MPC_TRY(mpc_cuda_beaver_mask_dev((const uint64_t*)L.a, (const uint64_t*)L.tx, (uint64_t*)L.ma, 2 * pn, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(masked_a, L.ma, bytes, cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(masked_b, L.mb, bytes, cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> lk(g_ws_mu);
    *state = g_ws_next++;
    g_ws[*state] = w;
    return MPC_CUDA_OK;
}

int32_t take_state(uint64_t state, bool erase, WitnessState* w) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto it = g_ws.find(state);
    if (it == g_ws.end()) {
        set_error("unknown witness_map state %llu", (unsigned long long)state);
        return MPC_CUDA_ERR_HANDLE;
    }
    *w = it->second;
    if (erase) g_ws.erase(it);
    return MPC_CUDA_OK;
}

// second half on the device; h = planes x n values at buf + 6*planes*n
int32_t finish_core(const WitnessState& w, const uint64_t* tz, const uint64_t* sx, const uint64_t* oy, uint32_t is_leader,
                    cudaStream_t s, Fr** h) {
    MPC_ARG_CHECK(tz && sx && oy);
    if (w.cuda_device != current_device_info()->cuda_device) {
        set_error("witness_map state lives on another device");
        return MPC_CUDA_ERR_HANDLE;
    }
    Layout L(w);
    const size_t n = w.n, pn = (size_t)w.planes * n;
    // a', b' are no longer needed: their planes receive the triple's z shares and the opened values
    Fr *vz = L.a, *vsx = L.b, *voy = L.b + n, *vs = L.mb;
    Scratch tmp;                                  // additive layout: b has one plane, oy needs its own buffer
    if (w.planes == 1) MPC_TRY(tmp.alloc(&voy, n, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(vz, tz, pn * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(vsx, sx, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(voy, oy, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    // ab = z - y*sx - x*oy (+ sx*oy on the leader; mac plane + mac_share*sx*oy)   share/field.rs:118-128
    MPC_TRY(mpc_cuda_beaver_combine_dev((const uint64_t*)L.tx, (const uint64_t*)L.ty, (const uint64_t*)vz,
                                        (const uint64_t*)vsx, (const uint64_t*)voy, (uint64_t*)vs, n, is_leader,
                                        w.planes == 2, s));
    // ab -= c' ; ab /= Z_H on the coset ; h = coset_ifft(ab)   src/groth16.rs:298-303 (out of place: ma is free)
    MPC_TRY(mpc_cuda_vec_op_dev(MPC_CUDA_VEC_SUB, (const uint64_t*)vs, (const uint64_t*)L.c, nullptr, (uint64_t*)L.ma, pn, s));
    for (uint32_t p = 0; p < w.planes; p++)
        MPC_TRY(mpc_cuda_divide_by_vanishing_on_coset_dev((uint64_t*)(L.ma + p * n), w.log_n, s));
    MPC_TRY(mpc_cuda_ntt_fr_dev((uint64_t*)L.ma, w.log_n, MPC_CUDA_NTT_COSET_IFFT, w.planes, s));
    *h = L.ma;
    return MPC_CUDA_OK;
}

}  // namespace

extern "C" {

int32_t mpc_cuda_witness_map_begin_ex(const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                                      const uint64_t* tx, const uint64_t* ty, uint32_t spdz, uint64_t* masked_a,
                                      uint64_t* masked_b, uint64_t* state) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(a && b && c && tx && ty && masked_a && masked_b && state);
    WitnessState w;
    MPC_TRY(alloc_state(&w, log_n, spdz));
    Layout L(w);
    const size_t bytes = (size_t)w.planes * w.n * sizeof(Fr);
    int32_t rc = MPC_CUDA_OK;
    if (cudaMemcpyAsync(L.a, a, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(L.b, b, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(L.c, c, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) {
        set_error("witness_map_begin: host to device copy failed");
        rc = MPC_CUDA_ERR_CUDA;
    }
    if (rc == MPC_CUDA_OK) rc = begin_tail(w, tx, ty, masked_a, masked_b, state, s);
    if (rc != MPC_CUDA_OK) cudaFree(w.buf);
    return rc;
}

int32_t mpc_cuda_witness_map_begin(const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                                   const uint64_t* tx, const uint64_t* ty, uint64_t* masked_a, uint64_t* masked_b,
                                   uint64_t* state) {
    return mpc_cuda_witness_map_begin_ex(a, b, c, log_n, tx, ty, 0, masked_a, masked_b, state);
}

int32_t mpc_cuda_witness_map_begin_r1cs(uint64_t csr_a, uint64_t csr_b, uint64_t csr_c, const uint64_t* assignment,
                                        size_t num_inputs, uint32_t log_n, const uint64_t* tx, const uint64_t* ty,
                                        uint32_t spdz, uint64_t* masked_a, uint64_t* masked_b, uint64_t* state) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(assignment && tx && ty && masked_a && masked_b && state);
    size_t rows = 0, cols = 0, rb = 0, cb = 0, rc_ = 0, cc = 0;
    MPC_TRY(mpc_cuda_csr_dims(csr_a, &rows, &cols, nullptr));
    MPC_TRY(mpc_cuda_csr_dims(csr_b, &rb, &cb, nullptr));
    MPC_TRY(mpc_cuda_csr_dims(csr_c, &rc_, &cc, nullptr));
    // domain = D::new(num_constraints + num_inputs) (src/groth16.rs:258-260); the caller passes its log size
    MPC_ARG_CHECK(rb == rows && rc_ == rows && cb == cols && cc == cols && num_inputs <= cols);
    MPC_ARG_CHECK(log_n <= 28 && rows + num_inputs <= ((size_t)1 << log_n));
    WitnessState w;
    MPC_TRY(alloc_state(&w, log_n, spdz));
    Layout L(w);
    const size_t n = w.n, pn = (size_t)w.planes * n;
    Scratch sz;
    Fr* z;                                         // the assignment, planes x cols
    int32_t rc = sz.alloc(&z, (size_t)w.planes * cols, s);
    auto step = [&](int32_t r) { if (rc == MPC_CUDA_OK) rc = r; };
    auto cu = [&](cudaError_t e) {
        if (rc == MPC_CUDA_OK && e != cudaSuccess) {
            set_error("witness_map_begin_r1cs: %s", cudaGetErrorString(e));
            rc = MPC_CUDA_ERR_CUDA;
        }
    };
    if (rc == MPC_CUDA_OK) {
        cu(cudaMemcpyAsync(z, assignment, (size_t)w.planes * cols * sizeof(Fr), cudaMemcpyHostToDevice, s));
        cu(cudaMemsetAsync(L.a, 0, 3 * pn * sizeof(Fr), s));       // rows beyond the constraints stay zero
        step(mpc_cuda_csr_spmv_dev(csr_a, (const uint64_t*)z, cols, w.planes, (uint64_t*)L.a, n, s));
        step(mpc_cuda_csr_spmv_dev(csr_b, (const uint64_t*)z, cols, w.planes, (uint64_t*)L.b, n, s));
        step(mpc_cuda_csr_spmv_dev(csr_c, (const uint64_t*)z, cols, w.planes, (uint64_t*)L.c, n, s));
        // a[num_constraints .. + num_inputs) = the instance assignment (src/groth16.rs:272-276)
        for (uint32_t p = 0; p < w.planes; p++)
            cu(cudaMemcpyAsync(L.a + p * n + rows, z + p * cols, num_inputs * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        step(rc == MPC_CUDA_OK ? begin_tail(w, tx, ty, masked_a, masked_b, state, s) : rc);
    }
    if (rc != MPC_CUDA_OK) cudaFree(w.buf);
    return rc;
}

int32_t mpc_cuda_witness_map_finish(uint64_t state, const uint64_t* tz, const uint64_t* sx, const uint64_t* oy,
                                    uint32_t is_leader, uint64_t* h_out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    WitnessState w;
    MPC_TRY(take_state(state, true, &w));
    Fr* h = nullptr;
    int32_t rc = h_out ? finish_core(w, tz, sx, oy, is_leader, s, &h) : MPC_CUDA_ERR_ARG;
    if (!h_out) set_error("witness_map_finish: null argument");
    if (rc == MPC_CUDA_OK &&
        (cudaMemcpyAsync(h_out, h, (size_t)w.planes * w.n * sizeof(Fr), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
         cudaStreamSynchronize(s) != cudaSuccess)) {
        set_error("witness_map_finish: device to host copy failed");
        rc = MPC_CUDA_ERR_CUDA;
    }
    cudaFree(w.buf);            // finish always releases the state
    return rc;
}

int32_t mpc_cuda_witness_map_finish_dev(uint64_t state, const uint64_t* tz, const uint64_t* sx, const uint64_t* oy,
                                        uint32_t is_leader, uint64_t** h_dev) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(h_dev != nullptr);
    WitnessState w;
    MPC_TRY(take_state(state, false, &w));
    if (w.finished) {
        set_error("witness_map state %llu was already finished", (unsigned long long)state);
        return MPC_CUDA_ERR_HANDLE;
    }
    Fr* h = nullptr;
    MPC_TRY(finish_core(w, tz, sx, oy, is_leader, s, &h));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));       // other streams (the h_query MSM) may read h from here on
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        auto it = g_ws.find(state);
        if (it != g_ws.end()) it->second.finished = true;
    }
    *h_dev = (uint64_t*)h;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_witness_map_release(uint64_t state) {
    MPC_TRY(enter(nullptr));
    WitnessState w;
    MPC_TRY(take_state(state, true, &w));
    int cur = 0;
    MPC_CUDA_TRY(cudaGetDevice(&cur));
    MPC_CUDA_TRY(cudaSetDevice(w.cuda_device));
    cudaFree(w.buf);
    MPC_CUDA_TRY(cudaSetDevice(cur));
    return MPC_CUDA_OK;
}

}  // extern "C"
