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
    Fr* buf = nullptr;          // [a | b | c | tx | ty | ma | mb], each `planes` x n; later reused for the combine;
                                // then [opened sx | opened oy] (n each) and the assignment (planes x cols, _begin_r1cs)
    size_t n = 0, cols = 0;
    bool opened[2] = {false, false};   // the open of masked_a / masked_b arrived through _open_payloads
    uint32_t log_n = 0, planes = 1;
    int cuda_device = 0;
    bool finished = false;      // finish_dev ran: buf + 6*planes*n holds h
};
std::mutex g_ws_mu;
std::unordered_map<uint64_t, WitnessState> g_ws;
uint64_t g_ws_next = 1;

struct Layout {
    Fr *a, *b, *c, *tx, *ty, *ma, *mb, *open, *z;
    Layout(const WitnessState& w) {
        size_t pn = (size_t)w.planes * w.n;
        a = w.buf; b = a + pn; c = b + pn; tx = c + pn; ty = tx + pn; ma = ty + pn; mb = ma + pn;
        open = mb + pn; z = open + 2 * w.n;
    }
};

// The state's buffer comes from the device's stream-ordered pool (never trimmed): a cudaMalloc / cudaFree pair per
// proof synchronises the device and stalls the other parties' streams.
int32_t alloc_state(WitnessState* w, uint32_t log_n, uint32_t spdz, size_t cols, cudaStream_t s) {
    MPC_ARG_CHECK(log_n <= 28);
    w->n = (size_t)1 << log_n;
    w->log_n = log_n;
    w->planes = spdz ? 2 : 1;
    w->cols = cols;
    w->cuda_device = current_device_info()->cuda_device;
    MPC_CUDA_TRY(cudaMallocAsync((void**)&w->buf, ((7 * (size_t)w->planes + 2) * w->n + (size_t)w->planes * cols) * sizeof(Fr), s));
    return MPC_CUDA_OK;
}

int32_t check_device(const WitnessState& w) {
    if (w.cuda_device != current_device_info()->cuda_device) {
        set_error("witness_map state lives on another device");
        return MPC_CUDA_ERR_HANDLE;
    }
    return MPC_CUDA_OK;
}

// flag[0] |= 1 when any element differs from zero (the zero test of the SPDZ MAC check's opened sum)
__global__ void k_any_nonzero(const Fr* __restrict__ v, size_t n, unsigned int* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = load_fe_ro(v + i);
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < Fr::N; k++) acc |= x.v[k];
    if (acc) atomicOr(flag, 1u);
}

// the received payloads (host pointers, 8 + 32 n bytes each) summed on the device into `out`
int32_t sum_payloads(const uint8_t* const* payloads, uint32_t P, size_t n, Fr* out, cudaStream_t s) {
    MPC_ARG_CHECK(payloads && P >= 1 && P <= 4096);
    for (uint32_t p = 0; p < P; p++) MPC_ARG_CHECK(payloads[p] != nullptr);
    const size_t len = 8 + 32 * n;
    Scratch si, sf;
    uint8_t* din;
    uint64_t* dflags;
    MPC_TRY(si.alloc(&din, (size_t)P * len, s));
    MPC_TRY(sf.alloc(&dflags, 2, s));
    for (uint32_t p = 0; p < P; p++)
        MPC_CUDA_TRY(cudaMemcpyAsync(din + p * len, payloads[p], len, cudaMemcpyHostToDevice, s));
    MPC_TRY(mpc_cuda_open_sum_deserialize_dev(din, P, n, (uint64_t*)out, dflags, s));
    uint64_t flags[2];
    MPC_CUDA_TRY(cudaMemcpyAsync(flags, dflags, sizeof(flags), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));       // the caller may reuse the payload buffers from here on
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

// n elements on the device -> wire payload in host memory
int32_t payload_out(const Fr* v, size_t n, uint8_t* out, cudaStream_t s) {
    Scratch so;
    uint8_t* dout;
    MPC_TRY(so.alloc(&dout, 8 + 32 * n, s));
    MPC_TRY(mpc_cuda_beaver_mask_serialize_dev((const uint64_t*)v, nullptr, n, dout, s));
    MPC_CUDA_TRY(cudaMemcpyAsync(out, dout, 8 + 32 * n, cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
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
    if (masked_a) {                               // NULL: the opens go through _masked_payload / _open_payloads
        MPC_CUDA_TRY(cudaMemcpyAsync(masked_a, L.ma, bytes, cudaMemcpyDeviceToHost, s));
        MPC_CUDA_TRY(cudaMemcpyAsync(masked_b, L.mb, bytes, cudaMemcpyDeviceToHost, s));
    }
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
    MPC_ARG_CHECK(tz && (sx || w.opened[0]) && (oy || w.opened[1]));
    MPC_TRY(check_device(w));
    Layout L(w);
    const size_t n = w.n, pn = (size_t)w.planes * n;
    // a', b' are no longer needed: their planes receive the triple's z shares and the opened values
    Fr *vz = L.a, *vsx = L.open, *voy = L.open + n, *vs = L.mb;
    MPC_CUDA_TRY(cudaMemcpyAsync(vz, tz, pn * sizeof(Fr), cudaMemcpyHostToDevice, s));
    if (sx) MPC_CUDA_TRY(cudaMemcpyAsync(vsx, sx, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    if (oy) MPC_CUDA_TRY(cudaMemcpyAsync(voy, oy, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
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
    MPC_ARG_CHECK(a && b && c && tx && ty && state && !masked_a == !masked_b);
    WitnessState w;
    MPC_TRY(alloc_state(&w, log_n, spdz, 0, s));
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
    if (rc != MPC_CUDA_OK) cudaFreeAsync(w.buf, s);
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
    MPC_ARG_CHECK(assignment && tx && ty && state && !masked_a == !masked_b);
    size_t rows = 0, cols = 0, rb = 0, cb = 0, rc_ = 0, cc = 0;
    MPC_TRY(mpc_cuda_csr_dims(csr_a, &rows, &cols, nullptr));
    MPC_TRY(mpc_cuda_csr_dims(csr_b, &rb, &cb, nullptr));
    MPC_TRY(mpc_cuda_csr_dims(csr_c, &rc_, &cc, nullptr));
    // domain = D::new(num_constraints + num_inputs) (src/groth16.rs:258-260); the caller passes its log size
    MPC_ARG_CHECK(rb == rows && rc_ == rows && cb == cols && cc == cols && num_inputs <= cols);
    MPC_ARG_CHECK(log_n <= 28 && rows + num_inputs <= ((size_t)1 << log_n));
    WitnessState w;
    MPC_TRY(alloc_state(&w, log_n, spdz, cols, s));
    Layout L(w);
    const size_t n = w.n, pn = (size_t)w.planes * n;
    Fr* z = L.z;                                   // the assignment, planes x cols, kept for _assignment_dev
    int32_t rc = MPC_CUDA_OK;
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
    if (rc != MPC_CUDA_OK) cudaFreeAsync(w.buf, s);
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
    cudaFreeAsync(w.buf, s);    // finish always releases the state (the copy out has completed)
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

int32_t mpc_cuda_witness_map_masked_payload(uint64_t state, uint32_t which, uint8_t* out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(which <= 1 && out);
    WitnessState w;
    MPC_TRY(take_state(state, false, &w));
    MPC_TRY(check_device(w));
    Layout L(w);
    return payload_out(which ? L.mb : L.ma, w.n, out, s);          // SPDZ: the sh plane comes first
}

int32_t mpc_cuda_witness_map_open_payloads(uint64_t state, uint32_t which, const uint8_t* const* payloads,
                                           uint32_t n_parties) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(which <= 1);
    WitnessState w;
    MPC_TRY(take_state(state, false, &w));
    MPC_TRY(check_device(w));
    if (w.finished) {
        set_error("witness_map state %llu was already finished", (unsigned long long)state);
        return MPC_CUDA_ERR_HANDLE;
    }
    Layout L(w);
    MPC_TRY(sum_payloads(payloads, n_parties, w.n, L.open + which * w.n, s));
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto it = g_ws.find(state);
    if (it != g_ws.end()) it->second.opened[which] = true;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_witness_map_mac_payload(uint64_t state, uint32_t which, uint32_t is_leader, uint8_t* out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(which <= 1 && out);
    WitnessState w;
    MPC_TRY(take_state(state, false, &w));
    MPC_TRY(check_device(w));
    if (w.planes != 2 || !w.opened[which]) {
        set_error("witness_map_mac_payload: needs an SPDZ state whose open %u arrived through _open_payloads", which);
        return MPC_CUDA_ERR_ARG;
    }
    Layout L(w);
    Scratch sd;
    Fr* dx;
    MPC_TRY(sd.alloc(&dx, w.n, s));
    MPC_TRY(mpc_cuda_spdz_mac_check_dev((const uint64_t*)(L.open + which * w.n), (const uint64_t*)((which ? L.mb : L.ma) + w.n),
                                        (uint64_t*)dx, w.n, is_leader, s));
    return payload_out(dx, w.n, out, s);
}

int32_t mpc_cuda_witness_map_mac_verify(uint64_t state, const uint8_t* const* payloads, uint32_t n_parties) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    WitnessState w;
    MPC_TRY(take_state(state, false, &w));
    MPC_TRY(check_device(w));
    Scratch sd, sf;
    Fr* sum;
    unsigned int* dflag;
    MPC_TRY(sd.alloc(&sum, w.n, s));
    MPC_TRY(sf.alloc(&dflag, 1, s));
    MPC_CUDA_TRY(cudaMemsetAsync(dflag, 0, sizeof(unsigned int), s));
    MPC_TRY(sum_payloads(payloads, n_parties, w.n, sum, s));
    k_any_nonzero<<<(unsigned)((w.n + 255) / 256), 256, 0, s>>>(sum, w.n, dflag);
    MPC_KERNEL_CHECK();
    unsigned int flag = 0;
    MPC_CUDA_TRY(cudaMemcpyAsync(&flag, dflag, sizeof(flag), cudaMemcpyDeviceToHost, s));
    MPC_CUDA_TRY(cudaStreamSynchronize(s));
    if (flag) {
        set_error("SPDZ MAC check failed on an opened value");
        return MPC_CUDA_ERR_MAC;
    }
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_witness_map_assignment_dev(uint64_t state, uint64_t** z_dev, size_t* cols) {
    MPC_TRY(enter(nullptr));
    MPC_ARG_CHECK(z_dev && cols);
    WitnessState w;
    MPC_TRY(take_state(state, false, &w));
    MPC_TRY(check_device(w));
    if (!w.cols) {
        set_error("witness_map state %llu was not begun from an assignment", (unsigned long long)state);
        return MPC_CUDA_ERR_HANDLE;
    }
    Layout L(w);
    *z_dev = (uint64_t*)L.z;
    *cols = w.cols;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_witness_map_release(uint64_t state) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    WitnessState w;
    MPC_TRY(take_state(state, true, &w));
    int cur = 0;
    MPC_CUDA_TRY(cudaGetDevice(&cur));
    if (cur == w.cuda_device) {
        // stream-ordered: everything this thread enqueued on its stream has finished with the buffer by then; readers
        // on other streams (the MSMs fed from h / the assignment) were synchronised by the caller
        MPC_CUDA_TRY(cudaFreeAsync(w.buf, s));
    } else {                                       // released from a thread bound to another device: synchronous
        MPC_CUDA_TRY(cudaSetDevice(w.cuda_device));
        cudaFree(w.buf);
        MPC_CUDA_TRY(cudaSetDevice(cur));
    }
    return MPC_CUDA_OK;
}

}  // extern "C"
