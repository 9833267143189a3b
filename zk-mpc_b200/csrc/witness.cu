// Fused R1CStoQAP::witness_map on one party's local share values (src/groth16.rs:278-303): the three
// vectors stay resident in HBM across 3 iFFT, 3 coset FFT, the Beaver mask, and — after the two network
// opens that stay on the host's mpc-net — the Beaver combine, the subtraction, the division by the
// vanishing polynomial and the coset iFFT.  Only the 2 x n masked values leave the device in between
// (SURVEY.md §8 f1).  Everything is composed from the library's own stream-ordered entry points.
#include <mutex>
#include <unordered_map>

#include "common.cuh"

using namespace mpc;

namespace {

struct WitnessState {
    Fr* buf = nullptr;          // [a' | b' | c' | tx | ty] 5 x n, later reused for the combine
    size_t n = 0;
    uint32_t log_n = 0;
    int cuda_device = 0;
};
std::mutex g_ws_mu;
std::unordered_map<uint64_t, WitnessState> g_ws;
uint64_t g_ws_next = 1;

}  // namespace

extern "C" {

int32_t mpc_cuda_witness_map_begin(const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                                   const uint64_t* tx, const uint64_t* ty, uint64_t* masked_a, uint64_t* masked_b,
                                   uint64_t* state) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    MPC_ARG_CHECK(a && b && c && tx && ty && masked_a && masked_b && state && log_n <= 28);
    const size_t n = (size_t)1 << log_n, bytes = n * sizeof(Fr);
    WitnessState w;
    w.n = n;
    w.log_n = log_n;
    w.cuda_device = current_device_info()->cuda_device;
    MPC_CUDA_TRY(cudaMalloc((void**)&w.buf, 7 * bytes));
    Fr *va = w.buf, *vb = va + n, *vc = vb + n, *vtx = vc + n, *vty = vtx + n, *ma = vty + n, *mb = ma + n;
    int32_t rc = MPC_CUDA_OK;
    auto fail = [&](int32_t r) { cudaFree(w.buf); return r; };
    if (cudaMemcpyAsync(va, a, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(vb, b, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(vc, c, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(vtx, tx, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(vty, ty, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) {
        set_error("witness_map_begin: host to device copy failed");
        return fail(MPC_CUDA_ERR_CUDA);
    }
    // ifft then coset_fft of a, b, c: three equal-size transforms per call (batch = 3)
    if ((rc = mpc_cuda_ntt_fr_dev((uint64_t*)va, log_n, MPC_CUDA_NTT_IFFT, 3, s)) != MPC_CUDA_OK) return fail(rc);
    if ((rc = mpc_cuda_ntt_fr_dev((uint64_t*)va, log_n, MPC_CUDA_NTT_COSET_FFT, 3, s)) != MPC_CUDA_OK) return fail(rc);
    // Beaver masks of the batch product a' * b' (share/field.rs:108-117): one launch over both planes
    if ((rc = mpc_cuda_beaver_mask_dev((const uint64_t*)va, (const uint64_t*)vtx, (uint64_t*)ma, 2 * n, s)) != MPC_CUDA_OK)
        return fail(rc);
    if (cudaMemcpyAsync(masked_a, ma, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(masked_b, mb, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        set_error("witness_map_begin: device to host copy failed");
        return fail(MPC_CUDA_ERR_CUDA);
    }
    std::lock_guard<std::mutex> lk(g_ws_mu);
    *state = g_ws_next++;
    g_ws[*state] = w;
    return MPC_CUDA_OK;
}

int32_t mpc_cuda_witness_map_finish(uint64_t state, const uint64_t* tz, const uint64_t* sx, const uint64_t* oy,
                                    uint32_t is_leader, uint64_t* h_out) {
    cudaStream_t s;
    MPC_TRY(enter(&s));
    WitnessState w;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        auto it = g_ws.find(state);
        if (it == g_ws.end()) {
            set_error("unknown witness_map state %llu", (unsigned long long)state);
            return MPC_CUDA_ERR_HANDLE;
        }
        w = it->second;
        g_ws.erase(it);
    }
    int32_t rc = MPC_CUDA_OK;
    auto done = [&](int32_t r) { cudaFree(w.buf); return r; };
    if (!(tz && sx && oy && h_out)) {
        set_error("witness_map_finish: null argument");
        return done(MPC_CUDA_ERR_ARG);
    }
    if (w.cuda_device != current_device_info()->cuda_device) {
        set_error("witness_map state lives on another device");
        return done(MPC_CUDA_ERR_HANDLE);
    }
    const size_t n = w.n, bytes = n * sizeof(Fr);
    Fr *va = w.buf, *vb = va + n, *vc = vb + n, *vtx = vc + n, *vty = vtx + n, *vz = vty + n, *vs = vz + n;
    // a', b' are no longer needed: reuse them for the opened values
    if (cudaMemcpyAsync(vz, tz, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(va, sx, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(vb, oy, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) {
        set_error("witness_map_finish: host to device copy failed");
        return done(MPC_CUDA_ERR_CUDA);
    }
    // ab = z - y*sx - x*oy (+ sx*oy on the leader)          share/field.rs:118-128
    if ((rc = mpc_cuda_beaver_combine_dev((const uint64_t*)vtx, (const uint64_t*)vty, (const uint64_t*)vz,
                                          (const uint64_t*)va, (const uint64_t*)vb, (uint64_t*)vs, n, is_leader, 0, s)))
        return done(rc);
    // ab -= c' ; ab /= Z_H on the coset ; h = coset_ifft(ab)   src/groth16.rs:298-303
    if ((rc = mpc_cuda_vec_op_dev(MPC_CUDA_VEC_SUB, (const uint64_t*)vs, (const uint64_t*)vc, nullptr, (uint64_t*)vs, n, s)))
        return done(rc);
    if ((rc = mpc_cuda_divide_by_vanishing_on_coset_dev((uint64_t*)vs, w.log_n, s))) return done(rc);
    if ((rc = mpc_cuda_ntt_fr_dev((uint64_t*)vs, w.log_n, MPC_CUDA_NTT_COSET_IFFT, 1, s))) return done(rc);
    if (cudaMemcpyAsync(h_out, vs, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        set_error("witness_map_finish: device to host copy failed");
        return done(MPC_CUDA_ERR_CUDA);
    }
    return done(MPC_CUDA_OK);
}

}  // extern "C"
