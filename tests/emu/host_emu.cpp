// Host-side emulation harness: compiles the device templates (fp.cuh, fp2.cuh, ec.cuh, ...) with
// plain g++ using the carry-flag emulation in ptx.cuh, so their logic can be checked against the
// oracle on a CPU-only box.  Test scaffolding only — never linked into libmpc_cuda.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "consts.cuh"
#include "fp.cuh"

using Fr = Fp<consts::FrParams>;
using Fq = Fp<consts::FqParams>;

template <class F>
static void vec_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    constexpr int N = F::N;
    for (size_t i = 0; i < n; i++) {
        F x, y, r;
        memcpy(x.v, a + N * i, 4 * N);
        if (b) memcpy(y.v, b + N * i, 4 * N);
        switch (op) {
            case 0: r = add(x, y); break;
            case 1: r = sub(x, y); break;
            case 2: r = mul(x, y); break;
            case 3: r = neg(x); break;
            case 4: r = inv(x); break;
            case 5: r = from_mont(x); break;
            case 6: r = to_mont(x); break;
            default: r = sqr(x); break;
        }
        memcpy(out + N * i, r.v, 4 * N);
    }
}

extern "C" {
void emu_fr_vec(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) { vec_op<Fr>(op, a, b, out, n); }
void emu_fq_vec(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) { vec_op<Fq>(op, a, b, out, n); }
}
