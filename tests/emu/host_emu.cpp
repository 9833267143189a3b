// Host-side emulation harness: compiles the device templates (fp.cuh, fp2.cuh, ec.cuh, ...) with
// plain g++ using the carry-flag emulation in ptx.cuh, so their logic can be checked against the
// oracle on a CPU-only box.  Test scaffolding only — never linked into libmpc_cuda.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "consts.cuh"
#include "fp.cuh"
#include "fp2.cuh"
#include "ec.cuh"
#include "msm_digits.cuh"

using Fr = Fp<consts::FrParams>;
using Fq = Fp<consts::FqParams>;
using Fq2 = Fp2<consts::FqParams>;

template <class P> static Fp<P> sqr_wide_of(const Fp<P>& x) { return sqr_wide(x); }
template <class P> static Fp2<P> sqr_wide_of(const Fp2<P>& x) { return sqr(x); }

// the NTT's lazily reduced butterfly arithmetic against the strict one: operands lifted to [p, 2p), then
// (x - y) * y and x + y through add_lazy / sub_lazy / mul_lazy, brought back with reduce_full; must equal
// mul(sub(x, y), y) + add(x, y) computed strictly
template <class P> static Fp<P> lazy_chain(const Fp<P>& x, const Fp<P>& y) {
    Fp<P> xl = add_lazy(x, Fp<P>::modulus()), yl = add_lazy(y, Fp<P>::modulus());      // x + p, y + p (< 2p)
    Fp<P> prod = mul_lazy(sub_lazy(xl, yl), y);                                        // < 2p
    Fp<P> sum = add_lazy(xl, yl);                                                      // < 2p
    Fp<P> both = add_lazy(prod, sum);
    return reduce_full(both);
}
template <class P> static Fp2<P> lazy_chain(const Fp2<P>& x, const Fp2<P>&) { return x; }

template <class F>
static void vec_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    constexpr int N = F::N;
    for (size_t i = 0; i < n; i++) {
        F x, y, r;
        memcpy(&x, a + N * i, 4 * N);
        if (b) memcpy(&y, b + N * i, 4 * N);
        switch (op) {
            case 0: r = add(x, y); break;
            case 1: r = sub(x, y); break;
            case 2: r = mul(x, y); break;
            case 3: r = neg(x); break;
            case 4: r = inv(x); break;
            case 7: r = sqr(x); break;
            case 10: r = sqr_wide_of(x); break;
            case 9: r = dbl(x); break;
            case 11: r = inv_euclid(x); break;
            case 12: r = lazy_chain(x, y); break;
            default: r = x; break;
        }
        memcpy(out + N * i, &r, 4 * N);
    }
}

template <class F>
static void vec_op_mont(int op, const uint32_t* a, uint32_t* out, size_t n) {
    constexpr int N = F::N;
    for (size_t i = 0; i < n; i++) {
        F x, r;
        memcpy(&x, a + N * i, 4 * N);
        r = op == 5 ? from_mont(x) : to_mont(x);
        memcpy(out + N * i, &r, 4 * N);
    }
}

// sum_i (+-)P_i three ways: a madd chain; two half chains joined by xyzz_add; the madd chain sent
// through Jacobian and back and doubled-and-halved via xyzz_add with itself (exercises dbl).
// out_xy[0]: chain, out_xy[1]: halves, out_xy[2]: (2*chain) via add(p,p)  (affine, N limbs x 2 each)
template <class F>
static void ec_sums(const uint32_t* pts, const uint8_t* negate, size_t n, uint32_t* out_xy, uint8_t* out_inf) {
    constexpr int N = F::N;
    XYZZ<F> acc = XYZZ<F>::infinity(), lo = XYZZ<F>::infinity(), hi = XYZZ<F>::infinity();
    for (size_t i = 0; i < n; i++) {
        F x, y;
        memcpy(&x, pts + 2 * N * i, 4 * N);
        memcpy(&y, pts + 2 * N * i + N, 4 * N);
        if (negate && negate[i]) y = neg(y);
        xyzz_madd(acc, x, y);
        xyzz_madd(i < n / 2 ? lo : hi, x, y);
    }
    xyzz_add(lo, hi);
    XYZZ<F> twice = jac_to_xyzz(xyzz_to_jac(acc));
    XYZZ<F> same = twice;
    xyzz_add(twice, same);
    XYZZ<F>* res[3] = {&acc, &lo, &twice};
    for (int k = 0; k < 3; k++) {
        F ox, oy;
        out_inf[k] = xyzz_to_affine(*res[k], ox, oy) ? 0 : 1;
        memcpy(out_xy + 2 * N * k, &ox, 4 * N);
        memcpy(out_xy + 2 * N * k + N, &oy, 4 * N);
    }
}

template <class F>
static void ec_mul_small(const uint32_t* pt, uint64_t k, uint32_t* out_xy, uint8_t* out_inf) {
    constexpr int N = F::N;
    XYZZ<F> p;
    memcpy(&p.x, pt, 4 * N);
    memcpy(&p.y, pt + N, 4 * N);
    p.zz = F::one();
    p.zzz = F::one();
    XYZZ<F> r = xyzz_mul_small(p, k);
    F ox, oy;
    *out_inf = xyzz_to_affine(r, ox, oy) ? 0 : 1;
    memcpy(out_xy, &ox, 4 * N);
    memcpy(out_xy + N, &oy, 4 * N);
}

// 2^k * P through the Jacobian doubling of the window-table builder, normalised with the Euclid inversion
template <class F>
static void ec_jac_dbl_chain(const uint32_t* pt, uint32_t k, uint32_t* out_xy) {
    constexpr int N = F::N;
    Jac<F> j;
    memcpy(&j.x, pt, 4 * N);
    memcpy(&j.y, pt + N, 4 * N);
    j.z = F::one();
    for (uint32_t i = 0; i < k; i++) jac_dbl(j);
    F zi = inv_euclid(j.z), zi2 = sqr(zi);
    F ox = mul(j.x, zi2), oy = mul(j.y, mul(zi2, zi));
    memcpy(out_xy, &ox, 4 * N);
    memcpy(out_xy + N, &oy, 4 * N);
}

extern "C" {
void emu_g1_jac_dbl_chain(const uint32_t* pt, uint32_t k, uint32_t* out_xy) { ec_jac_dbl_chain<Fq>(pt, k, out_xy); }
void emu_g2_jac_dbl_chain(const uint32_t* pt, uint32_t k, uint32_t* out_xy) { ec_jac_dbl_chain<Fq2>(pt, k, out_xy); }
void emu_fr_vec(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    if (op == 5 || op == 6) vec_op_mont<Fr>(op, a, out, n); else vec_op<Fr>(op, a, b, out, n);
}
void emu_fq_vec(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    if (op == 5 || op == 6) vec_op_mont<Fq>(op, a, out, n); else vec_op<Fq>(op, a, b, out, n);
}
void emu_fq2_vec(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) { vec_op<Fq2>(op, a, b, out, n); }
void emu_g1_sums(const uint32_t* pts, const uint8_t* negate, size_t n, uint32_t* out_xy, uint8_t* out_inf) { ec_sums<Fq>(pts, negate, n, out_xy, out_inf); }
void emu_g2_sums(const uint32_t* pts, const uint8_t* negate, size_t n, uint32_t* out_xy, uint8_t* out_inf) { ec_sums<Fq2>(pts, negate, n, out_xy, out_inf); }
void emu_g1_mul_small(const uint32_t* pt, uint64_t k, uint32_t* out_xy, uint8_t* out_inf) { ec_mul_small<Fq>(pt, k, out_xy, out_inf); }
void emu_g2_mul_small(const uint32_t* pt, uint64_t k, uint32_t* out_xy, uint8_t* out_inf) { ec_mul_small<Fq2>(pt, k, out_xy, out_inf); }

// signed-digit decomposition of canonical 256-bit scalars (msm_digits.cuh); out[w*n + i]
void emu_signed_digits(const uint32_t* canon, size_t n, uint32_t c, uint32_t nwin, uint32_t* out) {
    for (size_t i = 0; i < n; i++) {
        uint32_t carry = 0;
        for (uint32_t w = 0; w < nwin; w++) out[(size_t)w * n + i] = msm::signed_digit(canon + 8 * i, w, c, nwin, carry);
    }
}
}

// the k_generate loop of msm.cu (double-and-madd over a 64-bit multiplier) on the host
template <class F>
static void ec_generate(const uint32_t* gx, const uint32_t* gy, uint64_t k, uint32_t* out_xy) {
    constexpr int N = F::N;
    F x, y;
    memcpy(&x, gx, 4 * N);
    memcpy(&y, gy, 4 * N);
    XYZZ<F> acc = XYZZ<F>::infinity();
    bool started = false;
    for (int b = 63; b >= 0; b--) {
        if (started) xyzz_dbl(acc);
        if ((k >> b) & 1) { xyzz_madd(acc, x, y); started = true; }
    }
    F ox, oy;
    xyzz_to_affine(acc, ox, oy);
    memcpy(out_xy, &ox, 4 * N);
    memcpy(out_xy + N, &oy, 4 * N);
}
extern "C" {
void emu_g1_generate(uint64_t k, uint32_t* out_xy) { ec_generate<Fq>(consts::G1_GEN_X, consts::G1_GEN_Y, k, out_xy); }
void emu_g2_generate(uint64_t k, uint32_t* out_xy) { ec_generate<Fq2>(consts::G2_GEN_X, consts::G2_GEN_Y, k, out_xy); }
}
