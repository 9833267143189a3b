// Multi-limb Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces (computes the same function as) the reference's 64-bit-limb CIOS code:
//   mul_assign        arkworks/algebra/ff/src/fields/arithmetic.rs:7-57
//   square_in_place   arkworks/algebra/ff/src/fields/arithmetic.rs:85-172
//   into_repr         arkworks/algebra/ff/src/fields/arithmetic.rs:59-83
//   add/sub/neg/dbl   arkworks/algebra/ff/src/fields/macros.rs:317-323,638-651,698-717
//   from_repr         arkworks/algebra/ff/src/fields/macros.rs:464-474
// R = 2^(32 N) equals the reference's 2^(64 N/2), so Montgomery bit patterns are identical and
// every value handed in or out is fully reduced in [0, p): results are bit-exact with the CPU path.
//
// Multiplication keeps the running total as two interleaved accumulators of 64-bit columns
// (`ev` aligned to even limbs, `od` to odd limbs) so each 32x32->64 product is one
// IMAD.WIDE.U32 with carry (mad.lo.cc + madc.hi.cc pair), 2N^2 wide MADs + O(N) per product.
#pragma once
#include "ptx.cuh"

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one(i);
        return r;
    }
    HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::mod(i);
        return r;
    }
    HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i];
        return acc == 0;
    }
    HD bool operator==(const Fp& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i] ^ o.v[i];
        return acc == 0;
    }
    HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

namespace fp_detail {

// acc[0..n) += a[0], a[2], a[4] ... * b as 64-bit columns; leaves the carry-out in CC
template <int n>
HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t b) {
    acc[0] = ptx::mad_lo_cc(a[0], b, acc[0]);
    acc[1] = ptx::madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
    for (int j = 2; j < n; j += 2) {
        acc[j] = ptx::madc_lo_cc(a[j], b, acc[j]);
        acc[j + 1] = ptx::madc_hi_cc(a[j], b, acc[j + 1]);
    }
}

// same with the modulus as the multiplicand: acc += p[off], p[off+2], ... * m
template <class P, int off>
HD void cmad_mod(uint32_t* acc, uint32_t m) {
    constexpr int n = P::N;
    acc[0] = ptx::mad_lo_cc(P::mod(off), m, acc[0]);
    acc[1] = ptx::madc_hi_cc(P::mod(off), m, acc[1]);
#pragma unroll
    for (int j = 2; j < n; j += 2) {
        acc[j] = ptx::madc_lo_cc(P::mod(off + j), m, acc[j]);
        acc[j + 1] = ptx::madc_hi_cc(P::mod(off + j), m, acc[j + 1]);
    }
}

// One Montgomery reduction step on (ev, od): adds m*p so that ev[0] becomes 0.
// The od chain cannot carry out (total < 2^(32(N+1))); the ev chain's carry lands on limb N = od[N-1].
template <class P>
HD void redc_row(uint32_t* ev, uint32_t* od) {
    constexpr int N = P::N;
    uint32_t m = ptx::mul_lo(ev[0], P::INV);
    cmad_mod<P, 1>(od, m);
    cmad_mod<P, 0>(ev, m);
    od[N - 1] = ptx::addc(od[N - 1], 0);
}

// First row: ev = a_even * b, od = a_odd * b (no carries between disjoint 64-bit columns).
template <int N>
HD void mul_row_first(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        ev[j] = ptx::mul_lo(a[j], b);
        ev[j + 1] = ptx::mul_hi(a[j], b);
        od[j] = ptx::mul_lo(a[j + 1], b);
        od[j + 1] = ptx::mul_hi(a[j + 1], b);
    }
}

// Subsequent rows.  On entry the total is  ev_prev + od_prev * 2^32  with ev_prev[0] == 0.
// Dividing by 2^32 turns od_prev into the new even-aligned accumulator `ev` and
// (ev_prev >> 64) into the new odd-aligned accumulator, stored in place in `od`;
// the stray limb ev_prev[1] is folded into ev[0], its carry entering the odd chain at limb 1.
template <int N>
HD void mul_row(uint32_t* ev /* = od_prev */, uint32_t* od /* = ev_prev */, const uint32_t* a, uint32_t b) {
    ev[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        od[j] = ptx::madc_lo_cc(a[j + 1], b, od[j + 2]);
        od[j + 1] = ptx::madc_hi_cc(a[j + 1], b, od[j + 3]);
    }
    od[N - 2] = ptx::madc_lo_cc(a[N - 1], b, 0);
    od[N - 1] = ptx::madc_hi(a[N - 1], b, 0);
    cmad_n<N>(ev, a, b);
    od[N - 1] = ptx::addc(od[N - 1], 0);
}

// r = (r >= p) ? r - p : r
template <class P>
HD void final_sub(uint32_t* r) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = ptx::sub_cc(r[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(r[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);      // 0xffffffff when r < p
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

}  // namespace fp_detail

// 32-bit formulation (mad.lo.cc/madc.hi.cc): ptxas 12.9 lowers it to IMAD + IMAD.HI.U32.X + IADD3.X,
// about 3 instructions per limb product.  Kept for A/B measurements (bench_fp kernels).
template <class P>
HD Fp<P> mul_narrow(const Fp<P>& a, const Fp<P>& b) {
    using namespace fp_detail;
    constexpr int N = P::N;
    uint32_t ev[N], od[N];
    mul_row_first<N>(ev, od, a.v, b.v[0]);
    redc_row<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mul_row<N>(od, ev, a.v, b.v[i]);        // roles swap every row
        redc_row<P>(od, ev);
        if (i + 1 < N) {
            mul_row<N>(ev, od, a.v, b.v[i + 1]);
            redc_row<P>(ev, od);
        }
    }
    // N is even: after the last (odd-indexed) row the even-aligned accumulator lives in `od`
    // and holds a zero low limb; result = ev + (od >> 32)
    Fp<P> r;
    r.v[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(ev[i], od[i + 1]);
    r.v[N - 1] = ptx::addc(ev[N - 1], 0);
    final_sub<P>(r.v);
    return r;
}

namespace fp_detail {

HD uint32_t lo32(uint64_t x) { return (uint32_t)x; }
HD uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
HD uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// 64-bit-column formulation: ev[k] is the column at limbs (2k, 2k+1), od[k] at (2k+1, 2k+2).
// Every limb product is one IMAD.WIDE.U32[.X] (see ptx.cuh).
template <class P>
HD void wredc_row(uint64_t* ev, uint64_t* od) {
    constexpr int H = P::N / 2;
    // both BLS12-377 moduli are 1 mod 2^32, so -p^-1 = -1: m is a negation and m * p_0 a plain 64-bit addition
    // (the low limb cancels to zero, its carry enters limb 1) - one multiplier slot less per row
    constexpr bool UNIT = P::mod(0) == 1u && P::INV == 0xffffffffu;
    uint32_t m = UNIT ? 0u - lo32(ev[0]) : ptx::mul_lo(lo32(ev[0]), P::INV);
    od[0] = ptx::madw_cc(P::mod(1), m, od[0]);
#pragma unroll
    for (int k = 1; k < H; k++) od[k] = ptx::madwc_cc(P::mod(2 * k + 1), m, od[k]);   // no carry out
    ev[0] = UNIT ? ptx::add64_cc(ev[0], (uint64_t)m) : ptx::madw_cc(P::mod(0), m, ev[0]);
#pragma unroll
    for (int k = 1; k < H; k++) ev[k] = ptx::madwc_cc(P::mod(2 * k), m, ev[k]);
    od[H - 1] = pack64(lo32(od[H - 1]), ptx::addc(hi32(od[H - 1]), 0));                  // carry -> limb N
}

template <int N>
HD void wmul_row_first(uint64_t* ev, uint64_t* od, const uint32_t* a, uint32_t b) {
#pragma unroll
    for (int k = 0; k < N / 2; k++) {
        ev[k] = ptx::mulw(a[2 * k], b);
        od[k] = ptx::mulw(a[2 * k + 1], b);
    }
}

// ev = od_prev, od = ev_prev (ev_prev's low limb is zero); see mul_row above for the derivation
template <int N>
HD void wmul_row(uint64_t* ev, uint64_t* od, const uint32_t* a, uint32_t b) {
    constexpr int H = N / 2;
    ev[0] = pack64(ptx::add_cc(lo32(ev[0]), hi32(od[0])), hi32(ev[0]));   // stray limb; carry enters limb 1
#pragma unroll
    for (int k = 0; k < H - 1; k++) od[k] = ptx::madwc_cc(a[2 * k + 1], b, od[k + 1]);
    od[H - 1] = ptx::madwc(a[N - 1], b, 0);
    ev[0] = ptx::madw_cc(a[0], b, ev[0]);
#pragma unroll
    for (int k = 1; k < H; k++) ev[k] = ptx::madwc_cc(a[2 * k], b, ev[k]);
    od[H - 1] = pack64(lo32(od[H - 1]), ptx::addc(hi32(od[H - 1]), 0));
}

}  // namespace fp_detail

template <class P>
HD Fp<P> mul_wide(const Fp<P>& a, const Fp<P>& b) {
    using namespace fp_detail;
    constexpr int N = P::N, H = N / 2;
    uint64_t ev[H], od[H];
    wmul_row_first<N>(ev, od, a.v, b.v[0]);
    wredc_row<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        wmul_row<N>(od, ev, a.v, b.v[i]);
        wredc_row<P>(od, ev);
        if (i + 1 < N) {
            wmul_row<N>(ev, od, a.v, b.v[i + 1]);
            wredc_row<P>(ev, od);
        }
    }
    // result = ev + (od >> 32), od's low limb being zero
    Fp<P> r;
    r.v[0] = ptx::add_cc(lo32(ev[0]), hi32(od[0]));
#pragma unroll
    for (int i = 1; i < N - 1; i++) {
        uint32_t e = (i & 1) ? hi32(ev[i / 2]) : lo32(ev[i / 2]);
        uint32_t o = ((i + 1) & 1) ? hi32(od[(i + 1) / 2]) : lo32(od[(i + 1) / 2]);
        r.v[i] = ptx::addc_cc(e, o);
    }
    r.v[N - 1] = ptx::addc(hi32(ev[H - 1]), 0);
    final_sub<P>(r.v);
    return r;
}

template <class P>
HD Fp<P> mul(const Fp<P>& a, const Fp<P>& b) {
#ifdef FP_MUL_NARROW
    return mul_narrow(a, b);
#else
    return mul_wide(a, b);
#endif
}

namespace fp_detail {

// window shift of the Montgomery reduction without a product row: new total =
// od_prev + (ev_prev >> 32) + top * 2^(32(N-1)); called as wshift_row(od_prev, ev_prev, top) like wmul_row
template <int N>
HD void wshift_row(uint64_t* ev, uint64_t* od, uint32_t top) {
    constexpr int H = N / 2;
    ev[0] = pack64(ptx::add_cc(lo32(ev[0]), hi32(od[0])), hi32(ev[0]));   // stray limb; carry enters limb 1
#pragma unroll
    for (int k = 0; k < H - 1; k++) od[k] = ptx::addc64_cc(od[k + 1], 0);
    od[H - 1] = ptx::addc64(pack64(top, 0), 0);
}

}  // namespace fp_detail

// Dedicated Montgomery squaring (the function of square_in_place, ff/src/fields/arithmetic.rs:85-172):
// off-diagonal products once (N(N-1)/2 wide MACs), doubled, plus the N diagonal squares, then a
// product-free Montgomery reduction: N(N+1)/2 + N^2 wide MACs instead of 2 N^2.
template <class P>
HD Fp<P> sqr_wide(const Fp<P>& a) {
    using namespace fp_detail;
    constexpr int N = P::N, H = N / 2;
    // E[k]: 64-bit column over limbs (2k, 2k+1);  O[k]: column over limbs (2k+1, 2k+2)
    uint64_t E[N], O[N];
#pragma unroll
    for (int k = 0; k < N; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
        // a_i * a_j, j - i odd: limbs i+j = 2i+1, 2i+3, ... -> O[i], O[i+1], ...
        {
            int k = i;
            O[k] = ptx::madw_cc(a.v[i], a.v[i + 1], O[k]);
            k++;
#pragma unroll
            for (int j = i + 3; j < N; j += 2, k++) O[k] = ptx::madwc_cc(a.v[i], a.v[j], O[k]);
            O[k] = ptx::addc64(O[k], 0);        // columns above a row's top only hold earlier carries
        }
        // j - i even: limbs 2i+2, 2i+4, ... -> E[i+1], E[i+2], ...
        if (i + 2 < N) {
            int k = i + 1;
            E[k] = ptx::madw_cc(a.v[i], a.v[i + 2], E[k]);
            k++;
#pragma unroll
            for (int j = i + 4; j < N; j += 2, k++) E[k] = ptx::madwc_cc(a.v[i], a.v[j], E[k]);
            E[k] = ptx::addc64(E[k], 0);
        }
    }
    // t = E + (O << 32) as 2N limbs
    uint32_t t[2 * N];
    t[0] = lo32(E[0]);
    t[1] = ptx::add_cc(hi32(E[0]), lo32(O[0]));
#pragma unroll
    for (int l = 2; l < 2 * N - 1; l++) {
        uint32_t e = (l & 1) ? hi32(E[l / 2]) : lo32(E[l / 2]);
        uint32_t o = ((l - 1) & 1) ? hi32(O[(l - 1) / 2]) : lo32(O[(l - 1) / 2]);
        t[l] = ptx::addc_cc(e, o);
    }
    t[2 * N - 1] = ptx::addc(hi32(E[N - 1]), lo32(O[N - 1]));
    // t = 2 t
#pragma unroll
    for (int l = 2 * N - 1; l > 0; l--) t[l] = (t[l] << 1) | (t[l - 1] >> 31);
    t[0] <<= 1;
    // t += sum a_k^2 2^(64 k)
    uint64_t col[N];
    col[0] = ptx::madw_cc(a.v[0], a.v[0], pack64(t[0], t[1]));
#pragma unroll
    for (int k = 1; k < N - 1; k++) col[k] = ptx::madwc_cc(a.v[k], a.v[k], pack64(t[2 * k], t[2 * k + 1]));
    col[N - 1] = ptx::madwc(a.v[N - 1], a.v[N - 1], pack64(t[2 * N - 2], t[2 * N - 1]));
    // Montgomery reduction of the 2N-limb square: the low half seeds the window, the high limbs enter
    // at the top as the window slides
    uint64_t ev[H], od[H];
#pragma unroll
    for (int k = 0; k < H; k++) { ev[k] = col[k]; od[k] = 0; }
    auto hi_limb = [&](int l) -> uint32_t { return (l & 1) ? hi32(col[l / 2]) : lo32(col[l / 2]); };
    wredc_row<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        wshift_row<N>(od, ev, hi_limb(N + i - 1));
        wredc_row<P>(od, ev);
        if (i + 1 < N) {
            wshift_row<N>(ev, od, hi_limb(N + i));
            wredc_row<P>(ev, od);
        }
    }
    Fp<P> r;
    r.v[0] = ptx::add_cc(lo32(ev[0]), hi32(od[0]));
#pragma unroll
    for (int i = 1; i < N - 1; i++) {
        uint32_t e = (i & 1) ? hi32(ev[i / 2]) : lo32(ev[i / 2]);
        uint32_t o = ((i + 1) & 1) ? hi32(od[(i + 1) / 2]) : lo32(od[(i + 1) / 2]);
        r.v[i] = ptx::addc_cc(e, o);
    }
    r.v[N - 1] = ptx::addc(hi32(ev[H - 1]), hi_limb(2 * N - 1));
    final_sub<P>(r.v);
    return r;
}

template <class P>
HD Fp<P> sqr(const Fp<P>& a) {
    // sqr_wide saves 23 % of the wide MACs but adds shifts and 64-bit adds on the ALU pipe; inside the
    // bucket-accumulation kernel it measured 2 % slower than mul(a, a) on B200 (tools/tune_msm.py), so it
    // is opt-in until the reduction is restructured
#ifdef FP_SQR_DEDICATED
    return sqr_wide(a);
#else
    return mul(a, a);
#endif
}

template <class P>
HD Fp<P> add(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);      // p has spare top bits: no overflow
    fp_detail::final_sub<P>(r.v);
    return r;
}

template <class P>
HD Fp<P> sub(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = ptx::subc(0, 0);      // all-ones when a < b
    r.v[0] = ptx::add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod(i) & borrow);
    r.v[N - 1] = ptx::addc(r.v[N - 1], P::mod(N - 1) & borrow);
    return r;
}

template <class P>
HD Fp<P> dbl(const Fp<P>& a) {
    return add(a, a);
}

template <class P>
HD Fp<P> neg(const Fp<P>& a) {
    constexpr int N = P::N;
    Fp<P> r;
    r.v[0] = ptx::sub_cc(P::mod(0), a.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::subc_cc(P::mod(i), a.v[i]);
    r.v[N - 1] = ptx::subc(P::mod(N - 1), a.v[N - 1]);
    uint32_t keep = a.is_zero() ? 0u : 0xffffffffu;       // -0 = 0
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] &= keep;
    return r;
}

// Montgomery form -> canonical integer: one Montgomery product with 1 (value * R^-1)
template <class P>
HD Fp<P> from_mont(const Fp<P>& a) {
    Fp<P> one_raw = Fp<P>::zero();
    one_raw.v[0] = 1;
    return mul(a, one_raw);
}

// canonical integer (< p) -> Montgomery form
template <class P>
HD Fp<P> to_mont(const Fp<P>& a) {
    Fp<P> r2;
#pragma unroll
    for (int i = 0; i < P::N; i++) r2.v[i] = P::r2(i);
    return mul(a, r2);
}

// a^(p-2) by square-and-multiply over the constant exponent; 0 -> 0.
// (The reference uses binary Euclid, macros.rs:389-443; the value is the same.)
template <class P>
HD Fp<P> inv(const Fp<P>& a) {
    constexpr int N = P::N;
    Fp<P> acc = Fp<P>::one();
    bool started = false;
    for (int i = N * 32 - 1; i >= 0; i--) {
        uint32_t bit = (P::pm2(i / 32) >> (i % 32)) & 1u;      // exponent p - 2
        if (started) acc = sqr(acc);
        if (bit) {
            acc = mul(acc, a);
            started = true;
        }
    }
    return acc;
}

// Binary extended Euclid, the algorithm of the reference's `inverse` (ff/src/fields/macros.rs:389-443:
// Guajardo-Kumar-Paar-Pelzl, u = a, v = p, b = R^2, c = 0 on plain integers), so a Montgomery input xR
// yields x^-1 R.  About 2 log2(p) shift/subtract rounds of O(N) word operations: ~20x fewer instructions
// than the Fermat ladder above, but data-dependent control flow — meant for single-thread tails (result
// normalisation) where latency, not divergence, is what matters.  0 -> 0.
namespace fp_detail {
template <int N> HD void shr1(uint32_t* a) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] >>= 1;
}
template <int N> HD bool is_one(const uint32_t* a) {
    uint32_t acc = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < N; i++) acc |= a[i];
    return acc == 0;
}
template <int N> HD bool less_than(const uint32_t* a, const uint32_t* b) {      // a < b
    ptx::sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N; i++) ptx::subc_cc(a[i], b[i]);
    return ptx::subc(0, 0) != 0;
}
template <int N> HD void sub_raw(uint32_t* a, const uint32_t* b) {              // a -= b, a >= b
    a[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) a[i] = ptx::subc_cc(a[i], b[i]);
    a[N - 1] = ptx::subc(a[N - 1], b[N - 1]);
}
// b = b / 2 mod p for b in [0, p): b even -> b >> 1, else (b + p) >> 1 (p has spare top bits)
template <class P> HD void half_mod(uint32_t* b) {
    constexpr int N = P::N;
    if (b[0] & 1u) {
        b[0] = ptx::add_cc(b[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) b[i] = ptx::addc_cc(b[i], P::mod(i));
        b[N - 1] = ptx::addc(b[N - 1], P::mod(N - 1));
    }
    shr1<N>(b);
}
}  // namespace fp_detail

template <class P>
HD Fp<P> inv_euclid(const Fp<P>& a) {
    using namespace fp_detail;
    constexpr int N = P::N;
    if (a.is_zero()) return Fp<P>::zero();
    Fp<P> u = a, v = Fp<P>::modulus(), b, c = Fp<P>::zero();
#pragma unroll
    for (int i = 0; i < N; i++) b.v[i] = P::r2(i);
    while (!is_one<N>(u.v) && !is_one<N>(v.v)) {
        while (!(u.v[0] & 1u)) { shr1<N>(u.v); half_mod<P>(b.v); }
        while (!(v.v[0] & 1u)) { shr1<N>(v.v); half_mod<P>(c.v); }
        if (less_than<N>(v.v, u.v)) { sub_raw<N>(u.v, v.v); b = sub(b, c); }
        else { sub_raw<N>(v.v, u.v); c = sub(c, b); }
    }
    return is_one<N>(u.v) ? b : c;
}

// ---- lazy reduction for the NTT butterflies --------------------------------------------------------------
// Values are kept in [0, 2p) between butterflies instead of [0, p): p has at least 3 spare bits in its top limb,
// so 4p fits the limbs.  u + v is brought back below 2p with one conditional subtraction of 2p; u - v is
// computed as u - v + 2p in (0, 4p) and only ever feeds a Montgomery product with a fully reduced twiddle,
// whose result a * w / R + p * (m / R) < 4p * p / R + p < 2p needs no final subtraction at all (p / R < 1/8).
// reduce_full brings a value from [0, 4p) to the canonical [0, p) where the transform ends.
namespace fp_detail {
template <class P> HD uint32_t mod2(int i) {      // limb i of 2p
    return (P::mod(i) << 1) | (i ? (P::mod(i - 1) >> 31) : 0u);
}
// r = (r >= m) ? r - m : r for m = 2p
template <class P>
HD void cond_sub_2p(uint32_t* r) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = ptx::sub_cc(r[0], mod2<P>(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(r[i], mod2<P>(i));
    uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}
}  // namespace fp_detail

// a, b in [0, 2p) -> a + b in [0, 2p)
template <class P>
HD Fp<P> add_lazy(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);
    fp_detail::cond_sub_2p<P>(r.v);
    return r;
}

// a, b in [0, 2p) -> a - b + 2p in (0, 4p): input of mul_lazy only
template <class P>
HD Fp<P> sub_lazy(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = ptx::subc(a.v[N - 1], b.v[N - 1]);
    r.v[0] = ptx::add_cc(r.v[0], fp_detail::mod2<P>(0));
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], fp_detail::mod2<P>(i));
    r.v[N - 1] = ptx::addc(r.v[N - 1], fp_detail::mod2<P>(N - 1));
    return r;
}

// Montgomery product of a in [0, 4p) with b in [0, p), without the final subtraction: result in [0, 2p)
template <class P>
HD Fp<P> mul_lazy(const Fp<P>& a, const Fp<P>& b) {
    using namespace fp_detail;
    constexpr int N = P::N, H = N / 2;
    uint64_t ev[H], od[H];
    wmul_row_first<N>(ev, od, a.v, b.v[0]);
    wredc_row<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        wmul_row<N>(od, ev, a.v, b.v[i]);
        wredc_row<P>(od, ev);
        if (i + 1 < N) {
            wmul_row<N>(ev, od, a.v, b.v[i + 1]);
            wredc_row<P>(ev, od);
        }
    }
    Fp<P> r;
    r.v[0] = ptx::add_cc(lo32(ev[0]), hi32(od[0]));
#pragma unroll
    for (int i = 1; i < N - 1; i++) {
        uint32_t e = (i & 1) ? hi32(ev[i / 2]) : lo32(ev[i / 2]);
        uint32_t o = ((i + 1) & 1) ? hi32(od[(i + 1) / 2]) : lo32(od[(i + 1) / 2]);
        r.v[i] = ptx::addc_cc(e, o);
    }
    r.v[N - 1] = ptx::addc(hi32(ev[H - 1]), 0);
    return r;
}

// [0, 4p) -> [0, p)
template <class P>
HD Fp<P> reduce_full(const Fp<P>& a) {
    Fp<P> r = a;
    fp_detail::cond_sub_2p<P>(r.v);
    fp_detail::final_sub<P>(r.v);
    return r;
}
