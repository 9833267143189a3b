// Quadratic extension Fq2 = Fq[u]/(u^2 + 5) of the BLS12-377 base field, on top of fp.cuh.
//
// Computes the same functions as the reference's
//   QuadExtField mul/square/inverse  arkworks/algebra/ff/src/fields/models/quadratic_extension.rs:228-330,632-643
//   Fq2Parameters (NONRESIDUE = -5)  arkworks/curves/bls12_377/src/fields/fq2.rs:13,29-34
// Values are exact field elements with both coordinates fully reduced, so any correct formula is
// bit-identical to the CPU path; the formulas here are chosen for the GPU (3 / 2 base products).
#pragma once
#include "fp.cuh"

template <class P>
struct Fp2 {
    using Base = Fp<P>;
    static constexpr int N = 2 * P::N;      // 32-bit limbs per element (c0 | c1)
    Base c0, c1;

    HD static Fp2 zero() { Fp2 r; r.c0 = Base::zero(); r.c1 = Base::zero(); return r; }
    HD static Fp2 one() { Fp2 r; r.c0 = Base::one(); r.c1 = Base::zero(); return r; }
    HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    HD bool operator!=(const Fp2& o) const { return !(*this == o); }
};

template <class P> HD Fp2<P> add(const Fp2<P>& a, const Fp2<P>& b) { Fp2<P> r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); return r; }
template <class P> HD Fp2<P> sub(const Fp2<P>& a, const Fp2<P>& b) { Fp2<P> r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); return r; }
template <class P> HD Fp2<P> dbl(const Fp2<P>& a) { Fp2<P> r; r.c0 = dbl(a.c0); r.c1 = dbl(a.c1); return r; }
template <class P> HD Fp2<P> neg(const Fp2<P>& a) { Fp2<P> r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }

// 5 * a
template <class P> HD Fp<P> mul5(const Fp<P>& a) { return add(dbl(dbl(a)), a); }

// Karatsuba: (a0 + a1 u)(b0 + b1 u) = (a0 b0 - 5 a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u
// mul_fast / sqr_fast are the inlined bodies (the bucket-accumulation kernel: operands stay in registers);
// mul / sqr are the out-of-line versions every other G2 kernel calls (compile time, see ptx.cuh).
template <class P>
HD Fp2<P> mul_fast(const Fp2<P>& a, const Fp2<P>& b) {
    Fp<P> v0 = mul(a.c0, b.c0);
    Fp<P> v1 = mul(a.c1, b.c1);
    Fp<P> s = mul(add(a.c0, a.c1), add(b.c0, b.c1));
    Fp2<P> r;
    r.c1 = sub(sub(s, v0), v1);
    r.c0 = sub(v0, mul5(v1));
    return r;
}
template <class P>
HD_NOINLINE Fp2<P> mul(const Fp2<P>& a, const Fp2<P>& b) { return mul_fast(a, b); }

// complex squaring: c1 = 2 a0 a1, c0 = (a0 + a1)(a0 - 5 a1) + 4 a0 a1
template <class P>
HD Fp2<P> sqr_fast(const Fp2<P>& a) {
    Fp<P> v = mul(a.c0, a.c1);
    Fp<P> t = mul(add(a.c0, a.c1), sub(a.c0, mul5(a.c1)));
    Fp2<P> r;
    r.c1 = dbl(v);
    r.c0 = add(t, dbl(r.c1));
    return r;
}
template <class P>
HD_NOINLINE Fp2<P> sqr(const Fp2<P>& a) { return sqr_fast(a); }

// base field: the inlined product is the only one
template <class P> HD Fp<P> mul_fast(const Fp<P>& a, const Fp<P>& b) { return mul(a, b); }
template <class P> HD Fp<P> sqr_fast(const Fp<P>& a) { return sqr(a); }

// 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + 5 a1^2); 0 -> 0
template <class P>
HD Fp2<P> inv(const Fp2<P>& a) {
    Fp<P> norm = add(sqr(a.c0), mul5(sqr(a.c1)));
    Fp<P> ni = inv(norm);
    Fp2<P> r;
    r.c0 = mul(a.c0, ni);
    r.c1 = neg(mul(a.c1, ni));
    return r;
}

// same value through the binary-Euclid base-field inversion (single-thread tails, see fp.cuh)
template <class P>
HD Fp2<P> inv_euclid(const Fp2<P>& a) {
    Fp<P> norm = add(sqr(a.c0), mul5(sqr(a.c1)));
    Fp<P> ni = inv_euclid(norm);
    Fp2<P> r;
    r.c0 = mul(a.c0, ni);
    r.c1 = neg(mul(a.c1, ni));
    return r;
}
