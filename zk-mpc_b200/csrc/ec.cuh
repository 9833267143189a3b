// Short-Weierstrass (a = 0) group law for BLS12-377 G1 (F = Fq) and G2 (F = Fq2) on the device.
//
// The reference accumulates in Jacobian coordinates
//   add_assign_mixed  arkworks/algebra/ec/src/models/short_weierstrass_jacobian.rs:628-693
//   add_assign        ...:721-783        double_in_place ...:557-600        into affine ...:823-845
// and only the affine normal form of a sum is observable (mpc-algebra/src/share/msm.rs:33-37 converts
// immediately), so the device is free to use extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ = 0), whose mixed addition needs 8M + 2S instead
// of 7M + 4S and whose accumulator start is a plain copy.  Formulas: EFD shortw/xyzz madd-2008-s,
// add-2008-s, dbl-2008-s-1 (a = 0).  All exceptional cases (infinity, P + P, P - P) are handled, the
// result is the exact group element, hence bit-identical to the CPU path after normalisation.
#pragma once
#include "fp.cuh"
#include "fp2.cuh"

template <class F>
struct Affine {
    F x, y;
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    HD static XYZZ infinity() {
        XYZZ r;
        r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    HD bool is_inf() const { return zz.is_zero(); }
};

template <class F>
struct Jac {
    F x, y, z;      // x = X/Z^2, y = Y/Z^3; infinity <=> Z = 0 (the reference's zero() is (1,1,0))
};

// p = 2 * (qx, qy) for an affine point (never infinity; y != 0 on these prime-order-cofactor curves'
// points we see, and y = 0 would correctly yield ZZ = 0 = infinity)
template <class F>
HD void xyzz_mdbl(XYZZ<F>& p, const F& qx, const F& qy) {
    F u = dbl(qy);
    F v = sqr(u);
    F w = mul(u, v);
    F s = mul(qx, v);
    F xx = sqr(qx);
    F m = add(dbl(xx), xx);
    F x3 = sub(sqr(m), dbl(s));
    p.y = sub(mul(m, sub(s, x3)), mul(w, qy));
    p.x = x3;
    p.zz = v;
    p.zzz = w;
}

template <class F>
HD void xyzz_dbl(XYZZ<F>& p) {
    if (p.is_inf()) return;
    F u = dbl(p.y);
    F v = sqr(u);
    F w = mul(u, v);
    F s = mul(p.x, v);
    F xx = sqr(p.x);
    F m = add(dbl(xx), xx);
    F x3 = sub(sqr(m), dbl(s));
    F y3 = sub(mul(m, sub(s, x3)), mul(w, p.y));
    p.x = x3;
    p.y = y3;
    p.zz = mul(v, p.zz);
    p.zzz = mul(w, p.zzz);
}

// The same addition with every product inlined (mul_fast): the hot loop of the bucket accumulation.  For G2 the
// out-of-line Fq2 products of xyzz_madd pass their operands through local memory on every call.
template <class F>
HD void xyzz_madd_fast(XYZZ<F>& p, const F& qx, const F& qy) {
    if (p.is_inf()) {
        p.x = qx; p.y = qy; p.zz = F::one(); p.zzz = F::one();
        return;
    }
    F pp = sub(mul_fast(qx, p.zz), p.x);
    F r = sub(mul_fast(qy, p.zzz), p.y);
    if (pp.is_zero()) {
        if (r.is_zero()) xyzz_mdbl(p, qx, qy);      // rare: the out-of-line tangent
        else p = XYZZ<F>::infinity();
        return;
    }
    F p2 = sqr_fast(pp);
    F p3 = mul_fast(pp, p2);
    F q = mul_fast(p.x, p2);
    F x3 = sub(sub(sqr_fast(r), p3), dbl(q));
    p.y = sub(mul_fast(r, sub(q, x3)), mul_fast(p.y, p3));
    p.x = x3;
    p.zz = mul_fast(p.zz, p2);
    p.zzz = mul_fast(p.zzz, p3);
}

// p += (qx, qy), the affine point being finite
template <class F>
HD void xyzz_madd(XYZZ<F>& p, const F& qx, const F& qy) {
    if (p.is_inf()) {
        p.x = qx; p.y = qy; p.zz = F::one(); p.zzz = F::one();
        return;
    }
    F pp = sub(mul(qx, p.zz), p.x);       // P = U2 - X1
    F r = sub(mul(qy, p.zzz), p.y);       // R = S2 - Y1
    if (pp.is_zero()) {
        if (r.is_zero()) xyzz_mdbl(p, qx, qy);
        else p = XYZZ<F>::infinity();
        return;
    }
    F p2 = sqr(pp);
    F p3 = mul(pp, p2);
    F q = mul(p.x, p2);
    F x3 = sub(sub(sqr(r), p3), dbl(q));
    p.y = sub(mul(r, sub(q, x3)), mul(p.y, p3));
    p.x = x3;
    p.zz = mul(p.zz, p2);
    p.zzz = mul(p.zzz, p3);
}

// p += q (both XYZZ)
template <class F>
HD void xyzz_add(XYZZ<F>& p, const XYZZ<F>& q) {
    if (q.is_inf()) return;
    if (p.is_inf()) { p = q; return; }
    F u1 = mul(p.x, q.zz);
    F u2 = mul(q.x, p.zz);
    F s1 = mul(p.y, q.zzz);
    F s2 = mul(q.y, p.zzz);
    F pp = sub(u2, u1);
    F r = sub(s2, s1);
    if (pp.is_zero()) {
        if (r.is_zero()) xyzz_dbl(p);
        else p = XYZZ<F>::infinity();
        return;
    }
    F p2 = sqr(pp);
    F p3 = mul(pp, p2);
    F q2 = mul(u1, p2);
    F x3 = sub(sub(sqr(r), p3), dbl(q2));
    p.y = sub(mul(r, sub(q2, x3)), mul(s1, p3));
    p.x = x3;
    p.zz = mul(mul(p.zz, q.zz), p2);
    p.zzz = mul(mul(p.zzz, q.zzz), p3);
}

// XYZZ -> Jacobian with Z = ZZ: (X*ZZ, Y*ZZZ, ZZ)   [x = X*ZZ/ZZ^2, y = Y*ZZZ/ZZ^3 = Y/ZZZ]
template <class F>
HD Jac<F> xyzz_to_jac(const XYZZ<F>& p) {
    Jac<F> j;
    if (p.is_inf()) { j.x = F::one(); j.y = F::one(); j.z = F::zero(); return j; }
    j.x = mul(p.x, p.zz);
    j.y = mul(p.y, p.zzz);
    j.z = p.zz;
    return j;
}

// Jacobian -> XYZZ: (X, Y, Z^2, Z^3)
template <class F>
HD XYZZ<F> jac_to_xyzz(const Jac<F>& j) {
    XYZZ<F> p;
    if (j.z.is_zero()) return XYZZ<F>::infinity();
    p.x = j.x; p.y = j.y;
    p.zz = sqr(j.z);
    p.zzz = mul(p.zz, j.z);
    return p;
}

// affine normal form; returns false (and the reference's affine zero (0, 1)) for infinity
// (short_weierstrass_jacobian.rs:167-169,823-845)
template <class F>
HD bool xyzz_to_affine(const XYZZ<F>& p, F& ox, F& oy) {
    if (p.is_inf()) { ox = F::zero(); oy = F::one(); return false; }
    // one inversion: 1/(ZZ*ZZZ) -> 1/ZZ = that * ZZZ, 1/ZZZ = that * ZZ
    F t = inv(mul(p.zz, p.zzz));
    ox = mul(p.x, mul(t, p.zzz));
    oy = mul(p.y, mul(t, p.zz));
    return true;
}

// p = k * (qx, qy) for a small unsigned multiplier (double-and-add, MSB first)
template <class F>
HD XYZZ<F> xyzz_mul_small(const XYZZ<F>& q, uint64_t k) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        if (started) xyzz_dbl(acc);
        if ((k >> i) & 1) { xyzz_add(acc, q); started = true; }
    }
    return acc;
}

// affine normal form through the binary-Euclid inversion: for the one-thread result emission
template <class F>
HD bool xyzz_to_affine_tail(const XYZZ<F>& p, F& ox, F& oy) {
    if (p.is_inf()) { ox = F::zero(); oy = F::one(); return false; }
    F t = inv_euclid(mul(p.zz, p.zzz));
    ox = mul(p.x, mul(t, p.zzz));
    oy = mul(p.y, mul(t, p.zz));
    return true;
}

// Jacobian doubling for a = 0 (EFD dbl-2009-l, the formula of double_in_place,
// short_weierstrass_jacobian.rs:557-600): 2M + 5S against XYZZ's 6M + 3S.  Used by the window-table
// builder, whose cost is the 253 doublings per base.  Z = 0 stays Z = 0.
template <class F>
HD void jac_dbl(Jac<F>& p) {
    F a = sqr(p.x);
    F b = sqr(p.y);
    F c = sqr(b);
    F d = dbl(sub(sub(sqr(add(p.x, b)), a), c));
    F e = add(dbl(a), a);
    F f = sqr(e);
    F z3 = dbl(mul(p.y, p.z));
    F x3 = sub(f, dbl(d));
    F c8 = dbl(dbl(dbl(c)));
    p.y = sub(mul(e, sub(d, x3)), c8);
    p.x = x3;
    p.z = z3;
}

#if defined(__CUDACC__)
// ---- warp-cooperative group law for the serial tails ----------------------------------------------------
// The Horner combine of the window sums and the last few additions of an MSM are one dependent chain of
// field products; a single thread pays the full carry-chain latency of every product (~1 us each).  Here
// all lanes of a warp hold the same point and the independent products of one formula level are computed
// by different lanes at once (SIMD lanes are free), then broadcast: a doubling is 3 product latencies
// instead of 9, an addition 5 instead of 14.  Callers run these with a full, converged warp.
template <class P>
DEV Fp<P> warp_bcast(const Fp<P>& a, int src) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
    return r;
}
template <class P>
DEV Fp2<P> warp_bcast(const Fp2<P>& a, int src) {
    Fp2<P> r;
    r.c0 = warp_bcast(a.c0, src);
    r.c1 = warp_bcast(a.c1, src);
    return r;
}

template <class P>
DEV Fp<P> warp_shfl_up(const Fp<P>& a, int delta) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], delta);
    return r;
}
template <class P>
DEV Fp2<P> warp_shfl_up(const Fp2<P>& a, int delta) {
    Fp2<P> r;
    r.c0 = warp_shfl_up(a.c0, delta);
    r.c1 = warp_shfl_up(a.c1, delta);
    return r;
}
template <class P>
DEV Fp<P> warp_shfl_down(const Fp<P>& a, int delta) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
    return r;
}
template <class P>
DEV Fp2<P> warp_shfl_down(const Fp2<P>& a, int delta) {
    Fp2<P> r;
    r.c0 = warp_shfl_down(a.c0, delta);
    r.c1 = warp_shfl_down(a.c1, delta);
    return r;
}

// r[j] = a[j] * b[j] for j < K (K <= 4), product j computed by lane j, every lane receives all results
template <int K, class F>
DEV void warp_mul(F* r, const F* a, const F* b) {
    const int lane = threadIdx.x & 31;
    F x = a[0], y = b[0];
#pragma unroll
    for (int j = 1; j < K; j++)
        if (lane == j) { x = a[j]; y = b[j]; }
    F p = mul(x, y);
#pragma unroll
    for (int j = 0; j < K; j++) r[j] = warp_bcast(p, j);
}

template <class F>
DEV void xyzz_dbl_warp(XYZZ<F>& p) {
    if (p.is_inf()) return;                  // uniform across the warp: every lane holds the same point
    F u = dbl(p.y);
    F a1[2] = {u, p.x}, r1[2];
    warp_mul<2>(r1, a1, a1);                 // v = u^2, xx = x^2
    F v = r1[0], m = add(dbl(r1[1]), r1[1]);
    F a2[4] = {u, p.x, m, v}, b2[4] = {v, v, m, p.zz}, r2[4];
    warp_mul<4>(r2, a2, b2);                 // w = u v, s = x v, m^2, zz' = v zz
    F w = r2[0], s = r2[1];
    F x3 = sub(r2[2], dbl(s));
    F a3[3] = {m, w, w}, b3[3] = {sub(s, x3), p.y, p.zzz}, r3[3];
    warp_mul<3>(r3, a3, b3);                 // m (s - x3), w y, zzz' = w zzz
    p.x = x3;
    p.y = sub(r3[0], r3[1]);
    p.zz = r2[3];
    p.zzz = r3[2];
}

template <class F>
DEV void xyzz_add_warp(XYZZ<F>& p, const XYZZ<F>& q) {
    if (q.is_inf()) return;
    if (p.is_inf()) { p = q; return; }
    F a1[4] = {p.x, q.x, p.y, q.y}, b1[4] = {q.zz, p.zz, q.zzz, p.zzz}, r1[4];
    warp_mul<4>(r1, a1, b1);                 // u1, u2, s1, s2
    F pp = sub(r1[1], r1[0]);
    F r = sub(r1[3], r1[2]);
    if (pp.is_zero()) {
        if (r.is_zero()) xyzz_dbl_warp(p);
        else p = XYZZ<F>::infinity();
        return;
    }
    F a2[4] = {pp, r, p.zz, p.zzz}, b2[4] = {pp, r, q.zz, q.zzz}, r2[4];
    warp_mul<4>(r2, a2, b2);                 // p2, r^2, zz1 zz2, zzz1 zzz2
    F p2 = r2[0];
    F a3[3] = {pp, r1[0], r2[2]}, b3[3] = {p2, p2, p2}, r3[3];
    warp_mul<3>(r3, a3, b3);                 // p3, q2 = u1 p2, zz3
    F p3 = r3[0], q2 = r3[1];
    F x3 = sub(sub(r2[1], p3), dbl(q2));
    F a4[3] = {r, r1[2], r2[3]}, b4[3] = {sub(q2, x3), p3, p3}, r4[3];
    warp_mul<3>(r4, a4, b4);                 // r (q2 - x3), s1 p3, zzz3
    p.x = x3;
    p.y = sub(r4[0], r4[1]);
    p.zz = r3[2];
    p.zzz = r4[2];
}

// k * q for a small unsigned multiplier, warp-cooperative (double-and-add, MSB first)
template <class F>
DEV XYZZ<F> xyzz_mul_small_warp(const XYZZ<F>& q, uint64_t k) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        if (started) xyzz_dbl_warp(acc);
        if ((k >> i) & 1) { xyzz_add_warp(acc, q); started = true; }
    }
    return acc;
}
#endif
