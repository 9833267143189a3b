/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped GPU path.
 *
 * Short-Weierstrass (a = 0) Jacobian group law and Pippenger MSM template; a CPU
 * restatement of
 *   zero / is_zero        arkworks/algebra/ec/src/models/short_weierstrass_jacobian.rs:487-503
 *   double_in_place a=0   ...short_weierstrass_jacobian.rs:557-600   (dbl-2009-l)
 *   add_assign_mixed      ...short_weierstrass_jacobian.rs:628-693   (madd-2007-bl)
 *   add_assign            ...short_weierstrass_jacobian.rs:721-783   (add-2007-bl)
 *   From<Projective>      ...short_weierstrass_jacobian.rs:823-845   (X/Z^2, Y/Z^3)
 *   mul (bits MSB first)  ...short_weierstrass_jacobian.rs:424-436 / ec/src/lib.rs `mul_bits`
 *   VariableBaseMSM       arkworks/algebra/ec/src/msm/variable_base.rs:12-106
 *   ln_without_floats     arkworks/algebra/ec/src/msm/mod.rs:10-13, ark_std::log2 std/src/lib.rs:67-75
 *
 * Instantiate with  CP(name) symbol prefix, BF(name) base-field prefix (fq / fq2),
 * BF_LIMBS (u64 limbs per base-field element).
 */

typedef struct { BF(t) x, y; uint8_t inf; } CP(aff_t);
typedef struct { BF(t) x, y, z; } CP(jac_t);

static inline void CP(jac_zero)(CP(jac_t) *p) { BF(one)(&p->x); BF(one)(&p->y); BF(zero)(&p->z); }
static inline int CP(jac_is_zero)(const CP(jac_t) *p) { return BF(is_zero)(&p->z); }

static inline void CP(aff_zero)(CP(aff_t) *p) { BF(zero)(&p->x); BF(one)(&p->y); p->inf = 1; }

static void CP(jac_double)(CP(jac_t) *p) {
    if (CP(jac_is_zero)(p)) return;
    BF(t) a, b, c, d, e, f, t;
    BF(sqr)(&a, &p->x);                 /* A = X1^2 */
    BF(sqr)(&b, &p->y);                 /* B = Y1^2 */
    BF(sqr)(&c, &b);                    /* C = B^2 */
    BF(add)(&t, &p->x, &b);             /* D = 2*((X1+B)^2-A-C) */
    BF(sqr)(&t, &t);
    BF(sub)(&t, &t, &a);
    BF(sub)(&t, &t, &c);
    BF(dbl)(&d, &t);
    BF(dbl)(&t, &a);                    /* E = 3*A */
    BF(add)(&e, &a, &t);
    BF(sqr)(&f, &e);                    /* F = E^2 */
    BF(mul)(&p->z, &p->z, &p->y);       /* Z3 = 2*Y1*Z1 */
    BF(dbl)(&p->z, &p->z);
    BF(sub)(&t, &f, &d);                /* X3 = F-2*D */
    BF(sub)(&p->x, &t, &d);
    BF(dbl)(&c, &c); BF(dbl)(&c, &c); BF(dbl)(&c, &c);   /* 8*C */
    BF(sub)(&t, &d, &p->x);             /* Y3 = E*(D-X3)-8*C */
    BF(mul)(&t, &t, &e);
    BF(sub)(&p->y, &t, &c);
}

static void CP(jac_add_mixed)(CP(jac_t) *p, const CP(aff_t) *q) {
    if (q->inf) return;
    if (CP(jac_is_zero)(p)) { p->x = q->x; p->y = q->y; BF(one)(&p->z); return; }
    BF(t) z1z1, u2, s2;
    BF(sqr)(&z1z1, &p->z);
    BF(mul)(&u2, &q->x, &z1z1);
    BF(mul)(&s2, &q->y, &p->z);
    BF(mul)(&s2, &s2, &z1z1);
    if (BF(eq)(&p->x, &u2) && BF(eq)(&p->y, &s2)) { CP(jac_double)(p); return; }
    BF(t) h, hh, i, j, r, v, t;
    BF(sub)(&h, &u2, &p->x);            /* H = U2-X1 */
    BF(sqr)(&hh, &h);                   /* HH = H^2 */
    BF(dbl)(&i, &hh); BF(dbl)(&i, &i);  /* I = 4*HH */
    BF(mul)(&j, &h, &i);                /* J = H*I */
    BF(sub)(&r, &s2, &p->y);            /* r = 2*(S2-Y1) */
    BF(dbl)(&r, &r);
    BF(mul)(&v, &p->x, &i);             /* V = X1*I */
    BF(sqr)(&t, &r);                    /* X3 = r^2-J-2*V */
    BF(sub)(&t, &t, &j);
    BF(sub)(&t, &t, &v);
    BF(sub)(&p->x, &t, &v);
    BF(mul)(&j, &j, &p->y);             /* Y3 = r*(V-X3)-2*Y1*J */
    BF(dbl)(&j, &j);
    BF(sub)(&t, &v, &p->x);
    BF(mul)(&t, &t, &r);
    BF(sub)(&p->y, &t, &j);
    BF(add)(&t, &p->z, &h);             /* Z3 = (Z1+H)^2-Z1Z1-HH */
    BF(sqr)(&t, &t);
    BF(sub)(&t, &t, &z1z1);
    BF(sub)(&p->z, &t, &hh);
}

static void CP(jac_add)(CP(jac_t) *p, const CP(jac_t) *q) {
    if (CP(jac_is_zero)(p)) { *p = *q; return; }
    if (CP(jac_is_zero)(q)) return;
    BF(t) z1z1, z2z2, u1, u2, s1, s2;
    BF(sqr)(&z1z1, &p->z);
    BF(sqr)(&z2z2, &q->z);
    BF(mul)(&u1, &p->x, &z2z2);
    BF(mul)(&u2, &q->x, &z1z1);
    BF(mul)(&s1, &p->y, &q->z);
    BF(mul)(&s1, &s1, &z2z2);
    BF(mul)(&s2, &q->y, &p->z);
    BF(mul)(&s2, &s2, &z1z1);
    if (BF(eq)(&u1, &u2) && BF(eq)(&s1, &s2)) { CP(jac_double)(p); return; }
    BF(t) h, i, j, r, v, t;
    BF(sub)(&h, &u2, &u1);              /* H = U2-U1 */
    BF(dbl)(&i, &h);                    /* I = (2*H)^2 */
    BF(sqr)(&i, &i);
    BF(mul)(&j, &h, &i);                /* J = H*I */
    BF(sub)(&r, &s2, &s1);              /* r = 2*(S2-S1) */
    BF(dbl)(&r, &r);
    BF(mul)(&v, &u1, &i);               /* V = U1*I */
    BF(sqr)(&t, &r);                    /* X3 = r^2-J-2*V */
    BF(sub)(&t, &t, &j);
    BF(t) v2;
    BF(dbl)(&v2, &v);
    BF(sub)(&p->x, &t, &v2);
    BF(sub)(&t, &v, &p->x);             /* Y3 = r*(V-X3)-2*S1*J */
    BF(mul)(&t, &t, &r);
    BF(mul)(&s1, &s1, &j);
    BF(dbl)(&s1, &s1);
    BF(sub)(&p->y, &t, &s1);
    BF(add)(&t, &p->z, &q->z);          /* Z3 = ((Z1+Z2)^2-Z1Z1-Z2Z2)*H */
    BF(sqr)(&t, &t);
    BF(sub)(&t, &t, &z1z1);
    BF(sub)(&t, &t, &z2z2);
    BF(mul)(&p->z, &t, &h);
}

static void CP(jac_to_aff)(CP(aff_t) *out, const CP(jac_t) *p) {
    if (CP(jac_is_zero)(p)) { CP(aff_zero)(out); return; }
    BF(t) one;
    BF(one)(&one);
    out->inf = 0;
    if (BF(eq)(&p->z, &one)) { out->x = p->x; out->y = p->y; return; }
    BF(t) zi, zi2, zi3;
    BF(inv)(&zi, &p->z);
    BF(sqr)(&zi2, &zi);
    BF(mul)(&out->x, &p->x, &zi2);
    BF(mul)(&zi3, &zi2, &zi);
    BF(mul)(&out->y, &p->y, &zi3);
}

static void CP(aff_to_jac)(CP(jac_t) *out, const CP(aff_t) *p) {
    if (p->inf) { CP(jac_zero)(out); return; }
    out->x = p->x; out->y = p->y; BF(one)(&out->z);
}

/* k * P for a canonical little-endian scalar; double-and-add, MSB first, skipping leading zeros */
static void CP(scalar_mul)(CP(jac_t) *out, const CP(aff_t) *p, const uint64_t *k, int klimbs) {
    CP(jac_t) acc;
    CP(jac_zero)(&acc);
    int started = 0;
    for (int i = klimbs * 64 - 1; i >= 0; i--) {
        int bit = (k[i / 64] >> (i % 64)) & 1;
        if (started) CP(jac_double)(&acc);
        if (bit) { CP(jac_add_mixed)(&acc, p); started = 1; }
    }
    *out = acc;
}

static int CP(on_curve)(const CP(aff_t) *p, const BF(t) *coeff_b) {
    if (p->inf) return 1;
    BF(t) lhs, rhs;
    BF(sqr)(&lhs, &p->y);
    BF(sqr)(&rhs, &p->x);
    BF(mul)(&rhs, &rhs, &p->x);
    BF(add)(&rhs, &rhs, coeff_b);
    return BF(eq)(&lhs, &rhs);
}

#ifndef ORC_LOG2_CEIL_DEFINED
#define ORC_LOG2_CEIL_DEFINED
static inline unsigned orc_log2_ceil(size_t x) {       /* ark_std::log2 */
    if (x == 0) return 0;
    if ((x & (x - 1)) == 0) return (unsigned)__builtin_ctzll(x);
    return 64u - (unsigned)__builtin_clzll(x);
}
#endif

/* VariableBaseMSM::multi_scalar_mul over canonical 4-limb (Fr) scalars.
 * `threads` > 1 processes windows concurrently, which is what the reference's
 * cfg_into_iter!(window_starts) does under its (never enabled) `parallel` feature. */
static void CP(msm_bigint)(CP(jac_t) *out, const CP(aff_t) *bases, const uint64_t *scalars /* n*4 */,
                           size_t n, int threads) {
    const unsigned num_bits = 253;      /* FrParameters::MODULUS_BITS */
    unsigned c = n < 32 ? 3 : (orc_log2_ceil(n) * 69 / 100) + 2;
    unsigned nwin = (num_bits + c - 1) / c;
    uint64_t one_repr[4] = {1, 0, 0, 0};
    CP(jac_t) *window_sums = (CP(jac_t) *)malloc(nwin * sizeof(CP(jac_t)));
    size_t nbuckets = ((size_t)1 << c) - 1;
    if (threads < 1) threads = 1;

#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (unsigned w = 0; w < nwin; w++) {
        unsigned w_start = w * c;
        CP(jac_t) res;
        CP(jac_zero)(&res);
        CP(jac_t) *buckets = (CP(jac_t) *)malloc(nbuckets * sizeof(CP(jac_t)));
        for (size_t b = 0; b < nbuckets; b++) CP(jac_zero)(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            const uint64_t *s = scalars + 4 * i;
            if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;               /* filter(!is_zero) */
            if (s[0] == one_repr[0] && (s[1] | s[2] | s[3]) == 0) {
                if (w_start == 0) CP(jac_add_mixed)(&res, &bases[i]);     /* unit scalars once */
                continue;
            }
            /* divn(w_start) then low limb % 2^c */
            unsigned limb = w_start / 64, sh = w_start % 64;
            uint64_t lo = s[limb] >> sh;
            if (sh && limb + 1 < 4) lo |= s[limb + 1] << (64 - sh);
            uint64_t d = lo & (((uint64_t)1 << c) - 1);
            if (d != 0) CP(jac_add_mixed)(&buckets[d - 1], &bases[i]);
        }
        CP(jac_t) running;
        CP(jac_zero)(&running);
        for (size_t b = nbuckets; b-- > 0;) {
            CP(jac_add)(&running, &buckets[b]);
            CP(jac_add)(&res, &running);
        }
        free(buckets);
        window_sums[w] = res;
    }

    CP(jac_t) total;
    CP(jac_zero)(&total);
    for (unsigned w = nwin; w-- > 1;) {
        CP(jac_add)(&total, &window_sums[w]);
        for (unsigned k = 0; k < c; k++) CP(jac_double)(&total);
    }
    /* lowest + total */
    CP(jac_t) lowest = window_sums[0];
    CP(jac_add)(&lowest, &total);
    *out = lowest;
    free(window_sums);
}
