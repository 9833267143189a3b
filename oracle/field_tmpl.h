/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped GPU path.
 *
 * Prime-field template (Montgomery form, 64-bit little-endian limbs), a CPU
 * restatement of the arithmetic zk-mpc's vendored arkworks fork performs:
 *   mul        arkworks/algebra/ff/src/fields/arithmetic.rs:7-57   (CIOS, no-carry variant)
 *   into_repr  arkworks/algebra/ff/src/fields/arithmetic.rs:59-83  (Montgomery reduction)
 *   from_repr  arkworks/algebra/ff/src/fields/macros.rs:464-474    (x * R2)
 *   add/sub    arkworks/algebra/ff/src/fields/macros.rs:698-717    (add, conditional subtract)
 *   neg/double arkworks/algebra/ff/src/fields/macros.rs:317-323,638-651
 *   inverse    arkworks/algebra/ff/src/fields/macros.rs:389-443    (binary extended Euclid, alg. 16)
 *
 * Instantiate by defining  FN (limb count), FP(name) (symbol prefix macro),
 * F_MOD, F_R, F_R2 (const uint64_t[FN]) and F_INV (uint64_t) and including this file.
 * Every value handed in or out is fully reduced, i.e. in [0, p).
 */

typedef struct { uint64_t l[FN]; } FP(t);

static inline int FP(is_zero)(const FP(t) *a) {
    uint64_t acc = 0;
    for (int i = 0; i < FN; i++) acc |= a->l[i];
    return acc == 0;
}

static inline int FP(eq)(const FP(t) *a, const FP(t) *b) {
    uint64_t acc = 0;
    for (int i = 0; i < FN; i++) acc |= a->l[i] ^ b->l[i];
    return acc == 0;
}

/* lexicographic compare of raw limbs, most significant first: -1, 0, 1 */
static inline int FP(cmp_raw)(const uint64_t *a, const uint64_t *b) {
    for (int i = FN - 1; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}

static inline uint64_t FP(add_raw)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    unsigned __int128 c = 0;
    for (int i = 0; i < FN; i++) {
        c += (unsigned __int128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}

static inline uint64_t FP(sub_raw)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < FN; i++) {
        unsigned __int128 d = (unsigned __int128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

/* conditional final subtraction: macros.rs `reduce()` */
static inline void FP(reduce)(FP(t) *a) {
    if (FP(cmp_raw)(a->l, F_MOD) >= 0) FP(sub_raw)(a->l, a->l, F_MOD);
}

static inline void FP(add)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    /* both BLS12-377 moduli leave a spare top bit, so the raw sum cannot overflow FN limbs */
    FP(add_raw)(r->l, a->l, b->l);
    FP(reduce)(r);
}

static inline void FP(sub)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    if (FP(sub_raw)(r->l, a->l, b->l)) FP(add_raw)(r->l, r->l, F_MOD);
}

static inline void FP(dbl)(FP(t) *r, const FP(t) *a) { FP(add)(r, a, a); }

static inline void FP(neg)(FP(t) *r, const FP(t) *a) {
    if (FP(is_zero)(a)) { *r = *a; return; }
    FP(sub_raw)(r->l, F_MOD, a->l);
}

/* CIOS Montgomery product with the "no-carry" optimisation (valid for both
 * BLS12-377 moduli: top limb has its MSB clear and is not all-ones). */
static inline void FP(mul)(FP(t) *out, const FP(t) *a, const FP(t) *b) {
    uint64_t r[FN];
    for (int i = 0; i < FN; i++) r[i] = 0;
    for (int i = 0; i < FN; i++) {
        unsigned __int128 p = (unsigned __int128)a->l[0] * b->l[i] + r[0];
        uint64_t r0 = (uint64_t)p, carry1 = (uint64_t)(p >> 64);
        uint64_t k = r0 * F_INV;
        unsigned __int128 q = (unsigned __int128)k * F_MOD[0] + r0;
        uint64_t carry2 = (uint64_t)(q >> 64);
        for (int j = 1; j < FN; j++) {
            p = (unsigned __int128)a->l[j] * b->l[i] + r[j] + carry1;
            carry1 = (uint64_t)(p >> 64);
            q = (unsigned __int128)k * F_MOD[j] + (uint64_t)p + carry2;
            r[j - 1] = (uint64_t)q;
            carry2 = (uint64_t)(q >> 64);
        }
        r[FN - 1] = carry1 + carry2;
    }
    for (int i = 0; i < FN; i++) out->l[i] = r[i];
    FP(reduce)(out);
}

static inline void FP(sqr)(FP(t) *out, const FP(t) *a) { FP(mul)(out, a, a); }

/* Montgomery form -> canonical integer (arithmetic.rs:59-83) */
static inline void FP(from_mont)(uint64_t *out, const FP(t) *a) {
    uint64_t r[FN];
    for (int i = 0; i < FN; i++) r[i] = a->l[i];
    for (int i = 0; i < FN; i++) {
        uint64_t k = r[i] * F_INV;
        unsigned __int128 c = (unsigned __int128)k * F_MOD[0] + r[i];
        uint64_t carry = (uint64_t)(c >> 64);
        for (int j = 1; j < FN; j++) {
            int idx = (j + i) % FN;
            c = (unsigned __int128)k * F_MOD[j] + r[idx] + carry;
            r[idx] = (uint64_t)c;
            carry = (uint64_t)(c >> 64);
        }
        r[i % FN] = carry;
    }
    for (int i = 0; i < FN; i++) out[i] = r[i];
}

/* canonical integer (< p) -> Montgomery form (macros.rs:464-474) */
static inline void FP(to_mont)(FP(t) *out, const uint64_t *in) {
    FP(t) x, r2;
    for (int i = 0; i < FN; i++) { x.l[i] = in[i]; r2.l[i] = F_R2[i]; }
    FP(mul)(out, &x, &r2);
}

static inline void FP(one)(FP(t) *r) { for (int i = 0; i < FN; i++) r->l[i] = F_R[i]; }
static inline void FP(zero)(FP(t) *r) { for (int i = 0; i < FN; i++) r->l[i] = 0; }

static inline void FP(from_u64)(FP(t) *r, uint64_t v) {
    uint64_t c[FN];
    for (int i = 0; i < FN; i++) c[i] = 0;
    c[0] = v;
    FP(to_mont)(r, c);
}

static inline int FP(raw_is_one)(const uint64_t *a) {
    uint64_t acc = a[0] ^ 1;
    for (int i = 1; i < FN; i++) acc |= a[i];
    return acc == 0;
}

static inline void FP(raw_div2)(uint64_t *a) {
    for (int i = 0; i < FN - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[FN - 1] >>= 1;
}

/* Binary extended Euclid; `a` must be non-zero.  Returns 0 when a == 0. */
static int FP(inv)(FP(t) *out, const FP(t) *a) {
    if (FP(is_zero)(a)) return 0;
    uint64_t u[FN], v[FN];
    FP(t) b, c;
    for (int i = 0; i < FN; i++) { u[i] = a->l[i]; v[i] = F_MOD[i]; b.l[i] = F_R2[i]; c.l[i] = 0; }
    while (!FP(raw_is_one)(u) && !FP(raw_is_one)(v)) {
        while ((u[0] & 1) == 0) {
            FP(raw_div2)(u);
            if (b.l[0] & 1) FP(add_raw)(b.l, b.l, F_MOD);
            FP(raw_div2)(b.l);
        }
        while ((v[0] & 1) == 0) {
            FP(raw_div2)(v);
            if (c.l[0] & 1) FP(add_raw)(c.l, c.l, F_MOD);
            FP(raw_div2)(c.l);
        }
        if (FP(cmp_raw)(v, u) < 0) {
            FP(sub_raw)(u, u, v);
            FP(sub)(&b, &b, &c);
        } else {
            FP(sub_raw)(v, v, u);
            FP(sub)(&c, &c, &b);
        }
    }
    *out = FP(raw_is_one)(u) ? b : c;
    return 1;
}

/* a^e for a little-endian multi-limb exponent (square-and-multiply, MSB first) */
static void FP(pow)(FP(t) *out, const FP(t) *a, const uint64_t *e, int elimbs) {
    FP(t) acc;
    FP(one)(&acc);
    int started = 0;
    for (int i = elimbs * 64 - 1; i >= 0; i--) {
        int bit = (e[i / 64] >> (i % 64)) & 1;
        if (started) FP(sqr)(&acc, &acc);
        if (bit) { FP(mul)(&acc, &acc, a); started = 1; }
    }
    *out = acc;
}

/* Montgomery's trick over non-zero entries (fields/mod.rs:624-660) */
static void FP(batch_inv)(FP(t) *v, size_t n) {
    FP(t) *prod = (FP(t) *)malloc((n ? n : 1) * sizeof(FP(t)));
    FP(t) tmp;
    FP(one)(&tmp);
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        if (FP(is_zero)(&v[i])) continue;
        FP(mul)(&tmp, &tmp, &v[i]);
        prod[m++] = tmp;
    }
    if (m) {
        FP(inv)(&tmp, &tmp);
        size_t k = m;
        for (size_t i = n; i-- > 0;) {
            if (FP(is_zero)(&v[i])) continue;
            k--;
            FP(t) s, new_tmp;
            if (k == 0) FP(one)(&s); else s = prod[k - 1];
            FP(mul)(&new_tmp, &tmp, &v[i]);
            FP(mul)(&v[i], &tmp, &s);
            tmp = new_tmp;
        }
    }
    free(prod);
}
