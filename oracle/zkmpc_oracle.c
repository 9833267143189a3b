/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, one thread unless `threads` says otherwise) of the share-MSM /
 * share-NTT / Beaver hot path of Yoii-Inc/zk-mpc, used ONLY as the checker in tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The shipped
 * GPU path (zk-mpc_b200/csrc) never links, loads or calls anything in this directory.
 *
 * The reference is Rust and cannot be built in this image (no cargo/rustc), and its tests hold
 * no recorded input->output vectors for MSM/NTT/Beaver.  This oracle is pinned
 * by (1) the reference's parameter constants and generator KATs (tests/test_oracle_pins.py),
 * (2) an independent Python big-int implementation (tests/pyref.py), (3) algebraic identities.
 * For concrete MSM/NTT/Beaver values the reference offers nothing to pin against:
 * "parity unpinned" in that sense (stated in DESIGN.md).
 *
 * Reference files restated here (paths relative to /root/reference):
 *   constants   arkworks/curves/bls12_377/src/fields/{fr.rs:24-120,fq.rs:3-118,fq2.rs:4-38},
 *               arkworks/curves/bls12_377/src/curves/{g1.rs:9-51,g2.rs:9-86}
 *   Fq2         arkworks/algebra/ff/src/fields/models/quadratic_extension.rs:258-315,632-643
 *   FFT         arkworks/algebra/poly/src/domain/radix2/{mod.rs:51-114,fft.rs:22-307},
 *               arkworks/algebra/poly/src/domain/{mod.rs:92-157,183-190,utils.rs:22-40},
 *               arkworks/algebra/ff/src/fields/mod.rs:363-378 (get_root_of_unity)
 *   Beaver      mpc-algebra/src/share/field.rs:97-129, share/additive.rs:85-152,
 *               share/spdz.rs:31-47,135-146,177-219, wire/field.rs:44-63
 *   MSM glue    arkworks/algebra/ec/src/lib.rs:305-314, mpc-algebra/src/share/msm.rs:33-37,
 *               mpc-algebra/src/share/spdz.rs:482-488
 *   R1CS rows   src/groth16.rs:205-234 (evaluate_constraint)
 *   wire bytes  arkworks/algebra/serialize/src/lib.rs:263-272 ([T]: u64 length + items),
 *               arkworks/algebra/ff/src/fields/macros.rs:1-110 (Fr = 32 LE bytes of into_repr; from_repr on read)
 *   division    arkworks/algebra/poly/src/polynomial/univariate/mod.rs:133-172 (divide_with_q_and_r),
 *               mpc-algebra/src/share/additive.rs:154-162, dense.rs:71-75 (horner_evaluate)
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Fr */
static const uint64_t FR_MOD[4] = {725501752471715841ull, 6461107452199829505ull,
                                   6968279316240510977ull, 1345280370688173398ull};
static const uint64_t FR_R[4] = {9015221291577245683ull, 8239323489949974514ull,
                                 1646089257421115374ull, 958099254763297437ull};
static const uint64_t FR_R2[4] = {2726216793283724667ull, 14712177743343147295ull,
                                  12091039717619697043ull, 81024008013859129ull};
#define FR_INV 725501752471715839ull
static const uint64_t FR_TWO_ADIC_ROOT[4] = {12646347781564978760ull, 6783048705277173164ull,
                                             268534165941069093ull, 1121515446318641358ull};
static const uint64_t FR_GENERATOR[4] = {2984901390528151251ull, 10561528701063790279ull,
                                         5476750214495080041ull, 898978044469942640ull};
#define FR_TWO_ADICITY 47

#define FN 4
#define FP(name) fr_##name
#define F_MOD FR_MOD
#define F_R FR_R
#define F_R2 FR_R2
#define F_INV FR_INV
#include "field_tmpl.h"
#undef FN
#undef FP
#undef F_MOD
#undef F_R
#undef F_R2
#undef F_INV

/* ------------------------------------------------------------------ Fq */
static const uint64_t FQ_MOD[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull,
                                   0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x1ae3a4617c510eaull};
static const uint64_t FQ_R[6] = {202099033278250856ull, 5854854902718660529ull, 11492539364873682930ull,
                                 8885205928937022213ull, 5545221690922665192ull, 39800542322357402ull};
static const uint64_t FQ_R2[6] = {0xb786686c9400cd22ull, 0x329fcaab00431b1ull, 0x22a5f11162d6b46dull,
                                  0xbfdf7d03827dc3acull, 0x837e92f041790bf9ull, 0x6dfccb1e914b88ull};
#define FQ_INV 9586122913090633727ull

#define FN 6
#define FP(name) fq_##name
#define F_MOD FQ_MOD
#define F_R FQ_R
#define F_R2 FQ_R2
#define F_INV FQ_INV
#include "field_tmpl.h"
#undef FN
#undef FP
#undef F_MOD
#undef F_R
#undef F_R2
#undef F_INV

/* ------------------------------------------------------------------ Fq2 = Fq[u]/(u^2+5) */
typedef struct { fq_t c0, c1; } fq2_t;

static inline int fq2_is_zero(const fq2_t *a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static inline int fq2_eq(const fq2_t *a, const fq2_t *b) { return fq_eq(&a->c0, &b->c0) && fq_eq(&a->c1, &b->c1); }
static inline void fq2_zero(fq2_t *r) { fq_zero(&r->c0); fq_zero(&r->c1); }
static inline void fq2_one(fq2_t *r) { fq_one(&r->c0); fq_zero(&r->c1); }
static inline void fq2_add(fq2_t *r, const fq2_t *a, const fq2_t *b) { fq_add(&r->c0, &a->c0, &b->c0); fq_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_sub(fq2_t *r, const fq2_t *a, const fq2_t *b) { fq_sub(&r->c0, &a->c0, &b->c0); fq_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_dbl(fq2_t *r, const fq2_t *a) { fq_dbl(&r->c0, &a->c0); fq_dbl(&r->c1, &a->c1); }
static inline void fq2_neg(fq2_t *r, const fq2_t *a) { fq_neg(&r->c0, &a->c0); fq_neg(&r->c1, &a->c1); }

/* fe * NONRESIDUE with NONRESIDUE = -5 (fq2.rs:29-34: -(2fe) doubled, minus fe) */
static inline void fq_mul_by_nonresidue(fq_t *r, const fq_t *fe) {
    fq_t t;
    fq_dbl(&t, fe);
    fq_neg(&t, &t);
    fq_dbl(&t, &t);
    fq_sub(r, &t, fe);
}

/* Karatsuba (quadratic_extension.rs:632-643) */
static void fq2_mul(fq2_t *r, const fq2_t *a, const fq2_t *b) {
    fq_t v0, v1, s, t, n;
    fq_mul(&v0, &a->c0, &b->c0);
    fq_mul(&v1, &a->c1, &b->c1);
    fq_add(&s, &a->c1, &a->c0);
    fq_add(&t, &b->c0, &b->c1);
    fq_mul(&s, &s, &t);
    fq_sub(&s, &s, &v0);
    fq_sub(&r->c1, &s, &v1);
    fq_mul_by_nonresidue(&n, &v1);
    fq_add(&r->c0, &v0, &n);
}

static inline void fq2_sqr(fq2_t *r, const fq2_t *a) { fq2_mul(r, a, a); }

/* Guide to Pairing-based Cryptography alg. 5.19 (quadratic_extension.rs:297-315) */
static int fq2_inv(fq2_t *r, const fq2_t *a) {
    if (fq2_is_zero(a)) return 0;
    fq_t v0, v1, n;
    fq_sqr(&v1, &a->c1);
    fq_sqr(&v0, &a->c0);
    fq_mul_by_nonresidue(&n, &v1);
    fq_sub(&v0, &v0, &n);
    fq_inv(&v1, &v0);
    fq_mul(&r->c0, &a->c0, &v1);
    fq_mul(&n, &a->c1, &v1);
    fq_neg(&r->c1, &n);
    return 1;
}

/* ------------------------------------------------------------------ curves */
#define CP(name) g1_##name
#define BF(name) fq_##name
#include "curve_tmpl.h"
#undef CP
#undef BF

#define CP(name) g2_##name
#define BF(name) fq2_##name
#include "curve_tmpl.h"
#undef CP
#undef BF

/* generators, canonical decimal values from g1.rs:43-51 / g2.rs:68-86 turned into limbs by
 * tests/pyref.py and cross-checked there; stored canonical, converted to Montgomery on use */
static const uint64_t G1_GEN_X[6] = {0xeab9b16eb21be9efull, 0xd5481512ffcd394eull, 0x188282c8bd37cb5cull,
                                     0x85951e2caa9d41bbull, 0xc8fc6225bf87ff54ull, 0x008848defe740a67ull};
static const uint64_t G1_GEN_Y[6] = {0xfd82de55559c8ea6ull, 0xc2fe3d3634a9591aull, 0x6d182ad44fb82305ull,
                                     0xbd7fb348ca3e52d9ull, 0x1f674f5d30afeec4ull, 0x01914a69c5102effull};

/* ------------------------------------------------------------------ helpers */
static inline uint64_t mix64(uint64_t z) {          /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static void g1_load_aff(g1_aff_t *p, const uint64_t *xy, uint8_t inf) {
    memcpy(p->x.l, xy, 48);
    memcpy(p->y.l, xy + 6, 48);
    p->inf = inf ? 1 : 0;
}
static void g1_store_aff(uint64_t *xy, uint8_t *inf, const g1_aff_t *p) {
    memcpy(xy, p->x.l, 48);
    memcpy(xy + 6, p->y.l, 48);
    *inf = p->inf;
}
static void g2_load_aff(g2_aff_t *p, const uint64_t *xy, uint8_t inf) {
    memcpy(p->x.c0.l, xy, 48);
    memcpy(p->x.c1.l, xy + 6, 48);
    memcpy(p->y.c0.l, xy + 12, 48);
    memcpy(p->y.c1.l, xy + 18, 48);
    p->inf = inf ? 1 : 0;
}
static void g2_store_aff(uint64_t *xy, uint8_t *inf, const g2_aff_t *p) {
    memcpy(xy, p->x.c0.l, 48);
    memcpy(xy + 6, p->x.c1.l, 48);
    memcpy(xy + 12, p->y.c0.l, 48);
    memcpy(xy + 18, p->y.c1.l, 48);
    *inf = p->inf;
}

/* ================================================================== exported: field vectors
 * op: 0 add, 1 sub, 2 mul, 3 neg(a), 4 inv(a) (0 -> 0), 5 from_mont(a), 6 to_mont(a), 7 sqr(a) */
EXPORT void orc_fr_vec(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fr_t x, y, r;
        memcpy(x.l, a + 4 * i, 32);
        if (b) memcpy(y.l, b + 4 * i, 32);
        switch (op) {
            case 0: fr_add(&r, &x, &y); break;
            case 1: fr_sub(&r, &x, &y); break;
            case 2: fr_mul(&r, &x, &y); break;
            case 3: fr_neg(&r, &x); break;
            case 4: if (!fr_inv(&r, &x)) fr_zero(&r); break;
            case 5: fr_from_mont(r.l, &x); break;
            case 6: fr_to_mont(&r, x.l); break;
            default: fr_sqr(&r, &x); break;
        }
        memcpy(out + 4 * i, r.l, 32);
    }
}

EXPORT void orc_fq_vec(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fq_t x, y, r;
        memcpy(x.l, a + 6 * i, 48);
        if (b) memcpy(y.l, b + 6 * i, 48);
        switch (op) {
            case 0: fq_add(&r, &x, &y); break;
            case 1: fq_sub(&r, &x, &y); break;
            case 2: fq_mul(&r, &x, &y); break;
            case 3: fq_neg(&r, &x); break;
            case 4: if (!fq_inv(&r, &x)) fq_zero(&r); break;
            case 5: fq_from_mont(r.l, &x); break;
            case 6: fq_to_mont(&r, x.l); break;
            default: fq_sqr(&r, &x); break;
        }
        memcpy(out + 6 * i, r.l, 48);
    }
}

/* op: 0 add, 1 sub, 2 mul, 3 neg, 4 inv, 7 sqr ; elements are (c0,c1) = 12 limbs */
EXPORT void orc_fq2_vec(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fq2_t x, y, r;
        memcpy(&x, a + 12 * i, 96);
        if (b) memcpy(&y, b + 12 * i, 96);
        switch (op) {
            case 0: fq2_add(&r, &x, &y); break;
            case 1: fq2_sub(&r, &x, &y); break;
            case 2: fq2_mul(&r, &x, &y); break;
            case 3: fq2_neg(&r, &x); break;
            case 4: if (!fq2_inv(&r, &x)) fq2_zero(&r); break;
            default: fq2_sqr(&r, &x); break;
        }
        memcpy(out + 12 * i, &r, 96);
    }
}

EXPORT void orc_fr_batch_inv(uint64_t *v, size_t n) { fr_batch_inv((fr_t *)v, n); }

/* ================================================================== exported: G1 */
EXPORT void orc_g1_generator(uint64_t xy[12]) {
    fq_t x, y;
    fq_to_mont(&x, G1_GEN_X);
    fq_to_mont(&y, G1_GEN_Y);
    memcpy(xy, x.l, 48);
    memcpy(xy + 6, y.l, 48);
}

EXPORT int orc_g1_on_curve(const uint64_t *xy, uint8_t inf) {
    g1_aff_t p;
    fq_t b;
    g1_load_aff(&p, xy, inf);
    fq_one(&b);
    return g1_on_curve(&p, &b);
}

/* out = k * P, k canonical little-endian with `klimbs` limbs; affine out */
EXPORT void orc_g1_scalar_mul(const uint64_t *xy, uint8_t inf, const uint64_t *k, int klimbs,
                              uint64_t *out_xy, uint8_t *out_inf) {
    g1_aff_t p, r;
    g1_jac_t j;
    g1_load_aff(&p, xy, inf);
    g1_scalar_mul(&j, &p, k, klimbs);
    g1_jac_to_aff(&r, &j);
    g1_store_aff(out_xy, out_inf, &r);
}

/* affine + affine through mixed add (exercises add / double / inverse branches) */
EXPORT void orc_g1_add(const uint64_t *a_xy, uint8_t a_inf, const uint64_t *b_xy, uint8_t b_inf,
                       uint64_t *out_xy, uint8_t *out_inf) {
    g1_aff_t a, b, r;
    g1_jac_t j;
    g1_load_aff(&a, a_xy, a_inf);
    g1_load_aff(&b, b_xy, b_inf);
    g1_aff_to_jac(&j, &a);
    g1_jac_add_mixed(&j, &b);
    g1_jac_to_aff(&r, &j);
    g1_store_aff(out_xy, out_inf, &r);
}

/* Jacobian sum of `n` Jacobian points given as (x,y,z) 18-limb triples -> affine.
 * Used to fold per-GPU MSM partials exactly as SURVEY §8(e) describes. */
EXPORT void orc_g1_sum_jac(const uint64_t *pts, size_t n, uint64_t *out_xy, uint8_t *out_inf) {
    g1_jac_t acc, p;
    g1_aff_t r;
    g1_jac_zero(&acc);
    for (size_t i = 0; i < n; i++) {
        memcpy(&p, pts + 18 * i, 144);
        g1_jac_add(&acc, &p);
    }
    g1_jac_to_aff(&r, &acc);
    g1_store_aff(out_xy, out_inf, &r);
}

/* Synthetic bases shared with the GPU generator: P_i = k_i * G,
 * k_i = max(1, mix64(seed + (first+i+1) * 0x9E3779B97F4A7C15)). */
EXPORT void orc_g1_generate(uint64_t seed, size_t first, size_t n, uint64_t *out_xy) {
    g1_aff_t g;
    uint64_t gxy[12];
    orc_g1_generator(gxy);
    g1_load_aff(&g, gxy, 0);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint64_t k = mix64(seed + (uint64_t)(first + i + 1) * 0x9E3779B97F4A7C15ull);
        if (k == 0) k = 1;
        g1_jac_t j;
        g1_aff_t r;
        uint8_t inf;
        g1_scalar_mul(&j, &g, &k, 1);
        g1_jac_to_aff(&r, &j);
        g1_store_aff(out_xy + 12 * i, &inf, &r);
    }
}

/* AffineCurve::multi_scalar_mul + AffineMsm::msm: Montgomery scalars -> into_repr ->
 * VariableBaseMSM -> affine (ec/src/lib.rs:305-314, mpc-algebra/src/share/msm.rs:33-37). */
EXPORT void orc_g1_msm(const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars_mont,
                       size_t n, uint64_t *out_xy, uint8_t *out_inf, int threads) {
    g1_aff_t *bases = (g1_aff_t *)malloc((n ? n : 1) * sizeof(g1_aff_t));
    uint64_t *repr = (uint64_t *)malloc((n ? n : 1) * 32);
    for (size_t i = 0; i < n; i++) {
        g1_load_aff(&bases[i], bases_xy + 12 * i, inf ? inf[i] : 0);
        fr_t s;
        memcpy(s.l, scalars_mont + 4 * i, 32);
        fr_from_mont(repr + 4 * i, &s);
    }
    g1_jac_t j;
    g1_aff_t r;
    g1_msm_bigint(&j, bases, repr, n, threads);
    g1_jac_to_aff(&r, &j);
    g1_store_aff(out_xy, out_inf, &r);
    free(bases);
    free(repr);
}

/* naive sum_i s_i * P_i by double-and-add (test-templates/src/msm.rs:6-14) */
EXPORT void orc_g1_msm_naive(const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars_mont,
                             size_t n, uint64_t *out_xy, uint8_t *out_inf) {
    g1_jac_t acc;
    g1_jac_zero(&acc);
    for (size_t i = 0; i < n; i++) {
        g1_aff_t p;
        g1_jac_t t;
        fr_t s;
        uint64_t k[4];
        g1_load_aff(&p, bases_xy + 12 * i, inf ? inf[i] : 0);
        memcpy(s.l, scalars_mont + 4 * i, 32);
        fr_from_mont(k, &s);
        g1_scalar_mul(&t, &p, k, 4);
        g1_jac_add(&acc, &t);
    }
    g1_aff_t r;
    g1_jac_to_aff(&r, &acc);
    g1_store_aff(out_xy, out_inf, &r);
}

/* ================================================================== exported: G2
 * point layout: x.c0, x.c1, y.c0, y.c1 (24 limbs) */
static void g2_coeff_b(fq2_t *b) {
    /* (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906) */
    static const uint64_t B_C1[6] = {0x8072266666666685ull, 0x8df55926899999a9ull, 0x7fe4561ad64f34cfull,
                                     0xb95da6d8b6e4f01bull, 0x4b747cccfc142743ull, 0x0039c3fa70f49f43ull};
    /* stored as the Montgomery limbs the reference's field_new! produces; checked in tests */
    fq_zero(&b->c0);
    memcpy(b->c1.l, B_C1, 48);
}

EXPORT void orc_g2_coeff_b(uint64_t out[12]) {
    fq2_t b;
    g2_coeff_b(&b);
    memcpy(out, &b, 96);
}

EXPORT int orc_g2_on_curve(const uint64_t *xy, uint8_t inf) {
    g2_aff_t p;
    fq2_t b;
    g2_load_aff(&p, xy, inf);
    g2_coeff_b(&b);
    return g2_on_curve(&p, &b);
}

EXPORT void orc_g2_scalar_mul(const uint64_t *xy, uint8_t inf, const uint64_t *k, int klimbs,
                              uint64_t *out_xy, uint8_t *out_inf) {
    g2_aff_t p, r;
    g2_jac_t j;
    g2_load_aff(&p, xy, inf);
    g2_scalar_mul(&j, &p, k, klimbs);
    g2_jac_to_aff(&r, &j);
    g2_store_aff(out_xy, out_inf, &r);
}

EXPORT void orc_g2_add(const uint64_t *a_xy, uint8_t a_inf, const uint64_t *b_xy, uint8_t b_inf,
                       uint64_t *out_xy, uint8_t *out_inf) {
    g2_aff_t a, b, r;
    g2_jac_t j;
    g2_load_aff(&a, a_xy, a_inf);
    g2_load_aff(&b, b_xy, b_inf);
    g2_aff_to_jac(&j, &a);
    g2_jac_add_mixed(&j, &b);
    g2_jac_to_aff(&r, &j);
    g2_store_aff(out_xy, out_inf, &r);
}

/* P_i = k_i * G2gen with the same k_i schedule as orc_g1_generate; generator passed by caller */
EXPORT void orc_g2_generate(const uint64_t *gen_xy, uint64_t seed, size_t first, size_t n, uint64_t *out_xy) {
    g2_aff_t g;
    g2_load_aff(&g, gen_xy, 0);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint64_t k = mix64(seed + (uint64_t)(first + i + 1) * 0x9E3779B97F4A7C15ull);
        if (k == 0) k = 1;
        g2_jac_t j;
        g2_aff_t r;
        uint8_t inf;
        g2_scalar_mul(&j, &g, &k, 1);
        g2_jac_to_aff(&r, &j);
        g2_store_aff(out_xy + 24 * i, &inf, &r);
    }
}

EXPORT void orc_g2_msm(const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars_mont,
                       size_t n, uint64_t *out_xy, uint8_t *out_inf, int threads) {
    g2_aff_t *bases = (g2_aff_t *)malloc((n ? n : 1) * sizeof(g2_aff_t));
    uint64_t *repr = (uint64_t *)malloc((n ? n : 1) * 32);
    for (size_t i = 0; i < n; i++) {
        g2_load_aff(&bases[i], bases_xy + 24 * i, inf ? inf[i] : 0);
        fr_t s;
        memcpy(s.l, scalars_mont + 4 * i, 32);
        fr_from_mont(repr + 4 * i, &s);
    }
    g2_jac_t j;
    g2_aff_t r;
    g2_msm_bigint(&j, bases, repr, n, threads);
    g2_jac_to_aff(&r, &j);
    g2_store_aff(out_xy, out_inf, &r);
    free(bases);
    free(repr);
}

EXPORT void orc_g2_msm_naive(const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars_mont,
                             size_t n, uint64_t *out_xy, uint8_t *out_inf) {
    g2_jac_t acc;
    g2_jac_zero(&acc);
    for (size_t i = 0; i < n; i++) {
        g2_aff_t p;
        g2_jac_t t;
        fr_t s;
        uint64_t k[4];
        g2_load_aff(&p, bases_xy + 24 * i, inf ? inf[i] : 0);
        memcpy(s.l, scalars_mont + 4 * i, 32);
        fr_from_mont(k, &s);
        g2_scalar_mul(&t, &p, k, 4);
        g2_jac_add(&acc, &t);
    }
    g2_aff_t r;
    g2_jac_to_aff(&r, &acc);
    g2_store_aff(out_xy, out_inf, &r);
}

/* ================================================================== exported: radix-2 domain + FFT */
typedef struct {
    fr_t group_gen, group_gen_inv, size_inv, generator_inv, size_as_fe;
} domain_t;

static int domain_new(domain_t *d, unsigned log_n) {
    if (log_n > FR_TWO_ADICITY) return 0;
    fr_t omega;
    memcpy(omega.l, FR_TWO_ADIC_ROOT, 32);
    for (unsigned i = log_n; i < FR_TWO_ADICITY; i++) fr_sqr(&omega, &omega);
    d->group_gen = omega;
    fr_inv(&d->group_gen_inv, &omega);
    fr_from_u64(&d->size_as_fe, (uint64_t)1 << log_n);
    fr_inv(&d->size_inv, &d->size_as_fe);
    fr_t g;
    memcpy(g.l, FR_GENERATOR, 32);
    fr_inv(&d->generator_inv, &g);
    return 1;
}

/* out: group_gen, group_gen_inv, size_inv, generator_inv, size_as_field_element (5 x 4 limbs) */
EXPORT int orc_domain_params(unsigned log_n, uint64_t *out) {
    domain_t d;
    if (!domain_new(&d, log_n)) return 0;
    memcpy(out, &d, sizeof(d));
    return 1;
}

static fr_t *compute_powers_serial(size_t size, const fr_t *root) {
    fr_t *v = (fr_t *)malloc((size ? size : 1) * sizeof(fr_t));
    fr_t value;
    fr_one(&value);
    for (size_t i = 0; i < size; i++) { v[i] = value; fr_mul(&value, &value, root); }
    return v;
}

#define MIN_NUM_CHUNKS_FOR_COMPACTION 128

static void io_helper(fr_t *xi, size_t n, const fr_t *root) {
    size_t nroots = n / 2;
    fr_t *roots = compute_powers_serial(nroots, root);
    size_t step = 1;
    int first = 1;
    size_t gap = n / 2;
    while (gap > 0) {
        size_t chunk = 2 * gap, num_chunks = n / chunk;
        if (num_chunks >= MIN_NUM_CHUNKS_FOR_COMPACTION) {
            if (!first) {
                size_t s = step * 2, m = 0;
                for (size_t k = 0; k < nroots; k += s) roots[m++] = roots[k];
                nroots = m;
            }
            step = 1;
        } else {
            step = num_chunks;
        }
        first = 0;
        for (size_t c0 = 0; c0 < n; c0 += chunk) {
            fr_t *lo = xi + c0, *hi = xi + c0 + gap;
            for (size_t k = 0; k < gap; k++) {
                fr_t neg;
                fr_sub(&neg, &lo[k], &hi[k]);
                fr_add(&lo[k], &lo[k], &hi[k]);
                fr_mul(&hi[k], &neg, &roots[k * step]);
            }
        }
        gap /= 2;
    }
    free(roots);
}

static void oi_helper(fr_t *xi, size_t n, const fr_t *root) {
    size_t nroots = n / 2;
    fr_t *cache = compute_powers_serial(nroots, root);
    size_t cmax = nroots / 2 < nroots / MIN_NUM_CHUNKS_FOR_COMPACTION ? nroots / 2 : nroots / MIN_NUM_CHUNKS_FOR_COMPACTION;
    fr_t *compacted = (fr_t *)malloc((cmax ? cmax : 1) * sizeof(fr_t));
    size_t gap = 1;
    while (gap < n) {
        size_t chunk = 2 * gap, num_chunks = n / chunk;
        const fr_t *roots;
        size_t step;
        if (num_chunks >= MIN_NUM_CHUNKS_FOR_COMPACTION && gap < n / 2) {
            for (size_t k = 0; k < gap; k++) compacted[k] = cache[k * num_chunks];
            roots = compacted;
            step = 1;
        } else {
            roots = cache;
            step = num_chunks;
        }
        for (size_t c0 = 0; c0 < n; c0 += chunk) {
            fr_t *lo = xi + c0, *hi = xi + c0 + gap;
            for (size_t k = 0; k < gap; k++) {
                fr_t neg;
                fr_mul(&hi[k], &hi[k], &roots[k * step]);
                fr_sub(&neg, &lo[k], &hi[k]);
                fr_add(&lo[k], &lo[k], &hi[k]);
                hi[k] = neg;
            }
        }
        gap *= 2;
    }
    free(cache);
    free(compacted);
}

static void derange(fr_t *xi, unsigned log_len) {
    size_t n = (size_t)1 << log_len;
    if (n < 3) return;
    for (uint64_t idx = 1; idx < n - 1; idx++) {
        uint64_t r = 0, v = idx;
        for (unsigned b = 0; b < log_len; b++) { r = (r << 1) | (v & 1); v >>= 1; }
        if (idx < r) { fr_t t = xi[idx]; xi[idx] = xi[r]; xi[r] = t; }
    }
}

static void distribute_powers_and_mul_by_const(fr_t *x, size_t n, const fr_t *g, const fr_t *c) {
    fr_t pw = *c;
    for (size_t i = 0; i < n; i++) { fr_mul(&x[i], &x[i], &pw); fr_mul(&pw, &pw, g); }
}

/* kind: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft; data = 2^log_n Montgomery Fr elements, in place */
EXPORT int orc_ntt_fr(uint64_t *data, unsigned log_n, unsigned kind) {
    domain_t d;
    if (!domain_new(&d, log_n)) return 1;
    size_t n = (size_t)1 << log_n;
    fr_t *x = (fr_t *)data;
    fr_t one, g;
    fr_one(&one);
    memcpy(g.l, FR_GENERATOR, 32);
    switch (kind) {
        case 2:
            distribute_powers_and_mul_by_const(x, n, &g, &one);
            /* fall through */
        case 0:
            io_helper(x, n, &d.group_gen);
            derange(x, log_n);
            break;
        case 1:
            derange(x, log_n);
            oi_helper(x, n, &d.group_gen_inv);
            for (size_t i = 0; i < n; i++) fr_mul(&x[i], &x[i], &d.size_inv);
            break;
        case 3:
            derange(x, log_n);
            oi_helper(x, n, &d.group_gen_inv);
            distribute_powers_and_mul_by_const(x, n, &d.generator_inv, &d.size_inv);
            break;
        default:
            return 2;
    }
    return 0;
}

/* evals[i] *= (g^n - 1)^-1  (domain/mod.rs:183-190; vanishing poly at the coset generator 22) */
EXPORT void orc_divide_by_vanishing_on_coset(uint64_t *data, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    fr_t g, z, one, zi;
    memcpy(g.l, FR_GENERATOR, 32);
    uint64_t e[1] = {(uint64_t)n};
    fr_pow(&z, &g, e, 1);
    fr_one(&one);
    fr_sub(&z, &z, &one);
    fr_inv(&zi, &z);
    fr_t *x = (fr_t *)data;
    for (size_t i = 0; i < n; i++) fr_mul(&x[i], &x[i], &zi);
}

/* sum_j coeffs[j] * point^j by Horner (for random-point checks of interpolation) */
EXPORT void orc_fr_horner(const uint64_t *coeffs, size_t n, const uint64_t *point, uint64_t *out) {
    fr_t acc, p;
    fr_zero(&acc);
    memcpy(p.l, point, 32);
    const fr_t *c = (const fr_t *)coeffs;
    for (size_t i = n; i-- > 0;) { fr_mul(&acc, &acc, &p); fr_add(&acc, &acc, &c[i]); }
    memcpy(out, acc.l, 32);
}

/* ================================================================== exported: Beaver local halves */
/* out = s + x  (share/field.rs:108-112; SPDZ calls it once per plane: sh, then mac) */
EXPORT void orc_beaver_mask(const uint64_t *s, const uint64_t *x, uint64_t *out, size_t n) {
    orc_fr_vec(0, s, x, out, n);
}

/* out = z - y*sx - x*oy (+ sx*oy on the leader)        additive: share/additive.rs:132-152
 * SPDZ (spdz != 0): inputs are two planes [sh | mac] of n elements each; the mac plane gets
 * + mac_share * sx*oy with mac_share = leader ? 1 : 0   (share/spdz.rs:31-38,197-219) */
EXPORT void orc_beaver_combine(const uint64_t *x, const uint64_t *y, const uint64_t *z,
                               const uint64_t *sx_pub, const uint64_t *oy_pub, uint64_t *out,
                               size_t n, unsigned is_leader, unsigned spdz) {
    unsigned planes = spdz ? 2 : 1;
    for (unsigned p = 0; p < planes; p++) {
        const fr_t *xs = (const fr_t *)x + (size_t)p * n, *ys = (const fr_t *)y + (size_t)p * n;
        const fr_t *zs = (const fr_t *)z + (size_t)p * n;
        fr_t *o = (fr_t *)out + (size_t)p * n;
        for (size_t i = 0; i < n; i++) {
            const fr_t *sx = (const fr_t *)sx_pub + i, *oy = (const fr_t *)oy_pub + i;
            fr_t t, r;
            fr_mul(&t, &ys[i], sx);         /* y.scale(sx) */
            fr_sub(&r, &zs[i], &t);         /* z.sub(..)   */
            fr_mul(&t, &xs[i], oy);         /* x.scale(oy) */
            fr_sub(&r, &r, &t);
            if (is_leader) {                /* shift(sx*oy): leader-only on sh; mac += mac_share * v */
                fr_mul(&t, sx, oy);
                fr_add(&r, &r, &t);
            }
            o[i] = r;
        }
    }
}

/* out[i] = sum over P parties of parts[p][i]   (share/additive.rs:125-131) */
EXPORT void orc_open_sum(const uint64_t *parts, unsigned P, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fr_t acc;
        fr_zero(&acc);
        for (unsigned p = 0; p < P; p++) fr_add(&acc, &acc, (const fr_t *)parts + (size_t)p * n + i);
        memcpy(out + 4 * i, acc.l, 32);
    }
}

/* SPDZ MAC-check local half: dx_t[i] = mac_share * val[i] - mac[i]   (share/spdz.rs:185-189) */
EXPORT void orc_spdz_mac_check(const uint64_t *vals, const uint64_t *macs, uint64_t *out, size_t n,
                               unsigned is_leader) {
    for (size_t i = 0; i < n; i++) {
        fr_t v, r;
        if (is_leader) memcpy(v.l, vals + 4 * i, 32); else fr_zero(&v);
        fr_sub(&r, &v, (const fr_t *)macs + i);
        memcpy(out + 4 * i, r.l, 32);
    }
}

/* generic elementwise ops used by witness_map (src/groth16.rs:298-302) and Marlin rounds:
 * op 0: a - b   1: a * b   2: a * c (c one element)   3: a + c*b (c one element) */
EXPORT void orc_vec_op(unsigned op, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                       uint64_t *out, size_t n) {
    fr_t k;
    if (c) memcpy(k.l, c, 32);
    for (size_t i = 0; i < n; i++) {
        fr_t x, y, r;
        memcpy(x.l, a + 4 * i, 32);
        if (b) memcpy(y.l, b + 4 * i, 32);
        switch (op) {
            case 0: fr_sub(&r, &x, &y); break;
            case 1: fr_mul(&r, &x, &y); break;
            case 2: fr_mul(&r, &x, &k); break;
            default: fr_mul(&r, &k, &y); fr_add(&r, &x, &r); break;
        }
        memcpy(out + 4 * i, r.l, 32);
    }
}

/* ================================================================== exported: next rows */
/* f2 — evaluate_constraint (src/groth16.rs:205-234) for every row of a public CSR matrix against one party's
 * assignment values: sum += coeff.is_one() ? val : val * coeff.  Linear, so on local share values it yields
 * the party's share of the row. */
EXPORT void orc_spmv(const uint64_t *row_ptr, const uint32_t *col, const uint64_t *coeff, size_t rows,
                     const uint64_t *x, uint64_t *out) {
    fr_t one;
    fr_one(&one);
    for (size_t r = 0; r < rows; r++) {
        fr_t sum;
        fr_zero(&sum);
        for (uint64_t k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
            const fr_t *c = (const fr_t *)coeff + k, *v = (const fr_t *)x + col[k];
            if (fr_eq(c, &one)) fr_add(&sum, &sum, v);
            else { fr_t t; fr_mul(&t, v, c); fr_add(&sum, &sum, &t); }
        }
        memcpy(out + 4 * r, sum.l, 32);
    }
}

/* f3 — CanonicalSerialize of a Vec<Fr> as MpcSerNet::broadcast sends it (mpc-algebra/src/channel.rs:12-28):
 * u64 LE length, then 32 LE bytes of the canonical integer per element.  out holds 8 + 32 n bytes. */
EXPORT void orc_fr_vec_serialize(const uint64_t *mont, size_t n, uint8_t *out) {
    uint64_t len = n;
    memcpy(out, &len, 8);              /* x86-64 is little endian, as the wire format */
    for (size_t i = 0; i < n; i++) {
        uint64_t repr[4];
        fr_from_mont(repr, (const fr_t *)mont + i);
        memcpy(out + 8 + 32 * i, repr, 32);
    }
}

/* returns 0 on success, 1 if the length prefix differs from n, 2 + index of the first element >= modulus
 * (arkworks: from_repr fails -> SerializationError::InvalidData) */
EXPORT long orc_fr_vec_deserialize(const uint8_t *in, size_t n, uint64_t *mont_out) {
    uint64_t len;
    memcpy(&len, in, 8);
    if (len != n) return 1;
    for (size_t i = 0; i < n; i++) {
        uint64_t repr[4];
        memcpy(repr, in + 8 + 32 * i, 32);
        if (fr_cmp_raw(repr, FR_MOD) >= 0) return 2 + (long)i;
        fr_to_mont((fr_t *)mont_out + i, repr);
    }
    return 0;
}

/* f4 — DenseOrSparsePolynomial::divide_with_q_and_r (univariate/mod.rs:145-170) on plain coefficient vectors
 * (for shares: the local values, additive.rs:154-162).  q_out has room for n_num elements, r_out for n_num;
 * the lengths after arkworks' truncation of leading zeros are returned through q_len / r_len. */
static size_t poly_trim(const fr_t *c, size_t n) {
    while (n && fr_is_zero(&c[n - 1])) n--;
    return n;
}
EXPORT int orc_poly_div(const uint64_t *num, size_t n_num, const uint64_t *den, size_t n_den, uint64_t *q_out,
                        size_t *q_len, uint64_t *r_out, size_t *r_len) {
    const fr_t *a = (const fr_t *)num, *d = (const fr_t *)den;
    size_t na = poly_trim(a, n_num), nd = poly_trim(d, n_den);
    *q_len = 0; *r_len = 0;
    if (na == 0) return 0;                                   /* zero dividend: (0, 0) */
    if (nd == 0) return -1;                                  /* "Dividing by zero polynomial" */
    if (na < nd) { memcpy(r_out, num, na * 32); *r_len = na; return 0; }
    size_t nq = na - nd + 1;
    fr_t *q = (fr_t *)q_out, *rem = (fr_t *)r_out;
    for (size_t i = 0; i < nq; i++) fr_zero(&q[i]);
    memcpy(rem, a, na * 32);
    size_t nr = na;
    fr_t lead_inv;
    fr_inv(&lead_inv, &d[nd - 1]);
    while (nr != 0 && nr >= nd) {
        fr_t cq;
        fr_mul(&cq, &rem[nr - 1], &lead_inv);
        size_t deg = nr - nd;
        q[deg] = cq;
        for (size_t i = 0; i < nd; i++) {
            fr_t t;
            fr_mul(&t, &cq, &d[i]);
            fr_sub(&rem[deg + i], &rem[deg + i], &t);
        }
        nr = poly_trim(rem, nr);
    }
    *q_len = poly_trim(q, nq);
    *r_len = nr;
    return 0;
}

/* constants for tests */
EXPORT void orc_constants(uint64_t *fr_mod, uint64_t *fr_r, uint64_t *fr_r2, uint64_t *fr_gen, uint64_t *fr_root,
                          uint64_t *fq_mod, uint64_t *fq_r, uint64_t *fq_r2) {
    memcpy(fr_mod, FR_MOD, 32); memcpy(fr_r, FR_R, 32); memcpy(fr_r2, FR_R2, 32);
    memcpy(fr_gen, FR_GENERATOR, 32); memcpy(fr_root, FR_TWO_ADIC_ROOT, 32);
    memcpy(fq_mod, FQ_MOD, 48); memcpy(fq_r, FQ_R, 48); memcpy(fq_r2, FQ_R2, 48);
}
