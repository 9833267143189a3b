/*
 * C consumer of the drop-in boundary: includes include/mpc_cuda.h from plain C, links libmpc_cuda.so and
 * (as the checker) the CPU oracle, and drives MSM, NTT and the Beaver kernels through the real C ABI —
 * what the Rust `mpc-cuda` crate's `extern "C"` block binds.  Built and run by tests/test_c_abi.py.
 * Exit code 0 = every comparison was bit-exact.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpc_cuda.h"

/* oracle entry points (oracle/zkmpc_oracle.c) */
void orc_g1_generate(uint64_t seed, size_t first, size_t n, uint64_t *out_xy);
void orc_g1_msm(const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars_mont, size_t n, uint64_t *out_xy,
                uint8_t *out_inf, int threads);
int orc_ntt_fr(uint64_t *data, unsigned log_n, unsigned kind);
void orc_beaver_combine(const uint64_t *x, const uint64_t *y, const uint64_t *z, const uint64_t *sx_pub,
                        const uint64_t *oy_pub, uint64_t *out, size_t n, unsigned is_leader, unsigned spdz);
void orc_fr_vec_serialize(const uint64_t *mont, size_t n, uint8_t *out);

static const uint64_t FR_MOD[4] = {725501752471715841ull, 6461107452199829505ull, 6968279316240510977ull,
                                   1345280370688173398ull};

static uint64_t mix(uint64_t *state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* uniform below r, used directly as Montgomery limbs (the reference sampler's convention) */
static void fr_rand(uint64_t *out, size_t n, uint64_t seed) {
    for (size_t i = 0; i < n; i++) {
        for (;;) {
            uint64_t *e = out + 4 * i;
            for (int k = 0; k < 4; k++) e[k] = mix(&seed);
            e[3] &= 0xFFFFFFFFFFFFFFFFull >> 3;
            int lt = 0;
            for (int k = 3; k >= 0; k--) {
                if (e[k] != FR_MOD[k]) { lt = e[k] < FR_MOD[k]; break; }
            }
            if (lt) break;
        }
    }
}

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int32_t rc_ = (call);                                                             \
        if (rc_ != MPC_CUDA_OK) {                                                         \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mpc_cuda_last_error());         \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

int main(void) {
    printf("%s\n", mpc_cuda_version());
    CHECK(mpc_cuda_init(NULL, 0));
    CHECK(mpc_cuda_set_party(0, 3));
    int failures = 0;

    /* share MSM through the host-buffer call and a registered handle */
    size_t n = 3000;
    uint64_t *bases = malloc(n * 96), *scalars = malloc(n * 32);
    uint8_t *inf = calloc(n, 1);
    orc_g1_generate(0xABC, 0, n, bases);
    fr_rand(scalars, n, 1);
    inf[17] = 1;
    uint64_t got[12], exp[12];
    uint8_t gi = 0, ei = 0;
    CHECK(mpc_cuda_msm_g1(bases, inf, scalars, n, got, &gi));
    orc_g1_msm(bases, inf, scalars, n, exp, &ei, 4);
    if (memcmp(got, exp, sizeof(got)) || gi != ei) { fprintf(stderr, "msm_g1 mismatch\n"); failures++; }
    uint64_t handle = 0;
    CHECK(mpc_cuda_msm_g1_register_bases(bases, inf, n, &handle));
    CHECK(mpc_cuda_msm_g1_precompute(handle, 0));
    CHECK(mpc_cuda_msm_g1_handle(handle, 0, scalars, n, got, &gi));
    if (memcmp(got, exp, sizeof(got)) || gi != ei) { fprintf(stderr, "msm_g1_handle mismatch\n"); failures++; }
    CHECK(mpc_cuda_msm_release_bases(handle));

    /* share NTT, every kind */
    unsigned log_n = 12;
    size_t m = (size_t)1 << log_n;
    uint64_t *v = malloc(m * 32), *w = malloc(m * 32);
    for (unsigned kind = 0; kind < 4; kind++) {
        fr_rand(v, m, 10 + kind);
        memcpy(w, v, m * 32);
        CHECK(mpc_cuda_ntt_fr(v, log_n, kind, 1));
        if (orc_ntt_fr(w, log_n, kind)) return 3;
        if (memcmp(v, w, m * 32)) { fprintf(stderr, "ntt kind %u mismatch\n", kind); failures++; }
    }

    /* Beaver combine (leader, additive) and the wire bytes of a masked vector */
    uint64_t *x = malloc(m * 32), *y = malloc(m * 32), *z = malloc(m * 32), *sx = malloc(m * 32), *oy = malloc(m * 32);
    fr_rand(x, m, 21); fr_rand(y, m, 22); fr_rand(z, m, 23); fr_rand(sx, m, 24); fr_rand(oy, m, 25);
    CHECK(mpc_cuda_beaver_combine(x, y, z, sx, oy, v, m, 1, 0));
    orc_beaver_combine(x, y, z, sx, oy, w, m, 1, 0);
    if (memcmp(v, w, m * 32)) { fprintf(stderr, "beaver_combine mismatch\n"); failures++; }
    uint8_t *b1 = malloc(8 + 32 * m), *b2 = malloc(8 + 32 * m);
    CHECK(mpc_cuda_fr_serialize(x, m, b1));
    orc_fr_vec_serialize(x, m, b2);
    if (memcmp(b1, b2, 8 + 32 * m)) { fprintf(stderr, "fr_serialize mismatch\n"); failures++; }

    /* division by the vanishing polynomial x^64 - 1 from page-locked buffers: q (x^64 - 1) + r must give p back */
    {
        const size_t dn = 1000, dm = 64;
        uint64_t *p = NULL, *q = NULL, *r = NULL, *back = NULL;
        CHECK(mpc_cuda_host_alloc((void**)&p, dn * 32));
        CHECK(mpc_cuda_host_alloc((void**)&q, (dn - dm) * 32));
        CHECK(mpc_cuda_host_alloc((void**)&r, dm * 32));
        CHECK(mpc_cuda_host_alloc((void**)&back, dn * 32));
        fr_rand(p, dn, 31);
        CHECK(mpc_cuda_poly_div_vanishing(p, dn, dm, q, r));
        /* back = q (x^m - 1); p - back (in place, out = a) must equal r on the low m coefficients and 0 above */
        CHECK(mpc_cuda_poly_mul_vanishing(q, dn - dm, dm, back));
        CHECK(mpc_cuda_vec_op(MPC_CUDA_VEC_SUB, p, back, NULL, p, dn));
        if (memcmp(p, r, dm * 32)) { fprintf(stderr, "poly_div_vanishing: remainder mismatch\n"); failures++; }
        for (size_t i = dm * 4; i < dn * 4; i++)
            if (p[i]) { fprintf(stderr, "poly_div_vanishing: high coefficients differ\n"); failures++; break; }
        CHECK(mpc_cuda_host_free(p)); CHECK(mpc_cuda_host_free(q)); CHECK(mpc_cuda_host_free(r)); CHECK(mpc_cuda_host_free(back));
    }

    /* error convention: non-zero status + message, no abort */
    if (mpc_cuda_set_option("no_such_option", 1) == MPC_CUDA_OK) { fprintf(stderr, "bad option accepted\n"); failures++; }
    if (mpc_cuda_msm_g1_handle(12345678, 0, scalars, 1, got, &gi) != MPC_CUDA_ERR_HANDLE) {
        fprintf(stderr, "unknown handle not reported\n");
        failures++;
    }
    printf("abi_check: %d failure(s), %llu kernel launches\n", failures, (unsigned long long)mpc_cuda_launch_count());
    return failures ? 1 : 0;
}
