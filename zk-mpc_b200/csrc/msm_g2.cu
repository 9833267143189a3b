// C ABI of the share MSM over G2 (b_g2_query, src/groth16.rs:160): msm_g1.cu's entry points with F = Fq2.
// A separate translation unit so the two curves compile in parallel.
#define MSM_CURVE_G2 1
#include "msm_g1.cu"
