// 32-bit carry-chain primitives.
//
// On the device each helper is exactly one PTX instruction (mad.lo.cc / madc.hi.cc pairs are
// fused by ptxas into IMAD.WIDE.U32[.X]; add.cc/addc into IADD3[.X]).  Every statement is
// `asm volatile`, which keeps their relative order, so a carry produced by one helper is consumed
// by the next helper that reads it.
//
// On the host (plain g++ or nvcc host pass) the same helpers are emulated with an explicit carry
// flag so the field / curve templates built on top of them can be unit-tested on a CPU-only box
// against the oracle (tests/test_host_emulation.py).  The emulation is test scaffolding: no
// exported entry point of libmpc_cuda.so computes on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
// out-of-line on the device: used for the Fq2 products, whose full inlining into the G2 group law made
// single kernels of >100k instructions and a 25-minute ptxas run
#define HD_NOINLINE __host__ __device__ __noinline__
#else
#define HD inline
#define DEV inline
#define HD_NOINLINE inline
#endif

namespace ptx {

#if defined(__CUDA_ARCH__)

HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// 32x32 -> 64 multiply-accumulate on a 64-bit column with carry: ptxas fuses each
// mul.wide.u32 + add[c].cc.u64 pair into ONE  IMAD.WIDE.U32[.X] Rd, Pcarry, Ra, Rb, Rc[, Pcarry].
HD uint64_t mulw(uint32_t a, uint32_t b) { uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
HD uint64_t madw_cc(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %1, %2; add.cc.u64 %0, %3, t; }" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
HD uint64_t madwc_cc(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %1, %2; addc.cc.u64 %0, %3, t; }" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
HD uint64_t madwc(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %1, %2; addc.u64 %0, %3, t; }" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }

HD uint64_t add64_cc(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HD uint64_t addc64_cc(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HD uint64_t addc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

#else  // host emulation

static thread_local uint32_t g_cc = 0;

HD uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cc; }
// PTX semantics: after the sub family CC.CF holds the borrow-out, and subc consumes it as borrow-in
HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)(t >> 63); return (uint32_t)t; }
HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)(t >> 63); return (uint32_t)t; }
HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cc; }
HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
HD uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }

HD uint64_t mulw(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
HD uint64_t madw_cc(uint32_t a, uint32_t b, uint64_t c) { unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + c; g_cc = (uint32_t)(t >> 64); return (uint64_t)t; }
HD uint64_t madwc_cc(uint32_t a, uint32_t b, uint64_t c) { unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + c + g_cc; g_cc = (uint32_t)(t >> 64); return (uint64_t)t; }
HD uint64_t madwc(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c + g_cc; }
HD uint64_t add64_cc(uint64_t a, uint64_t b) { unsigned __int128 t = (unsigned __int128)a + b; g_cc = (uint32_t)(t >> 64); return (uint64_t)t; }
HD uint64_t addc64_cc(uint64_t a, uint64_t b) { unsigned __int128 t = (unsigned __int128)a + b + g_cc; g_cc = (uint32_t)(t >> 64); return (uint64_t)t; }
HD uint64_t addc64(uint64_t a, uint64_t b) { return a + b + g_cc; }

#endif

}  // namespace ptx
