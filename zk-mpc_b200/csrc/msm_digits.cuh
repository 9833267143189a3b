// Window plan and signed-digit recoding for the Pippenger bucket MSM (host + device).
//
// The reference (arkworks/algebra/ec/src/msm/variable_base.rs:21-31,50-64) cuts the canonical
// scalar into unsigned c-bit windows and keeps 2^c - 1 buckets per window.  Because only the group
// element Σ sᵢPᵢ is observable, the device recodes each scalar into signed digits
// dₖ ∈ [-(2^(c-1) - 1), 2^(c-1)] with Σ dₖ 2^(ck) = s, which halves the buckets per window
// (negative digits add -P, a free y-negation).  The last window is never folded, and the plan
// guarantees its value (≤ 2^(253 mod c) - 1 + carry) fits the bucket range.
#pragma once
#include <stdint.h>
#include "ptx.cuh"

namespace msm {

constexpr uint32_t SCALAR_BITS = 253;       // FrParameters::MODULUS_BITS (bls12_377/src/fields/fr.rs:28)
constexpr uint32_t DIGIT_NEG = 0x80000000u;

// number of windows for window width c: floor(253/c) + 1, so the top window holds 253 mod c < c bits
HD uint32_t num_windows(uint32_t c) { return SCALAR_BITS / c + 1; }

// Encoded signed digit of window w: 0 when the digit is zero, else magnitude (1 .. 2^(c-1)) with
// DIGIT_NEG set for negative digits.  `carry` threads through the windows (start at 0, w ascending).
HD uint32_t signed_digit(const uint32_t* s /* 8 canonical limbs */, uint32_t w, uint32_t c, uint32_t nwin,
                         uint32_t& carry) {
    uint32_t bit = w * c;
    uint32_t raw = 0;
    if (bit < 256) {
        uint32_t limb = bit >> 5, sh = bit & 31;
        uint64_t lo = s[limb];
        if (limb + 1 < 8) lo |= (uint64_t)s[limb + 1] << 32;
        raw = (uint32_t)(lo >> sh) & ((1u << c) - 1u);
    }
    uint32_t d = raw + carry;
    if (w + 1 < nwin && d > (1u << (c - 1))) {
        carry = 1;
        uint32_t mag = (1u << c) - d;           // d == 2^c: digit 0, carry 1
        return mag ? (mag | DIGIT_NEG) : 0u;
    }
    carry = 0;
    return d;
}

}  // namespace msm
