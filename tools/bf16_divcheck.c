/* bf16_divcheck.c — exhaustive check of the claim behind Bf16RowPacked (csrc/common.cuh):
 *     RN_bf16( RN_f32( x * RN_f32(1/s) ) )  ==  RN_bf16( RN_f32( x / s ) )
 * for bf16 x and bf16 s > 0, i.e. the packed fast path reproduces torch's bf16 division (fp32 division, result rounded
 * to bf16).  Scaling both operands by powers of two is exact, so all 2^7 x 2^7 significand pairs x both signs x a band of
 * exponents (well inside the fast path's guard) is exhaustive.   gcc -O2 -o bf16_divcheck bf16_divcheck.c -lm && ./bf16_divcheck
 * Also checks the rint identity  rne(t) == (t + 128) - 128  in bf16 arithmetic for every bf16 t in [0, 128). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static float bf16_round(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return f;
    u += 0x7fffu + ((u >> 16) & 1u);          /* round to nearest even on the upper 16 bits */
    u &= 0xffff0000u;
    memcpy(&f, &u, 4);
    return f;
}
static float bf16_from_bits(uint16_t b)
{
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int main(void)
{
    long long pairs = 0, bad = 0;
    for (int es = 127 - 30; es <= 127 + 30; es += 3)
        for (int ms = 0; ms < 128; ++ms) {
            const float s = bf16_from_bits((uint16_t)((es << 7) | ms));
            volatile float rcp = 1.0f / s;
            for (int ex = 127 - 30; ex <= 127 + 60; ex += 3)
                for (int mx = 0; mx < 128; ++mx)
                    for (int sign = 0; sign < 2; ++sign) {
                        const float x = bf16_from_bits((uint16_t)((sign << 15) | (ex << 7) | mx));
                        volatile float q_exact = x / s;
                        volatile float q_fast = x * rcp;
                        ++pairs;
                        if (bf16_round(q_exact) != bf16_round(q_fast)) {
                            if (bad < 10) printf("MISMATCH x=%a s=%a exact=%a fast=%a\n", x, s, q_exact, q_fast);
                            ++bad;
                        }
                    }
        }
    long long rbad = 0, rn = 0;
    for (uint32_t b = 0; b < 0x4300; ++b) {                      /* all bf16 in [0, 128) */
        const float t = bf16_from_bits((uint16_t)b);
        const float r = bf16_round(bf16_round(t + 128.0f) - 128.0f);
        ++rn;
        if (r != rintf(t)) { if (rbad < 10) printf("RINT MISMATCH t=%a got %a want %a\n", t, r, rintf(t)); ++rbad; }
    }
    printf("division: %lld pairs, %lld mismatches; rint identity: %lld values, %lld mismatches\n", pairs, bad, rn, rbad);
    return (bad || rbad) ? 1 : 0;
}
