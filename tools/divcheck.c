/* Evidence for the hoisted-reciprocal division of the quantise kernels (csrc/common.cuh RowDiv):
 *   y = RN(1/b) once per row;  q0 = RN(a*y);  r0 = fma(-q0,b,a);  q1 = fma(r0,y,q0);  r1 = fma(-q1,b,a);
 *   q = fma(r1,y,q1)
 * compared with the correctly rounded a/b over the guarded domain the kernel uses it on:
 *   2^-30 <= b <= 2^60 (row level),  |a| <= 2^60 (row level),  |q0| >= 2^-30 (element level; otherwise the
 *   kernel falls back to the plain IEEE division).
 * build/run: gcc -O2 -fopenmp -ffp-contract=off -o /tmp/divcheck tools/divcheck.c -lm && /tmp/divcheck [millions] */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float bf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t fb(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

int main(int argc, char **argv)
{
    long millions = argc > 1 ? atol(argv[1]) : 2000;
    long bad = 0, total = 0, skipped = 0;
#pragma omp parallel reduction(+ : bad, total, skipped)
    {
        uint64_t s = 0x9e3779b97f4a7c15ull;
#ifdef _OPENMP
        extern int omp_get_thread_num(void);
        s ^= (uint64_t)(omp_get_thread_num() + 1) * 0xbf58476d1ce4e5b9ull;
#endif
#pragma omp for schedule(static)
        for (long m = 0; m < millions; ++m) {
            for (int i = 0; i < 1000000; ++i) {
                uint64_t x = rng(&s), z = rng(&s);
                uint32_t eb = 127 - 30 + (uint32_t)((x >> 8) % 91);            /* b in [2^-30, 2^61) */
                /* a: any exponent up to 2^60, including denormals and zero */
                uint32_t ea = (uint32_t)(x % 188);                               /* 0 .. 187 = 2^60 */
                if (((x >> 40) & 3) == 0) ea = eb - 30 + (uint32_t)((x >> 44) % 4) - 2; /* around the |q0| guard */
                if ((int32_t)ea < 0) ea = 0;
                if (ea > 187) ea = 187;
                uint32_t ma = (uint32_t)(z & 0x7fffff), mb = (uint32_t)((z >> 23) & 0x7fffff);
                switch ((x >> 20) % 12) {
                case 0: mb = 0x7fffff; break;
                case 1: mb = 0; break;
                case 2: mb = 0x7ffffe - (uint32_t)((z >> 50) & 7); break;
                case 3: mb = (uint32_t)((z >> 50) & 7); break;
                case 4: ma = 0x7fffff - (uint32_t)((z >> 53) & 7); break;
                case 5: ma = (uint32_t)((z >> 53) & 7); break;
                default: break;
                }
                float a = bf(((uint32_t)(x >> 63) << 31) | (ea << 23) | ma);
                float b = bf((eb << 23) | mb);
                if (b > 0x1p60f) b = 0x1p60f;
                float y = 1.0f / b;
                float q0 = a * y;
                if (!(fabsf(q0) >= 0x1p-30f)) { ++skipped; continue; }           /* kernel: plain division */
                float r0 = fmaf(-q0, b, a);
                float q1 = fmaf(r0, y, q0);
                float r1 = fmaf(-q1, b, a);
                float q = fmaf(r1, y, q1), w = a / b;
                ++total;
                if (fb(q) != fb(w)) {
                    if (bad < 5) printf("MISMATCH a=%a b=%a got %a want %a\n", a, b, q, w);
                    ++bad;
                }
            }
        }
    }
    printf("%ld pairs checked, %ld outside the element guard, %ld mismatches\n", total, skipped, bad);
    return bad != 0;
}
