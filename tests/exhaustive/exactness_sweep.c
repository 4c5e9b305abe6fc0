/* Proof by exhaustion of the float32 forms that csrc/rf_tracer.cuh uses in place of the
 * reference's float64-typed sub-expressions. IEEE-754 arithmetic is identical on the CPU
 * and the GPU, so sweeping the input domain here settles it for the kernel.
 *
 *   exactness_sweep <stride>     stride 1 = every float32 of the domain (about a minute on
 *                                8 cores; run once, result recorded in DESIGN.md), larger
 *                                strides for the unit tests.
 * Prints "sky <n> <bad_a> <bad_b0> <bad_b1> <bad_b2>" and "lens <n> <bad>"; exit 1 on any
 * mismatch. Build: gcc -O2 -fopenmp -ffp-contract=off (no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float bits_to_float(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int main(int argc, char **argv) {
    const int64_t stride = argc > 1 ? atoll(argv[1]) : 1;
    const float c7 = 0.7f;

    /* sky gradient terms (reference physics.py:183-193) for |ny| <= 2, both signs */
    long bad_a = 0, bad_b0 = 0, bad_b1 = 0, bad_b2 = 0, n = 0;
#pragma omp parallel for reduction(+ : bad_a, bad_b0, bad_b1, bad_b2, n) schedule(static)
    for (int64_t i = 0; i <= 0x40000000LL * 2 + 1; i += stride) {
        const uint32_t b = (uint32_t)(i >> 1) | ((i & 1) ? 0x80000000u : 0);
        const float ny = bits_to_float(b);
        const double k = ((double)ny + 1.0) * 0.5;
        const float a = (float)(1.0 - k), b0 = (float)(k * 0.5), b1 = (float)(k * (double)c7),
                    b2 = (float)k;
        const float up = ny + 1.0f;
        n++;
        bad_a += (a != 0.5f * (1.0f - ny));
        bad_b0 += (b0 != 0.25f * up);
        bad_b1 += (b1 != fmaf(ny, c7 * 0.5f, c7 * 0.5f));
        bad_b2 += (b2 != 0.5f * up);
    }
    printf("sky %ld %ld %ld %ld %ld\n", n, bad_a, bad_b0, bad_b1, bad_b2);

    /* aperture offset (reference camera.py:327-334): f32(f64(p) * float64(0.05)) for the
     * disc sampler's outputs, p == 0 or 2^-24 <= |p| <= 1 */
    const double lens = 0.1 / 2.0;
    const float hi = (float)lens, lo = (float)(lens - (double)hi);
    long bad_l = 0, nl = 0;
#pragma omp parallel for reduction(+ : bad_l, nl) schedule(static)
    for (int64_t i = 0; i <= 0x3f800000LL * 2 + 1; i += stride) {
        const uint32_t b = (uint32_t)(i >> 1) | ((i & 1) ? 0x80000000u : 0);
        const float p = bits_to_float(b);
        if (p != 0.0f && fabsf(p) < 0x1p-24f) continue;
        nl++;
        bad_l += ((float)((double)p * lens) != fmaf(p, hi, p * lo));
    }
    printf("lens %ld %ld %a %a\n", nl, bad_l, hi, lo);
    return (bad_a | bad_b0 | bad_b1 | bad_b2 | bad_l) != 0;
}
