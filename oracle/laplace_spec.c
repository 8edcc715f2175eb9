/* ORACLE (test infrastructure only).
 *
 * CPU restatement of the integer Laplace CDF the codec feeds to the range coder:
 * src/real_life/bitstream.py:127-154 (get_y_cdf) followed by torchac's
 * float -> int16 normalisation (call sites bitstream.py:281,454).  All steps are fp32
 * exactly as torch performs them, except expm1, which is evaluated in double by a
 * fixed add/mul/fma recipe (range reduction by ln2 split in two, 12-term Taylor) and
 * rounded once to fp32 -- the arithmetic DESIGN.md specifies so that host and device
 * agree bit for bit.  Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static const double INV_FACT[11] = { /* 1/12! ... 1/2! */
    2.08767569878680989792e-09, 2.50521083854417187751e-08, 2.75573192239858906526e-07,
    2.75573192239858906526e-06, 2.48015873015873015873e-05, 1.98412698412698412698e-04,
    1.38888888888888888889e-03, 8.33333333333333333333e-03, 4.16666666666666666667e-02,
    1.66666666666666666667e-01, 0.5};

static double expm1_neg_spec(double x) {
    if (x <= -20.0) return -1.0;
    double k = rint(x * 1.4426950408889634);
    double r = fma(-k, 6.93147180369123816490e-01, x);
    r = fma(-k, 1.90821492927058770002e-10, r);
    double acc = INV_FACT[0];
    for (int j = 1; j < 11; ++j) acc = fma(acc, r, INV_FACT[j]);
    double er = acc * r;
    double p = fma(er, r, r);
    double two_k = ldexp(1.0, (int)k);
    return fma(two_k, p, two_k - 1.0);
}

/* cdf_int for entry i (0..513) given sigma (fp32) */
uint32_t laplace_spec_cdf_int(float sigma, int i) {
    volatile float sqrt2 = 1.41421354f;
    float b = sigma / sqrt2;
    float t = (float)i - 256.5f;
    float x = -fabsf(t) / b;
    float e = (float)expm1_neg_spec((double)x);
    float half_sign = (t < 0.0f) ? -0.5f : 0.5f;
    float prod = half_sign * e;
    float cdf = 0.5f - prod;
    float scaled = rintf(cdf * 65023.0f);
    return ((uint32_t)(int32_t)scaled + (uint32_t)i) & 0xFFFFu;
}

/* full [n, 514] uint16 table from n sigmas */
void laplace_spec_table(const float *sigma, size_t n, uint16_t *out) {
    for (size_t s = 0; s < n; ++s)
        for (int i = 0; i < 514; ++i) out[s * 514 + i] = (uint16_t)laplace_spec_cdf_int(sigma[s], i);
}

/* per-symbol bounds for symbols q in [-256, 255] */
void laplace_spec_bounds(const float *sigma, const int16_t *q, size_t n, uint32_t *c_low,
                         uint32_t *c_high) {
    for (size_t s = 0; s < n; ++s) {
        int sym = (int)q[s] + 256;
        c_low[s] = laplace_spec_cdf_int(sigma[s], sym);
        c_high[s] = laplace_spec_cdf_int(sigma[s], sym + 1);
    }
}
