// Laplace CDF -> 16-bit integer CDF, evaluated identically on the device (encoder-side
// bounds kernel) and on the host (range decoder's per-symbol search).
//
// Follows src/real_life/bitstream.py:127-154 (get_y_cdf: Laplace(0, sigma/sqrt(2)).cdf on the
// grid idx = i - 256.5, i in [0, 514)) and torchac's float->int16 normalisation as called at
// bitstream.py:281,454 (round(cdf * (2^16 - 513)) + i, mod 2^16).  Every step is the fp32
// operation torch performs, in the same order, except expm1f: torch's vectorised expm1f is
// not reproducible across devices, so it is replaced by a double-precision evaluation built
// only from IEEE add/mul/fma (bit-identical on x86-64 and sm_100a) and rounded once to fp32.
// Encoder and decoder therefore always agree, on any mix of host and device.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDA_ARCH__)
#define AIVC_HD __host__ __device__ __forceinline__
#define AIVC_FMA(a, b, c) __fma_rn((a), (b), (c))
#define AIVC_MUL(a, b) __dmul_rn((a), (b))
#define AIVC_SUB(a, b) __dsub_rn((a), (b))
#define AIVC_FDIV(a, b) __fdiv_rn((a), (b))
#define AIVC_FMULF(a, b) __fmul_rn((a), (b))
#define AIVC_FSUBF(a, b) __fsub_rn((a), (b))
#define AIVC_FADDF(a, b) __fadd_rn((a), (b))
#else
#if defined(__CUDACC__)
#define AIVC_HD __host__ __device__ inline
#else
#define AIVC_HD static inline
#endif
// host: translation units including this header are built with -ffp-contract=off
#define AIVC_FMA(a, b, c) fma((a), (b), (c))
#define AIVC_MUL(a, b) ((a) * (b))
#define AIVC_SUB(a, b) ((a) - (b))
#define AIVC_FDIV(a, b) ((a) / (b))
#define AIVC_FMULF(a, b) ((a) * (b))
#define AIVC_FSUBF(a, b) ((a) - (b))
#define AIVC_FADDF(a, b) ((a) + (b))
#endif

#define AIVC_AC_MAX_VAL 256
#define AIVC_AC_LP 514            /* 2*AC_MAX_VAL + 2, bitstream.py:134 */
#define AIVC_CDF_SCALE 65023.0f   /* 2^16 - (Lp - 1) */
#define AIVC_WIN_FIRST 253        /* first CDF entry of the 8-entry window handed to the host decoder */
#define AIVC_SQRT2_F 1.41421354f  /* fp32 sqrt(2), as torch.sqrt(torch.tensor([2.0])) */

// e^x = 2^k (1 + p): k = rint(x*log2e), r = x - k*ln2 (two-part ln2), p = Taylor(12)(r) - 1.
// |error| ~1e-16; only IEEE add/mul/fma, so host and device agree bit for bit.
AIVC_HD double aivc_exp_parts(double x, double *two_k_out) {
    const double k = rint(AIVC_MUL(x, 1.4426950408889634));
    double r = AIVC_FMA(-k, 6.93147180369123816490e-01, x);
    r = AIVC_FMA(-k, 1.90821492927058770002e-10, r);
    double e = 2.08767569878680989792e-09;                  // 1/12!
    e = AIVC_FMA(e, r, 2.50521083854417187751e-08);         // 1/11!
    e = AIVC_FMA(e, r, 2.75573192239858906526e-07);         // 1/10!
    e = AIVC_FMA(e, r, 2.75573192239858906526e-06);         // 1/9!
    e = AIVC_FMA(e, r, 2.48015873015873015873e-05);         // 1/8!
    e = AIVC_FMA(e, r, 1.98412698412698412698e-04);         // 1/7!
    e = AIVC_FMA(e, r, 1.38888888888888888889e-03);         // 1/6!
    e = AIVC_FMA(e, r, 8.33333333333333333333e-03);         // 1/5!
    e = AIVC_FMA(e, r, 4.16666666666666666667e-02);         // 1/4!
    e = AIVC_FMA(e, r, 1.66666666666666666667e-01);         // 1/3!
    e = AIVC_FMA(e, r, 0.5);                                // 1/2!
    union { uint64_t u; double d; } two_k;
    two_k.u = (uint64_t)(1023 + (int)k) << 52;              // 2^k, |k| small
    *two_k_out = two_k.d;
    return AIVC_FMA(AIVC_MUL(e, r), r, r);                  // e^r - 1
}

// expm1(x) for x <= 0
AIVC_HD double aivc_expm1_neg(double x) {
    if (x <= -20.0) return -1.0;     // e^x < 2^-28: the fp32 result is exactly -1
    double two_k;
    const double p = aivc_exp_parts(x, &two_k);
    return AIVC_FMA(two_k, p, AIVC_SUB(two_k, 1.0));
}

// sigma = exp(0.5 * clamp(v, LOG_VAR_MIN, LOG_VAR_MAX))   misc_layers.py:214-221
AIVC_HD float aivc_sigma_from_logvar(float v) {
    const float c = fminf(fmaxf(v, -18.4207f), 10.0f);
    double two_k;
    const double p = aivc_exp_parts((double)AIVC_FMULF(0.5f, c), &two_k);
    return (float)AIVC_FMA(two_k, p, two_k);
}

// b = sigma / sqrt(2) in fp32 (bitstream.py:141)
AIVC_HD float aivc_laplace_scale(float sigma) { return AIVC_FDIV(sigma, AIVC_SQRT2_F); }

// 16-bit integer CDF entry i (0 <= i < 514) of the Laplace with scale b.
AIVC_HD uint32_t aivc_laplace_cdf_int(float b, int i) {
    const float t = (float)i - 256.5f;                       // exact
    const float x = AIVC_FDIV(-fabsf(t), b);
    const float e = (float)aivc_expm1_neg((double)x);
    const float hs = t < 0.f ? -0.5f : 0.5f;                 // 0.5 * sign(t); t is never 0
    const float cdf = AIVC_FSUBF(0.5f, AIVC_FMULF(hs, e));
    const float scaled = rintf(AIVC_FMULF(cdf, AIVC_CDF_SCALE));
    return ((uint32_t)(int32_t)scaled + (uint32_t)i) & 0xFFFFu;
}

// fp32 Laplace(0, b) CDF at t (torch.distributions.Laplace.cdf: 0.5 - 0.5 sign(t) expm1(-|t| / b)) and the rate
// estimate of an integer symbol q in bits (pdf_estimator.py:27-70 with zero_mu, entropy_coder.py:25-30):
// -log2 clamp(cdf(q + 1/2) - cdf(q - 1/2), 2^-16, 1).
AIVC_HD float aivc_laplace_cdf_f(float b, float t) {
    const float x = AIVC_FDIV(-fabsf(t), b);
    const float e = (float)aivc_expm1_neg((double)x);
    const float hs = t < 0.f ? -0.5f : (t > 0.f ? 0.5f : 0.f);
    return AIVC_FSUBF(0.5f, AIVC_FMULF(hs, e));
}
AIVC_HD float aivc_laplace_rate_bits(float b, float q) {
    const float p = AIVC_FSUBF(aivc_laplace_cdf_f(b, AIVC_FADDF(q, 0.5f)), aivc_laplace_cdf_f(b, AIVC_FSUBF(q, 0.5f)));
    return -log2f(fminf(fmaxf(p, 1.52587890625e-05f), 1.0f));
}
