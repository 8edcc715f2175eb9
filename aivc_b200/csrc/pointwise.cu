// HBM-bound kernels of the hot path: layout bridges, pixel ends (4:2:0 <-> 4:4:4, 8-bit cast),
// motion-compensation warp + blend, hyperprior mu/sigma, quantisation + integer CDF bounds.
// One thread per pixel (or per symbol); channel counts here are 1..6 at full resolution and
// C_y at latent resolution, so the work is a single pass over the data.
#include "common.cuh"
extern int g_aivc_kernel_class;
#include "laplace_cdf.h"

namespace {

constexpr int PT = 256;

// ---------------------------------------------------------------- layout bridges
__global__ void nchw_to_fmap_kernel(const float *__restrict__ src, FMap dst) {
    const int hw = dst.h * dst.w;
    const size_t n = (size_t)hw * dst.c;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % dst.c);
        const int pix = (int)(i / dst.c);
        fm_store(dst, pix / dst.w, pix % dst.w, ch, src[(size_t)ch * hw + pix]);
    }
}

__global__ void fmap_to_nchw_kernel(FMap src, float *__restrict__ dst) {
    const int hw = src.h * src.w;
    const size_t n = (size_t)hw * src.c;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i % hw);
        const int ch = (int)(i / hw);
        dst[i] = fm_load(src, pix / src.w, pix % src.w, ch);
    }
}

__global__ void fmap_copy_kernel(FMap src, FMap dst) {
    const size_t n = (size_t)dst.h * dst.w * dst.c;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % dst.c), pix = (int)(i / dst.c);
        const int y = pix / dst.w, x = pix % dst.w;
        fm_store(dst, y, x, ch, fm_load(src, y, x, ch));
    }
}

__global__ void fill_border_kernel(FMap m) {
    // one thread per padded pixel of the border ring
    const int hp = m.h + 2 * m.pad, wp = m.w + 2 * m.pad;
    const size_t n = (size_t)hp * wp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int yp = (int)(i / wp), xp = (int)(i % wp);
        const int y = min(max(yp - m.pad, 0), m.h - 1), x = min(max(xp - m.pad, 0), m.w - 1);
        if (y + m.pad == yp && x + m.pad == xp) continue;
        for (int ch = 0; ch < m.c; ++ch) fm_store_raw(m, yp, xp, ch, fm_load(m, y, x, ch));
    }
}

__global__ void fmap_to_i16_kernel(FMap src, int16_t *__restrict__ dst) {
    const int hw = src.h * src.w;
    const size_t n = (size_t)hw * src.c;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i % hw), ch = (int)(i / hw);
        dst[i] = (int16_t)fm_load(src, pix / src.w, pix % src.w, ch);
    }
}

__global__ void i16_to_fmap_kernel(const int16_t *__restrict__ src, FMap dst) {
    const int hw = dst.h * dst.w;
    const size_t n = (size_t)hw * dst.c;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % dst.c), pix = (int)(i / dst.c);
        fm_store(dst, pix / dst.w, pix % dst.w, ch, (float)src[(size_t)ch * hw + pix]);
    }
}

// ---------------------------------------------------------------- InputLayer
template <bool U8>
__global__ void yuv420_to_fmap_kernel(const void *__restrict__ yp, const void *__restrict__ up,
                                      const void *__restrict__ vp, FMap dst, float scale) {
    const int wc = (dst.w + 1) / 2;
    const size_t n = (size_t)dst.h * dst.w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / dst.w), x = (int)(i % dst.w);
        const size_t ci = (size_t)(y / 2) * wc + (x / 2);
        float a, b, c;
        if (U8) {
            a = (float)((const uint8_t *)yp)[i];
            b = (float)((const uint8_t *)up)[ci];
            c = (float)((const uint8_t *)vp)[ci];
            if (scale != 1.f) { a /= 255.f; b /= 255.f; c /= 255.f; }   // [0,1] units, like x/255 in torch
        } else {
            a = ((const float *)yp)[i];
            b = ((const float *)up)[ci];
            c = ((const float *)vp)[ci];
            if (scale == 1.f) { a *= 255.f; b *= 255.f; c *= 255.f; }   // level units
        }
        fm_store(dst, y, x, 0, a);
        fm_store(dst, y, x, 1, b);
        fm_store(dst, y, x, 2, c);
    }
}


// Fused InputLayer for the tensor-core engines (bf16 / split-bf16 buffers): up to three 4:2:0 frames (code, prev, next; uint8 levels, a null
// luma pointer = the all-zero frame) -> channels 0..8 of the 16-channel level-unit pixel buffer, border
// replicas included, channels 9..15 zero: ONE 32-byte store per pixel instead of three launches writing 6
// bytes each.  `dst2` (optional) receives frame 0 alone in channels 0..2 (CodecNet input; its channels
// 3..5 are written later by warp_blend, or stay zero for I frames).
struct PackSrc { const uint8_t *y[3], *u[3], *v[3]; };
__global__ void yuv420_pack16_kernel(PackSrc s, FMap dst, FMap dst2) {
    const int P = dst.pad, W = dst.w + 2 * P, H = dst.h + 2 * P, wc = (dst.w + 1) / 2;
    const size_t n = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int py = (int)(i / W), px = (int)(i % W);
        const int y = min(max(py - P, 0), dst.h - 1), x = min(max(px - P, 0), dst.w - 1);
        const size_t li = (size_t)y * dst.w + x, ci = (size_t)(y / 2) * wc + (x / 2);
        float c[9];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const bool on = s.y[t] != nullptr;
            c[3 * t] = on ? (float)s.y[t][li] : 0.f;
            c[3 * t + 1] = on ? (float)s.u[t][ci] : 0.f;
            c[3 * t + 2] = on ? (float)s.v[t][ci] : 0.f;
        }
        auto pk = [](float a, float b) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
            return *reinterpret_cast<const uint32_t *>(&b2);
        };
        // (split-bf16 buffers: 8-bit levels are exact in the hi half, the lo half of the 32-element pixel is zero)
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        uint4 *q = reinterpret_cast<uint4 *>((__nv_bfloat16 *)dst.data + ((size_t)py * dst.pitch + px) * dst.c_stride);
        q[0] = make_uint4(pk(c[0], c[1]), pk(c[2], c[3]), pk(c[4], c[5]), pk(c[6], c[7]));
        q[1] = make_uint4(pk(c[8], 0.f), 0u, 0u, 0u);
        if (dst.dtype == AIVC_BF16X2) { q[2] = zero; q[3] = zero; }
        if (dst2.data) {
            uint4 *q2 = reinterpret_cast<uint4 *>((__nv_bfloat16 *)dst2.data + ((size_t)py * dst2.pitch + px) * dst2.c_stride);
            q2[0] = make_uint4(pk(c[0], c[1]), pk(c[2], 0.f), 0u, 0u);
            q2[1] = zero;
            if (dst2.dtype == AIVC_BF16X2) { q2[2] = zero; q2[3] = zero; }
        }
    }
}

// ---------------------------------------------------------------- warp
// grid_sample(bilinear, border, align_corners=True) at (x + fx, y + fy), following the fp32
// operation order of func_util/optical_flow.py:28-41 + ATen's grid sampler.
struct Bilerp {
    int x0, y0, x1, y1;
    float w00, w01, w10, w11;   // nw, ne, sw, se
};

__device__ __forceinline__ Bilerp bilerp_setup(int x, int y, float fx, float fy, int w, int h) {
    const float wm = (float)max(w - 1, 1), hm = (float)max(h - 1, 1);
    float gx = 2.0f * ((float)x + fx) / wm - 1.0f;       // normalise (optical_flow.py:31-32)
    float gy = 2.0f * ((float)y + fy) / hm - 1.0f;
    float ix = ((gx + 1.f) / 2.f) * (float)(w - 1);      // un-normalise (align_corners=True)
    float iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));          // border padding = clip coordinates
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
    const float xf = floorf(ix), yf = floorf(iy);
    Bilerp b;
    b.x0 = (int)xf; b.y0 = (int)yf;
    b.x1 = b.x0 + 1; b.y1 = b.y0 + 1;
    const float ex = (xf + 1.f) - ix, ey = (yf + 1.f) - iy;   // distance to the se corner
    const float dx = ix - xf, dy = iy - yf;
    b.w00 = ex * ey; b.w01 = dx * ey; b.w10 = ex * dy; b.w11 = dx * dy;
    return b;
}

template <typename L>
__device__ __forceinline__ float bilerp_sample(const Bilerp &b, int w, int h, L load) {
    // out-of-range corners (x1 == w or y1 == h) carry zero weight; skip their loads
    float acc = 0.f;
    acc += load(b.y0, b.x0) * b.w00;
    if (b.x1 < w) acc += load(b.y0, b.x1) * b.w01;
    if (b.y1 < h) acc += load(b.y1, b.x0) * b.w10;
    if (b.x1 < w && b.y1 < h) acc += load(b.y1, b.x1) * b.w11;
    return acc;
}

__global__ void warp_blend_kernel(FMap mof, FMap prev, FMap next, int frame_is_p, FMap pred,
                                  FMap skip, int levels, float *__restrict__ aux) {
    const int h = pred.h, w = pred.w;
    const size_t n = (size_t)h * w;
    // both references are channel slices 3..5 / 6..8 of ONE 16-channel bf16 pixel buffer (the tensor-core engines' mof_in)
    const bool x2 = prev.dtype == AIVC_BF16X2;                 // split-bf16 pixel: 16 hi + 16 lo elements
    const bool fast = (prev.dtype == AIVC_BF16 || x2) && next.dtype == prev.dtype && prev.data == next.data &&
                      prev.c_stride == (x2 ? 32 : 16) && prev.c_off == 3 && next.c_off == 6 && prev.pad == next.pad &&
                      prev.pitch == next.pitch && ((uintptr_t)prev.data & 15) == 0;
    const int pix_elems = x2 ? 32 : 16;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i % w);
        const float alpha = fminf(fmaxf(fm_load(mof, y, x, 0) + 0.5f, 0.f), 1.f);
        float beta = fminf(fmaxf(fm_load(mof, y, x, 1) + 0.5f, 0.f), 1.f);
        const Bilerp bp = bilerp_setup(x, y, fm_load(mof, y, x, 2), fm_load(mof, y, x, 3), w, h);
        Bilerp bn;
        if (frame_is_p) {
            beta = 1.f;
            bn = bilerp_setup(x, y, 0.f, 0.f, w, h);
        } else {
            bn = bilerp_setup(x, y, fm_load(mof, y, x, 4), fm_load(mof, y, x, 5), w, h);
        }
        float a3[3], b3[3];
        if (fast) {
            // 16-channel level-unit pixels (prev = channels 3..5, next = 6..8 of the same 32-byte pixel):
            // one 16-byte load per prev corner, an 8- and a 4-byte load per next corner
            const __nv_bfloat16 *base = (const __nv_bfloat16 *)prev.data;
            auto pix = [&](int yy, int xx) { return base + ((size_t)(yy + prev.pad) * prev.pitch + (xx + prev.pad)) * pix_elems; };
            auto hi16 = [](uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); };
            auto lo16 = [](uint32_t v) { return __uint_as_float(v << 16); };
            a3[0] = a3[1] = a3[2] = b3[0] = b3[1] = b3[2] = 0.f;
            auto tap_prev = [&](int yy, int xx, float wgt) {
                const uint4 v = *reinterpret_cast<const uint4 *>(pix(yy, xx));          // channels 0..7
                float c0 = hi16(v.y), c1 = lo16(v.z), c2 = hi16(v.z);
                if (x2) {                                                               // + lo halves
                    const uint4 l = *reinterpret_cast<const uint4 *>(pix(yy, xx) + 16);
                    c0 += hi16(l.y); c1 += lo16(l.z); c2 += hi16(l.z);
                }
                a3[0] += c0 * wgt; a3[1] += c1 * wgt; a3[2] += c2 * wgt;
            };
            auto tap_next = [&](int yy, int xx, float wgt) {
                const __nv_bfloat16 *q = pix(yy, xx);
                const uint32_t c67 = *reinterpret_cast<const uint32_t *>(q + 6), c89 = *reinterpret_cast<const uint32_t *>(q + 8);
                float c0 = lo16(c67), c1 = hi16(c67), c2 = lo16(c89);
                if (x2) {
                    const uint32_t l67 = *reinterpret_cast<const uint32_t *>(q + 22), l89 = *reinterpret_cast<const uint32_t *>(q + 24);
                    c0 += lo16(l67); c1 += hi16(l67); c2 += lo16(l89);
                }
                b3[0] += c0 * wgt; b3[1] += c1 * wgt; b3[2] += c2 * wgt;
            };
            tap_prev(bp.y0, bp.x0, bp.w00);
            if (bp.x1 < w) tap_prev(bp.y0, bp.x1, bp.w01);
            if (bp.y1 < h) tap_prev(bp.y1, bp.x0, bp.w10);
            if (bp.x1 < w && bp.y1 < h) tap_prev(bp.y1, bp.x1, bp.w11);
            tap_next(bn.y0, bn.x0, bn.w00);
            if (bn.x1 < w) tap_next(bn.y0, bn.x1, bn.w01);
            if (bn.y1 < h) tap_next(bn.y1, bn.x0, bn.w10);
            if (bn.x1 < w && bn.y1 < h) tap_next(bn.y1, bn.x1, bn.w11);
        } else {
            for (int ch = 0; ch < 3; ++ch) {
                a3[ch] = bilerp_sample(bp, w, h, [&](int yy, int xx) { return fm_load(prev, yy, xx, ch); });
                b3[ch] = bilerp_sample(bn, w, h, [&](int yy, int xx) { return fm_load(next, yy, xx, ch); });
            }
        }
        for (int ch = 0; ch < 3; ++ch) {
            float xw = beta * a3[ch] + (1.f - beta) * b3[ch];
            // level-unit refs/pred (bf16 engine: 8-bit levels are exact in bf16); skip is in [0,1]
            fm_store(pred, y, x, ch, xw * alpha);             // warped_ref * alpha  (decode.py:542)
            if (levels) xw /= 255.f;
            fm_store(skip, y, x, ch, (1.f - alpha) * xw);     // (1 - alpha) * warped (decode.py:536)
            if (aux) aux[(size_t)(2 + ch) * n + i] = xw;
        }
        if (aux) { aux[i] = alpha; aux[n + i] = beta; }
    }
}

__global__ void warp_blend_nchw_kernel(const float *__restrict__ prev, const float *__restrict__ next,
                                       const float *__restrict__ vp, const float *__restrict__ vn,
                                       const float *__restrict__ beta, float *__restrict__ out, int h,
                                       int w) {
    const size_t n = (size_t)h * w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i % w);
        const Bilerp bp = bilerp_setup(x, y, vp[i], vp[n + i], w, h);
        const Bilerp bn = bilerp_setup(x, y, vn[i], vn[n + i], w, h);
        for (int ch = 0; ch < 3; ++ch) {
            const float *pc = prev + ch * n, *nc = next + ch * n;
            const float a = bilerp_sample(bp, w, h, [&](int yy, int xx) { return pc[(size_t)yy * w + xx]; });
            const float b = bilerp_sample(bn, w, h, [&](int yy, int xx) { return nc[(size_t)yy * w + xx]; });
            const float be = beta[ch * n + i];
            out[ch * n + i] = be * a + (1.f - be) * b;
        }
    }
}

// ---------------------------------------------------------------- OutputLayer + 8-bit cast
__device__ __forceinline__ float level8(float v) {
    return rintf(255.f * fminf(fmaxf(v, 0.f), 1.f));
}

__global__ void finalize_frame_kernel(FMap codec, FMap skip, uint8_t *__restrict__ yo,
                                      uint8_t *__restrict__ uo, uint8_t *__restrict__ vo, FMap ref,
                                      int h, int w) {
    // one thread per chroma sample = 2x2 luma block
    const int hc = (h + 1) / 2, wc = (w + 1) / 2;
    const int hc_valid = h / 2, wc_valid = w / 2;      // bilinear x0.5 yields floor(h/2) rows
    const size_t n = (size_t)hc * wc;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int cy = (int)(i / wc), cx = (int)(i % wc);
        auto val = [&](int yy, int xx, int ch) {
            float v = fm_load(codec, yy, xx, ch);
            if (skip.data) v += fm_load(skip, yy, xx, ch);
            return v;
        };
        // luma
        float ylev[2][2];
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                const int yy = 2 * cy + dy, xx = 2 * cx + dx;
                if (yy < h && xx < w) {
                    ylev[dy][dx] = level8(val(yy, xx, 0));
                    yo[(size_t)yy * w + xx] = (uint8_t)ylev[dy][dx];
                }
            }
        // chroma: mean of the 2x2 block; the last row/col of an odd-sized frame replicates
        // its neighbour (decode.py:562-571)
        const int sy = min(cy, max(hc_valid - 1, 0)), sx = min(cx, max(wc_valid - 1, 0));
        float lev[2];
        for (int ch = 1; ch < 3; ++ch) {
            float m;
            if (hc_valid == 0 || wc_valid == 0) {
                m = val(min(2 * sy, h - 1), min(2 * sx, w - 1), ch);
            } else {
                const float a = val(2 * sy, 2 * sx, ch), b = val(2 * sy, 2 * sx + 1, ch);
                const float c = val(2 * sy + 1, 2 * sx, ch), d = val(2 * sy + 1, 2 * sx + 1, ch);
                m = 0.5f * (0.5f * a + 0.5f * b) + 0.5f * (0.5f * c + 0.5f * d);
            }
            lev[ch - 1] = level8(m);
        }
        uo[i] = (uint8_t)lev[0];
        vo[i] = (uint8_t)lev[1];
        if (ref.data) {
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const int yy = 2 * cy + dy, xx = 2 * cx + dx;
                    if (yy < h && xx < w) {
                        fm_store(ref, yy, xx, 0, ylev[dy][dx] / 255.f);
                        fm_store(ref, yy, xx, 1, lev[0] / 255.f);
                        fm_store(ref, yy, xx, 2, lev[1] / 255.f);
                    }
                }
        }
    }
}

// ---------------------------------------------------------------- hyperprior / quantisation
__global__ void mu_sigma_nchw_kernel(const float *__restrict__ hs, float *__restrict__ mu,
                                     float *__restrict__ sigma, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        mu[i] = hs[i];
        sigma[i] = aivc_sigma_from_logvar(hs[n + i]);
    }
}

__global__ void quantize_latent_kernel(FMap y, FMap hs, const float *__restrict__ dec_gain,
                                       int16_t *__restrict__ q, uint32_t *__restrict__ bounds,
                                       int32_t *__restrict__ nz, FMap yhat, float *__restrict__ rate) {
    const int c = y.c, hw = y.h * y.w;
    const size_t n = (size_t)c * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c), pix = (int)(i / c);
        const int yy = pix / y.w, xx = pix % y.w;
        const float mu = fm_load(hs, yy, xx, ch);
        const float sigma = aivc_sigma_from_logvar(fm_load(hs, yy, xx, c + ch));
        const float v = fm_load(y, yy, xx, ch);
        const float qf = fminf(fmaxf(rintf(v - mu), -256.f), 255.f);
        const int qi = (int)qf;
        const float b = aivc_laplace_scale(sigma);
        const uint32_t lo = aivc_laplace_cdf_int(b, qi + 256), hi = aivc_laplace_cdf_int(b, qi + 257);
        const size_t o = (size_t)ch * hw + pix;
        q[o] = (int16_t)qi;
        bounds[o] = lo | (hi << 16);
        if (qi != 0) nz[ch] = 1;
        if (yhat.data) fm_store(yhat, yy, xx, ch, (qf + mu) * dec_gain[ch]);
        if (rate) rate[o] = aivc_laplace_rate_bits(b, qf);
    }
}

__global__ void pdf_prob_kernel(const float *__restrict__ y, const float *__restrict__ mu,
                                const float *__restrict__ sigma, int family, int accumulate,
                                float *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mu ? mu[i] : 0.f, s = sigma[i];
        const float up = (y[i] + 0.5f) - m, dn = (y[i] - 0.5f) - m;
        float p;
        if (family == 0) {
            const float b = aivc_laplace_scale(s);
            p = aivc_laplace_cdf_f(b, up) - aivc_laplace_cdf_f(b, dn);
        } else {                                   // torch.distributions.Normal.cdf: 0.5 (1 + erf((x - mu) / (sigma sqrt 2)))
            const float r = 1.f / (s * 1.41421356237309504880f);
            p = 0.5f * (1.f + erff(up * r)) - 0.5f * (1.f + erff(dn * r));
        }
        out[i] = accumulate ? out[i] + p : p;
    }
}

__global__ void laplace_scale_kernel(FMap hs, int c, float *__restrict__ b) {
    const int hw = hs.h * hs.w;
    const size_t n = (size_t)c * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c), pix = (int)(i / c);
        const float sigma = aivc_sigma_from_logvar(fm_load(hs, pix / hs.w, pix % hs.w, c + ch));
        b[(size_t)ch * hw + pix] = aivc_laplace_scale(sigma);
    }
}

// Decoder side: besides the scale b, the eight integer CDF entries i = 253..260 (the bounds of the symbols
// q = -3..+3 around the mode) of every symbol, so that the host range decoder resolves almost every
// symbol with one 16-byte load instead of evaluating the Laplace CDF two or three times.
__global__ void laplace_window_kernel(FMap hs, int c, float *__restrict__ b, uint4 *__restrict__ win) {
    const int hw = hs.h * hs.w;
    const size_t n = (size_t)c * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c), pix = (int)(i / c);
        const float sigma = aivc_sigma_from_logvar(fm_load(hs, pix / hs.w, pix % hs.w, c + ch));
        const float bb = aivc_laplace_scale(sigma);
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = aivc_laplace_cdf_int(bb, AIVC_WIN_FIRST + j);
        const size_t o = (size_t)ch * hw + pix;
        b[o] = bb;
        win[o] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
    }
}

__global__ void dequantize_latent_kernel(const int16_t *__restrict__ q, FMap hs,
                                         const float *__restrict__ dec_gain, FMap yhat) {
    const int c = yhat.c, hw = yhat.h * yhat.w;
    const size_t n = (size_t)c * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c), pix = (int)(i / c);
        const int yy = pix / yhat.w, xx = pix % yhat.w;
        const float mu = fm_load(hs, yy, xx, ch);
        fm_store(yhat, yy, xx, ch, ((float)q[(size_t)ch * hw + pix] + mu) * dec_gain[ch]);
    }
}

// ---------------------------------------------------------------- transposed conv as GEMM + col2im
// The narrow output layers (ConvTranspose2d k5, 128 -> 3 or 6 channels; custom_conv_layers.py:214-224)
// are computed as P[pixel][(ky,kx,co)] = in[pixel][:] . W[:, co, ky, kx] on the tensor cores (a 1x1
// stage with 25*co output channels), after which every P element belongs to exactly one output
// pixel: out[2*iy - pad + ky][2*ix - pad + kx][co] += P[iy][ix][(ky,kx,co)].  This kernel is that
// gather (<= 9 terms per output element), plus bias and activation.
__global__ void col2im_tconv_kernel(FMap P, FMap out, const float *__restrict__ bias, int k, int act) {
    const int co_n = out.c, pad = (k + 1) / 2 - 1;
    const size_t n = (size_t)out.h * out.w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / out.w), ox = (int)(i % out.w);
        float acc[8];
        for (int c = 0; c < co_n; ++c) acc[c] = bias ? bias[c] : 0.f;
        for (int ky = (oy + pad) & 1; ky < k; ky += 2) {
            const int iy = (oy + pad - ky) / 2;
            if (oy + pad - ky < 0 || iy >= P.h) continue;
            for (int kx = (ox + pad) & 1; kx < k; kx += 2) {
                const int ix = (ox + pad - kx) / 2;
                if (ox + pad - kx < 0 || ix >= P.w) continue;
                const int base = (ky * k + kx) * co_n;
                if (P.dtype == AIVC_F16 && co_n == 6) {        // 12-byte group, 4-byte aligned: three half2 loads
                    const __half2 *r2 = reinterpret_cast<const __half2 *>((const __half *)P.data + fm_index(P, iy, ix, base));
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float2 f = __half22float2(r2[c]);
                        acc[2 * c] += f.x; acc[2 * c + 1] += f.y;
                    }
                } else if (P.dtype == AIVC_F16) {
                    const __half *r = (const __half *)P.data + fm_index(P, iy, ix, base);
                    for (int c = 0; c < co_n; ++c) acc[c] += __half2float(r[c]);
                } else {
                    for (int c = 0; c < co_n; ++c) acc[c] += fm_load(P, iy, ix, base + c);
                }
            }
        }
        for (int c = 0; c < co_n; ++c) fm_store(out, oy, ox, c, act_apply(act, acc[c]));
    }
}


// ---------------------------------------------------------------- weight re-layout
__global__ void pack_weight_kernel(const float *__restrict__ src, void *__restrict__ dst, int kind,
                                   int k, int cin, int cout, int engine, int cin_pad, int cout_pad,
                                   int cin_off, float scale) {
    const int taps = k * k;
    const size_t n = (size_t)taps * cout_pad * cin_pad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        int t, ci, co;
        if (engine == AIVC_ENGINE_SIMT) {          // [tap][cin_pad][cout_pad]
            co = (int)(i % cout_pad); ci = (int)((i / cout_pad) % cin_pad);
            t = (int)(i / ((size_t)cout_pad * cin_pad));
        } else {                                   // [tap][cout_pad][cin_pad]
            ci = (int)(i % cin_pad); co = (int)((i / cin_pad) % cout_pad);
            t = (int)(i / ((size_t)cin_pad * cout_pad));
        }
        ci -= cin_off;                             // buffer channel -> weight channel
        float v = 0.f;
        if (ci >= 0 && ci < cin && co < cout) {
            // Conv2d weight [cout][cin][k][k]; ConvTranspose2d weight [cin][cout][k][k]
            const size_t s = (kind == 0) ? (((size_t)co * cin + ci) * taps + t)
                                         : (((size_t)ci * cout + co) * taps + t);
            v = src[s] * scale;
        }
        if (engine == AIVC_ENGINE_SIMT) ((float *)dst)[i] = v;
        else if (engine == AIVC_ENGINE_TC_X3) {    // [tap][cout_pad][hi cin_pad | lo cin_pad]
            const size_t row = i / cin_pad, col = i % cin_pad;
            __nv_bfloat16 *d = (__nv_bfloat16 *)dst + row * 2 * cin_pad + col;
            bf16_split(v, d[0], d[cin_pad]);
        } else ((__nv_bfloat16 *)dst)[i] = __float2bfloat16_rn(v);
    }
}


// kind 3: space-to-depth repack of a bf16 map whose pixels are 16-byte multiples (see aivc_b200.h).
// One thread per (output pixel incl. border, dy, dx): copies in.c channels.
__global__ void space_to_depth_kernel(FMap in, FMap out) {
    const int vec = in.c / 8;                                  // uint4 per source pixel
    const size_t n = (size_t)(out.h + 2 * out.pad) * (out.w + 2 * out.pad) * 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i & 3);
        const size_t pix = i >> 2;
        const int px = (int)(pix % (out.w + 2 * out.pad)), py = (int)(pix / (out.w + 2 * out.pad));
        const int sy = min(max(2 * (py - out.pad) + (q >> 1), 0), in.h - 1);
        const int sx = min(max(2 * (px - out.pad) + (q & 1), 0), in.w - 1);
        const uint4 *s = reinterpret_cast<const uint4 *>((const __nv_bfloat16 *)in.data + fm_index(in, sy, sx, 0));
        uint4 *d = reinterpret_cast<uint4 *>((__nv_bfloat16 *)out.data + ((size_t)py * out.pitch + px) * out.c_stride +
                                             out.c_off + q * in.c);
        for (int v = 0; v < vec; ++v) d[v] = s[v];
        if (in.dtype == AIVC_BF16X2) {                         // lo halves: same layout, half a pixel further
            const uint4 *sl = reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(s) + (in.c_stride >> 1));
            uint4 *dl = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(d) + (out.c_stride >> 1));
            for (int v = 0; v < vec; ++v) dl[v] = sl[v];
        }
    }
}

int grid_for(size_t n) {
    size_t g = (n + PT - 1) / PT;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

int col2im_tconv_run(const aivc_conv_op *op, cudaStream_t st) {
    g_aivc_kernel_class = AIVC_KC_COL2IM;
    if (op->out.c > 8) AIVC_FAIL("col2im: at most 8 output channels, got %d", op->out.c);
    if (op->in.c < op->k * op->k * op->out.c) AIVC_FAIL("col2im: input has %d channels, needs %d", op->in.c, op->k * op->k * op->out.c);
    if (op->out.h != 2 * op->in.h || op->out.w != 2 * op->in.w) AIVC_FAIL("col2im: output must be exactly 2x");
    col2im_tconv_kernel<<<grid_for((size_t)op->out.h * op->out.w), PT, 0, st>>>(to_dev(op->in), to_dev(op->out), op->bias,
                                                                                op->k, op->act);
    AIVC_CHECK_LAUNCH("col2im_tconv");
    return 0;
}


int space_to_depth_run(const aivc_conv_op *op, cudaStream_t st) {
    g_aivc_kernel_class = AIVC_KC_S2D;
    const aivc_fmap &in = op->in, &out = op->out;
    if ((in.dtype != AIVC_BF16 && in.dtype != AIVC_BF16X2) || out.dtype != in.dtype) AIVC_FAIL("space_to_depth: (split) bf16 maps only");
    const int al = in.dtype == AIVC_BF16X2 ? 16 : 8;           // split maps: both halves 16-byte aligned
    if (in.c % 8 || in.c_off % 8 || in.c_stride % al || out.c_off % 8 || out.c_stride % al)
        AIVC_FAIL("space_to_depth: channel views must be 16-byte multiples");
    if (out.c != 4 * in.c || out.h != (in.h + 1) / 2 || out.w != (in.w + 1) / 2)
        AIVC_FAIL("space_to_depth: output must be ceil(h/2) x ceil(w/2) x 4c");
    const size_t n = (size_t)(out.h + 2 * out.pad) * (out.w + 2 * out.pad) * 4;
    space_to_depth_kernel<<<grid_for(n), 256, 0, st>>>(to_dev(in), to_dev(out));
    AIVC_CHECK_LAUNCH("space_to_depth_kernel");
    return 0;
}

extern "C" {

int aivc_nchw_to_fmap(const float *src, const aivc_fmap *dst, void *stream) {
    if (validate_fmap(dst, "nchw_to_fmap dst")) return 1;
    nchw_to_fmap_kernel<<<grid_for((size_t)dst->h * dst->w * dst->c), PT, 0, (cudaStream_t)stream>>>(
        src, to_dev(*dst));
    AIVC_CHECK_LAUNCH("nchw_to_fmap");
    return 0;
}

int aivc_fmap_to_nchw(const aivc_fmap *src, float *dst, void *stream) {
    if (validate_fmap(src, "fmap_to_nchw src")) return 1;
    fmap_to_nchw_kernel<<<grid_for((size_t)src->h * src->w * src->c), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*src), dst);
    AIVC_CHECK_LAUNCH("fmap_to_nchw");
    return 0;
}

int aivc_fmap_copy(const aivc_fmap *src, const aivc_fmap *dst, void *stream) {
    if (validate_fmap(src, "fmap_copy src") || validate_fmap(dst, "fmap_copy dst")) return 1;
    if (src->h != dst->h || src->w != dst->w || src->c != dst->c) AIVC_FAIL("fmap_copy: shape mismatch");
    fmap_copy_kernel<<<grid_for((size_t)dst->h * dst->w * dst->c), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*src), to_dev(*dst));
    AIVC_CHECK_LAUNCH("fmap_copy");
    return 0;
}

int aivc_fill_border(const aivc_fmap *m, void *stream) {
    if (validate_fmap(m, "fill_border")) return 1;
    if (m->pad == 0) return 0;
    fill_border_kernel<<<grid_for((size_t)(m->h + 2 * m->pad) * (m->w + 2 * m->pad)), PT, 0,
                         (cudaStream_t)stream>>>(to_dev(*m));
    AIVC_CHECK_LAUNCH("fill_border");
    return 0;
}

int aivc_fmap_to_i16(const aivc_fmap *src, int16_t *dst, void *stream) {
    if (validate_fmap(src, "fmap_to_i16 src")) return 1;
    fmap_to_i16_kernel<<<grid_for((size_t)src->h * src->w * src->c), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*src), dst);
    AIVC_CHECK_LAUNCH("fmap_to_i16");
    return 0;
}

int aivc_i16_to_fmap(const int16_t *src, const aivc_fmap *dst, void *stream) {
    if (validate_fmap(dst, "i16_to_fmap dst")) return 1;
    i16_to_fmap_kernel<<<grid_for((size_t)dst->h * dst->w * dst->c), PT, 0, (cudaStream_t)stream>>>(
        src, to_dev(*dst));
    AIVC_CHECK_LAUNCH("i16_to_fmap");
    return 0;
}

int aivc_yuv420_to_fmap(const void *y, const void *u, const void *v, int u8, int levels,
                        const aivc_fmap *dst, void *stream) {
    if (validate_fmap(dst, "yuv420_to_fmap dst")) return 1;
    if (dst->c != 3) AIVC_FAIL("yuv420_to_fmap: destination view must have 3 channels, got %d", dst->c);
    const int g = grid_for((size_t)dst->h * dst->w);
    const float scale = levels ? 1.f : 1.f / 255.f;
    if (u8) yuv420_to_fmap_kernel<true><<<g, PT, 0, (cudaStream_t)stream>>>(y, u, v, to_dev(*dst), scale);
    else yuv420_to_fmap_kernel<false><<<g, PT, 0, (cudaStream_t)stream>>>(y, u, v, to_dev(*dst), scale);
    AIVC_CHECK_LAUNCH("yuv420_to_fmap");
    return 0;
}

int aivc_yuv420_pack16(const void *y0, const void *u0, const void *v0, const void *y1, const void *u1,
                       const void *v1, const void *y2, const void *u2, const void *v2, const aivc_fmap *dst,
                       const aivc_fmap *dst2, void *stream) {
    if (validate_fmap(dst, "yuv420_pack16 dst")) return 1;
    const bool x2 = dst->dtype == AIVC_BF16X2;
    if ((dst->dtype != AIVC_BF16 && !x2) || dst->c_stride != (x2 ? 32 : 16) || dst->c_off != 0 || ((uintptr_t)dst->data & 15))
        AIVC_FAIL("yuv420_pack16: destination must be a whole 16-channel (split-)bf16 pixel buffer");
    FMap d2;
    memset(&d2, 0, sizeof(d2));
    if (dst2) {
        if (validate_fmap(dst2, "yuv420_pack16 dst2")) return 1;
        if (dst2->dtype != dst->dtype || dst2->c_stride != dst->c_stride || dst2->c_off != 0 || dst2->h != dst->h || dst2->w != dst->w ||
            dst2->pad != dst->pad || ((uintptr_t)dst2->data & 15))
            AIVC_FAIL("yuv420_pack16: second destination must match the first");
        d2 = to_dev(*dst2);
    }
    PackSrc s;
    s.y[0] = (const uint8_t *)y0; s.u[0] = (const uint8_t *)u0; s.v[0] = (const uint8_t *)v0;
    s.y[1] = (const uint8_t *)y1; s.u[1] = (const uint8_t *)u1; s.v[1] = (const uint8_t *)v1;
    s.y[2] = (const uint8_t *)y2; s.u[2] = (const uint8_t *)u2; s.v[2] = (const uint8_t *)v2;
    const size_t n = (size_t)(dst->h + 2 * dst->pad) * (dst->w + 2 * dst->pad);
    yuv420_pack16_kernel<<<grid_for(n), PT, 0, (cudaStream_t)stream>>>(s, to_dev(*dst), d2);
    AIVC_CHECK_LAUNCH("yuv420_pack16");
    return 0;
}

int aivc_warp_blend(const aivc_fmap *mof, const aivc_fmap *prev, const aivc_fmap *next,
                    int frame_is_p, int levels, const aivc_fmap *pred, const aivc_fmap *skip,
                    float *aux, void *stream) {
    if (validate_fmap(mof, "warp mof") || validate_fmap(prev, "warp prev") ||
        validate_fmap(next, "warp next") || validate_fmap(pred, "warp pred") ||
        validate_fmap(skip, "warp skip"))
        return 1;
    if (mof->c < 6 || mof->h < pred->h || mof->w < pred->w)
        AIVC_FAIL("warp_blend: MOFNet output must be >= 6 ch and cover the frame");
    if (prev->h != pred->h || prev->w != pred->w || next->h != pred->h || next->w != pred->w)
        AIVC_FAIL("warp_blend: reference / prediction size mismatch");
    warp_blend_kernel<<<grid_for((size_t)pred->h * pred->w), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*mof), to_dev(*prev), to_dev(*next), frame_is_p, to_dev(*pred), to_dev(*skip), levels, aux);
    AIVC_CHECK_LAUNCH("warp_blend");
    return 0;
}

int aivc_warp_blend_nchw(const float *prev, const float *next, const float *v_prev,
                         const float *v_next, const float *beta, float *out, int h, int w,
                         void *stream) {
    warp_blend_nchw_kernel<<<grid_for((size_t)h * w), PT, 0, (cudaStream_t)stream>>>(
        prev, next, v_prev, v_next, beta, out, h, w);
    AIVC_CHECK_LAUNCH("warp_blend_nchw");
    return 0;
}

int aivc_finalize_frame(const aivc_fmap *codec, const aivc_fmap *skip, uint8_t *y, uint8_t *u,
                        uint8_t *v, const aivc_fmap *ref444, void *stream) {
    if (validate_fmap(codec, "finalize codec")) return 1;
    FMap sk, rf;
    memset(&sk, 0, sizeof(sk));
    memset(&rf, 0, sizeof(rf));
    int h = codec->h, w = codec->w;
    if (skip && skip->data) {
        if (validate_fmap(skip, "finalize skip")) return 1;
        sk = to_dev(*skip);
        h = skip->h; w = skip->w;
    }
    if (ref444 && ref444->data) {
        if (validate_fmap(ref444, "finalize ref444")) return 1;
        rf = to_dev(*ref444);
        h = ref444->h; w = ref444->w;
    }
    if (codec->h < h || codec->w < w || codec->c < 3) AIVC_FAIL("finalize: codec output too small");
    finalize_frame_kernel<<<grid_for((size_t)((h + 1) / 2) * ((w + 1) / 2)), PT, 0,
                            (cudaStream_t)stream>>>(to_dev(*codec), sk, y, u, v, rf, h, w);
    AIVC_CHECK_LAUNCH("finalize_frame");
    return 0;
}

int aivc_mu_sigma_nchw(const float *hs, float *mu, float *sigma, int c, int hw, void *stream) {
    const size_t n = (size_t)c * hw;
    mu_sigma_nchw_kernel<<<grid_for(n), PT, 0, (cudaStream_t)stream>>>(hs, mu, sigma, n);
    AIVC_CHECK_LAUNCH("mu_sigma_nchw");
    return 0;
}

int aivc_quantize_latent(const aivc_fmap *y, const aivc_fmap *hs, const float *dec_gain, int16_t *q,
                         uint32_t *bounds, int32_t *nz, const aivc_fmap *yhat, float *rate, void *stream) {
    if (validate_fmap(y, "quantize y") || validate_fmap(hs, "quantize hs")) return 1;
    if (hs->c < 2 * y->c || hs->h < y->h || hs->w < y->w)
        AIVC_FAIL("quantize_latent: hyper-decoder output must be >= 2C channels and cover y");
    FMap yh;
    memset(&yh, 0, sizeof(yh));
    if (yhat && yhat->data) {
        if (validate_fmap(yhat, "quantize yhat")) return 1;
        yh = to_dev(*yhat);
    }
    quantize_latent_kernel<<<grid_for((size_t)y->c * y->h * y->w), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*y), to_dev(*hs), dec_gain, q, bounds, nz, yh, rate);
    AIVC_CHECK_LAUNCH("quantize_latent");
    return 0;
}

int aivc_pdf_prob(const float *y, const float *mu, const float *sigma, int family, int accumulate, float *out,
                  size_t n, void *stream) {
    if (!y || !sigma || !out) AIVC_FAIL("pdf_prob: null tensor");
    if (family != 0 && family != 1) AIVC_FAIL("pdf_prob: family %d (0 = laplace, 1 = normal)", family);
    pdf_prob_kernel<<<grid_for(n), PT, 0, (cudaStream_t)stream>>>(y, mu, sigma, family, accumulate, out, n);
    AIVC_CHECK_LAUNCH("pdf_prob");
    return 0;
}

int aivc_laplace_scale(const aivc_fmap *hs, int c, float *b, void *stream) {
    if (validate_fmap(hs, "laplace_scale hs")) return 1;
    if (hs->c < 2 * c) AIVC_FAIL("laplace_scale: need 2C channels");
    laplace_scale_kernel<<<grid_for((size_t)c * hs->h * hs->w), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*hs), c, b);
    AIVC_CHECK_LAUNCH("laplace_scale");
    return 0;
}

int aivc_laplace_window(const aivc_fmap *hs, int c, float *b, uint16_t *win, void *stream) {
    if (validate_fmap(hs, "laplace_window hs")) return 1;
    if (hs->c < 2 * c) AIVC_FAIL("laplace_window: need 2C channels");
    if ((uintptr_t)win & 15) AIVC_FAIL("laplace_window: window buffer must be 16-byte aligned");
    laplace_window_kernel<<<grid_for((size_t)c * hs->h * hs->w), PT, 0, (cudaStream_t)stream>>>(
        to_dev(*hs), c, b, reinterpret_cast<uint4 *>(win));
    AIVC_CHECK_LAUNCH("laplace_window");
    return 0;
}

int aivc_dequantize_latent(const int16_t *q, const aivc_fmap *hs, const float *dec_gain,
                           const aivc_fmap *yhat, void *stream) {
    if (validate_fmap(hs, "dequantize hs") || validate_fmap(yhat, "dequantize yhat")) return 1;
    dequantize_latent_kernel<<<grid_for((size_t)yhat->c * yhat->h * yhat->w), PT, 0,
                               (cudaStream_t)stream>>>(q, to_dev(*hs), dec_gain, to_dev(*yhat));
    AIVC_CHECK_LAUNCH("dequantize_latent");
    return 0;
}

size_t aivc_packed_weight_bytes(int k, int engine, int cin_pad, int cout_pad) {
    return (size_t)k * k * cin_pad * cout_pad * (engine == AIVC_ENGINE_TC ? 2 : 4);     // (TC_X3: two bf16 per weight)
}

int aivc_pack_conv_weight(const float *src, void *dst, int kind, int k, int cin, int cout, int engine,
                          int cin_pad, int cout_pad, int cin_off, float scale, void *stream) {
    if (cin_off < 0 || cin_pad < cin + cin_off || cout_pad < cout)
        AIVC_FAIL("pack_conv_weight: padded sizes smaller than logical sizes");
    const size_t n = (size_t)k * k * cin_pad * cout_pad;
    pack_weight_kernel<<<grid_for(n), PT, 0, (cudaStream_t)stream>>>(src, dst, kind, k, cin, cout,
                                                                     engine, cin_pad, cout_pad, cin_off,
                                                                     scale);
    AIVC_CHECK_LAUNCH("pack_weight");
    return 0;
}

}  // extern "C"
