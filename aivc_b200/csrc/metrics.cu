// Encoder-side quality metrics of one frame on the device:
//   MSE / PSNR over the Y, U, V planes                 model_mngt/loss_function.py:415-435, 234
//   MS-SSIM per plane, weighted by plane size          loss_function.py:438-470, func_util/ms_ssim.py:37-150
// The reference runs, per plane and per scale, five depth-wise 11x11 Gaussian convolutions (window =
// outer product of gaussian(11, 1.5), "valid" = no padding, ms_ssim.py:24-74), then reflection-pads odd
// sizes and average-pools by two (ms_ssim.py:116-126), five scales, and combines
// prod(cs[:4] ** w[:4]) * ssim[4] ** w[4] (ms_ssim.py:138-150).  Here one kernel per scale computes the five
// filtered maps separably in shared memory (11 + 11 taps instead of 121) and reduces the ssim / cs maps to
// per-block partial sums; a second kernel adds the partials in a FIXED order (run-to-run deterministic).
// The separable evaluation rounds differently from a direct 2-D convolution: results agree with the
// reference to ~1e-6, not bit for bit (the metric is logged, never coded).
#include "common.cuh"

namespace {

constexpr int WIN = 11, TX = 32, TY = 8, NT = 256;
constexpr int MAX_SCALES = 5;
__constant__ float c_gauss[WIN + 1][WIN];          // row k: gaussian(k, 1.5) normalised (k = window size actually used)
__constant__ float c_msw[MAX_SCALES] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};     // ms_ssim.py:98-100

__global__ void u8_to_unit_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, float *__restrict__ fa,
                                  float *__restrict__ fb, size_t n, double *__restrict__ sq_partial) {
    // [0,1] floats of both planes + this block's partial sum of squared differences
    __shared__ double red[NT];
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = (float)a[i] / 255.f, y = (float)b[i] / 255.f;
        fa[i] = x; fb[i] = y;
        const float d = x - y;
        acc += (double)(d * d);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) sq_partial[blockIdx.x] = red[0];
}

// one scale: partial[2*block] = sum of ssim map values, partial[2*block+1] = sum of cs map values
__global__ void __launch_bounds__(NT) ssim_scale_kernel(const float *__restrict__ a, const float *__restrict__ b, int h, int w,
                                                        int win, double *__restrict__ partial) {
    __shared__ float sa[TY + WIN - 1][TX + WIN - 1], sb[TY + WIN - 1][TX + WIN - 1];
    __shared__ float sh[5][TY + WIN - 1][TX];
    __shared__ double red[2][NT];
    const int oh = h - win + 1, ow = w - win + 1;               // "valid" output size
    const float *gw = c_gauss[win];
    const int tiles_x = (ow + TX - 1) / TX;
    const int x0 = (blockIdx.x % tiles_x) * TX, y0 = (blockIdx.x / tiles_x) * TY;
    for (int i = threadIdx.x; i < (TY + WIN - 1) * (TX + WIN - 1); i += NT) {
        const int ly = i / (TX + WIN - 1), lx = i % (TX + WIN - 1);
        const int gy = y0 + ly, gx = x0 + lx;
        const bool in = gy < h && gx < w;
        sa[ly][lx] = in ? a[(size_t)gy * w + gx] : 0.f;
        sb[ly][lx] = in ? b[(size_t)gy * w + gx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (TY + WIN - 1) * TX; i += NT) {          // horizontal pass
        const int ly = i / TX, lx = i % TX;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
        for (int k = 0; k < win; ++k) {
            const float g = gw[k], p = sa[ly][lx + k], q = sb[ly][lx + k];
            m1 += g * p; m2 += g * q; e11 += g * (p * p); e22 += g * (q * q); e12 += g * (p * q);
        }
        sh[0][ly][lx] = m1; sh[1][ly][lx] = m2; sh[2][ly][lx] = e11; sh[3][ly][lx] = e22; sh[4][ly][lx] = e12;
    }
    __syncthreads();
    const int lx = threadIdx.x % TX, ly = threadIdx.x / TX;                 // vertical pass: one output per thread
    double s_ssim = 0.0, s_cs = 0.0;
    if (x0 + lx < ow && y0 + ly < oh) {
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            float acc = 0.f;
            for (int k = 0; k < win; ++k) acc += gw[k] * sh[q][ly + k][lx];
            v[q] = acc;
        }
        const float mu1 = v[0], mu2 = v[1];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = v[2] - mu1_sq, s2 = v[3] - mu2_sq, s12 = v[4] - mu12;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;                 // L = 1 (val_range = 1, loss_function.py:443)
        const float v1 = 2.0f * s12 + C2, v2 = s1 + s2 + C2;
        s_cs = (double)(v1 / v2);
        s_ssim = (double)(((2.f * mu12 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2));
    }
    red[0][threadIdx.x] = s_ssim; red[1][threadIdx.x] = s_cs;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = red[0][0]; partial[2 * blockIdx.x + 1] = red[1][0]; }
}

// out[0..m) = sum over blocks of partial[m * block + j], in a fixed order (single block)
__global__ void sum_partials_kernel(const double *__restrict__ partial, int nblocks, int m, double *__restrict__ out, double scale) {
    __shared__ double red[NT];
    for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += NT) acc += partial[(size_t)m * i + j];
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int s = NT / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[j] = red[0] * scale;
        __syncthreads();
    }
}

// ReflectionPad2d((0, w & 1, 0, h & 1)) + avg_pool2d(2)   ms_ssim.py:116-126
__global__ void pool2_reflect_kernel(const float *__restrict__ src, int h, int w, float *__restrict__ dst) {
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    const size_t n = (size_t)oh * ow;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / ow), ox = (int)(i % ow);
        int y1 = 2 * oy + 1, x1 = 2 * ox + 1;
        if (y1 >= h) y1 = h - 2;                                 // reflected row / column (the edge is not repeated)
        if (x1 >= w) x1 = w - 2;
        const float s = src[(size_t)(2 * oy) * w + 2 * ox] + src[(size_t)(2 * oy) * w + x1] +
                        src[(size_t)y1 * w + 2 * ox] + src[(size_t)y1 * w + x1];
        dst[i] = s * 0.25f;
    }
}

// res: [3 planes][5 scales][2] means (ssim, cs) + [30] = sum of squared errors (all planes)
// out: mse, psnr, ms_ssim, ms_ssim_db   (loss_function.py:232-242)
__global__ void combine_kernel(const double *__restrict__ res, double n_y, double n_c, int scales, float *__restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    float total = 0.f;
    for (int p = 0; p < 3; ++p) {
        float prod = 1.f;
        for (int s = 0; s < scales; ++s) {
            const float ssim = (float)res[(p * MAX_SCALES + s) * 2], cs = (float)res[(p * MAX_SCALES + s) * 2 + 1];
            prod *= powf(s == scales - 1 ? ssim : cs, c_msw[s]);     // prod(cs[:-1] ** w[:-1]) * ssim[-1] ** w[-1]
        }
        total += prod * (float)(p == 0 ? n_y : n_c);
    }
    const float ms = total / (float)(n_y + 2.0 * n_c);
    const float mse = (float)(res[30] / (n_y + 2.0 * n_c));
    out[0] = mse;
    out[1] = 10.f * log10f(1.f / mse);
    out[2] = ms;
    out[3] = -10.f * log10f(1.f - ms);
}

int blocks_for(size_t n) { return (int)((n + NT - 1) / NT < 1184 ? (n + NT - 1) / NT : 1184); }

}  // namespace

extern "C" {

// bytes of device scratch aivc_frame_metrics needs for an h x w frame
size_t aivc_frame_metrics_scratch_bytes(int h, int w) {
    const size_t n = (size_t)h * w;
    return 2 * (n + n / 2 + 4096) * sizeof(float) + (2 * 9000 + 64) * sizeof(double);
}

// a, b: the two frames as uint8 4:2:0 planes (device).  out: 4 floats (device): mse, psnr, ms_ssim, ms_ssim_db.
int aivc_frame_metrics(const uint8_t *ya, const uint8_t *ua, const uint8_t *va, const uint8_t *yb, const uint8_t *ub,
                       const uint8_t *vb, int h, int w, void *scratch, size_t scratch_bytes, float *out, void *stream) {
    if (h < 32 || w < 32) AIVC_FAIL("frame_metrics: frame %dx%d too small for five scales", w, h);
    if (scratch_bytes < aivc_frame_metrics_scratch_bytes(h, w)) AIVC_FAIL("frame_metrics: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    static bool window_set = false;
    if (!window_set) {          // gaussian(k, 1.5) for every window size k <= 11 (ms_ssim.py:24-27, 58-60), fp32 like torch.Tensor
        float g[WIN + 1][WIN];
        memset(g, 0, sizeof(g));
        for (int k = 1; k <= WIN; ++k) {
            float sum = 0.f;
            for (int x = 0; x < k; ++x) { g[k][x] = (float)exp(-(double)((x - k / 2) * (x - k / 2)) / (2.0 * 1.5 * 1.5)); sum += g[k][x]; }
            for (int x = 0; x < k; ++x) g[k][x] /= sum;
        }
        AIVC_CHECK_CUDA(cudaMemcpyToSymbol(c_gauss, g, sizeof(g)));
        window_set = true;
    }
    const size_t n = (size_t)h * w, half = n + n / 2 + 4096;
    float *fa = (float *)scratch, *fb = fa + half;
    double *partial = (double *)(fb + half), *res = partial + 2 * 9000;
    AIVC_CHECK_CUDA(cudaMemsetAsync(res, 0, 64 * sizeof(double), st));
    const int hc = (h + 1) / 2, wc = (w + 1) / 2;
    const uint8_t *pa[3] = {ya, ua, va}, *pb[3] = {yb, ub, vb};
    for (int p = 0; p < 3; ++p) {
        int ph = p ? hc : h, pw = p ? wc : w;
        const size_t pn = (size_t)ph * pw;
        const int nb = blocks_for(pn);
        u8_to_unit_kernel<<<nb, NT, 0, st>>>(pa[p], pb[p], fa, fb, pn, partial);
        sum_partials_kernel<<<1, NT, 0, st>>>(partial, nb, 1, res + 31 + p, 1.0);
        float *ca = fa, *cb = fb;
        for (int s = 0; s < MAX_SCALES; ++s) {
            const int win = ph < WIN || pw < WIN ? (ph < pw ? ph : pw) : WIN;      // real_size = min(window_size, h, w)
            const int oh = ph - win + 1, ow = pw - win + 1;
            const int nblk = ((ow + TX - 1) / TX) * ((oh + TY - 1) / TY);
            if (nblk > 9000) AIVC_FAIL("frame_metrics: frame too large for the partial-sum buffer");
            ssim_scale_kernel<<<nblk, NT, 0, st>>>(ca, cb, ph, pw, win, partial);
            sum_partials_kernel<<<1, NT, 0, st>>>(partial, nblk, 2, res + (p * MAX_SCALES + s) * 2, 1.0 / ((double)oh * ow));
            if (s + 1 < MAX_SCALES) {
                const int nh = (ph + 1) / 2, nw = (pw + 1) / 2;
                float *na = ca + (size_t)ph * pw, *nbuf = cb + (size_t)ph * pw;
                pool2_reflect_kernel<<<blocks_for((size_t)nh * nw), NT, 0, st>>>(ca, ph, pw, na);
                pool2_reflect_kernel<<<blocks_for((size_t)nh * nw), NT, 0, st>>>(cb, ph, pw, nbuf);
                ca = na; cb = nbuf; ph = nh; pw = nw;
            }
        }
    }
    // res[30] = total squared error: add the three plane sums (fixed order)
    sum_partials_kernel<<<1, NT, 0, st>>>(res + 31, 3, 1, res + 30, 1.0);
    combine_kernel<<<1, 32, 0, st>>>(res, (double)h * w, (double)hc * wc, MAX_SCALES, out);
    AIVC_CHECK_LAUNCH("frame_metrics");
    return 0;
}

}  // extern "C"
