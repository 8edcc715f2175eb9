// Generic fp32-accumulate direct convolution / transposed convolution on NHWC feature maps.
//
// This is the exact-arithmetic engine (AIVC_ENGINE_SIMT): every product and sum is an fp32
// FFMA in a fixed order, like the reference's fp32 CPU path, so it serves (a) the layers
// whose channel counts do not fill a tensor-core tile (3/6/9-channel pixel ends), (b) the
// hyperprior transforms whose output (sigma) steers the range coder, and (c) as the
// "fp32" precision mode for parity runs.  The tcgen05 engine lives in conv_tc.cu.
//
// Replaces: CustomConvLayer / UpscalingLayer / GDN forward
//           (layers/misc/custom_conv_layers.py:129-253, layers/misc/misc_layers.py:113-154).
#include "common.cuh"
extern int g_aivc_kernel_class;

namespace {

constexpr int BM = 64;        // output pixels per block (8 x 8 in the phase grid)
constexpr int BN = 64;        // output channels per block
constexpr int BK = 16;        // input channels per slab
constexpr int NT = 256;       // threads
constexpr int MAX_TAPS = 25;

struct SimtParams {
    FMap in, out, res, gate, gx;
    const float *w;           // [tap][cin][cout]
    const float *bias;
    const float *out_scale;
    int cin, cout;
    int act, post, act_channels;
    int in_square;            // GDN pass: square the input on load
    int gdn_mode;             // 0 none, 1 x / sqrt(v), 2 x * sqrt(v)  (x read from gx)
    int clamp;                // 1 replicate (conv), 0 zero outside (transposed conv)
    int in_step;              // input coordinate = m * in_step + dy
    int out_step;             // output coordinate = m * out_step + phase offset
    int mh, mw;               // size of the m-grid handled by this launch
    int nphase;               // blockIdx.z selects the output phase (transposed conv: 4)
    struct Phase {
        int ntaps, out_py, out_px;
        signed char dy[MAX_TAPS], dx[MAX_TAPS];
        unsigned char widx[MAX_TAPS];
    } ph[4];
};

__global__ void __launch_bounds__(NT) conv_simt_kernel(const SimtParams p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const SimtParams::Phase &ph = p.ph[blockIdx.z];
    const int tiles_x = (p.mw + 7) / 8;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
    const int n0 = blockIdx.y * BN;

    const int ty = tid / 16, tx = tid % 16;      // 4 pixels x 4 channels per thread
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // A-slab loader: pixel = tid / 4, channels (tid % 4) * 4 .. +3
    const int lp = tid / 4, lc = (tid % 4) * 4;
    const int lmy = tile_y * 8 + lp / 8, lmx = tile_x * 8 + lp % 8;
    const bool lvalid = (lmy < p.mh) && (lmx < p.mw);

    for (int t = 0; t < ph.ntaps; ++t) {
        int iy = lmy * p.in_step + ph.dy[t];
        int ix = lmx * p.in_step + ph.dx[t];
        bool inb = lvalid;
        if (p.clamp) {
            iy = min(max(iy, 0), p.in.h - 1);
            ix = min(max(ix, 0), p.in.w - 1);
        } else {
            inb = inb && iy >= 0 && iy < p.in.h && ix >= 0 && ix < p.in.w;
        }
        const float *wt = p.w + (size_t)ph.widx[t] * p.cin * p.cout;
        for (int c0 = 0; c0 < p.cin; c0 += BK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = c0 + lc + j;
                float v = 0.f;
                if (inb && ci < p.cin) {
                    v = fm_load(p.in, iy, ix, ci);
                    if (p.in_square) v = v * v;
                }
                As[lc + j][lp] = v;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = tid + j * NT;           // 1024 weights per slab
                const int kk = e / BN, nn = e % BN;
                const int ci = c0 + kk, co = n0 + nn;
                Bs[kk][nn] = (ci < p.cin && co < p.cout) ? wt[(size_t)ci * p.cout + co] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pp = ty * 4 + i;
        const int my = tile_y * 8 + pp / 8, mx = tile_x * 8 + pp % 8;
        if (my >= p.mh || mx >= p.mw) continue;
        const int oy = my * p.out_step + ph.out_py, ox = mx * p.out_step + ph.out_px;
        if (oy >= p.out.h || ox >= p.out.w) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= p.cout) continue;
            float v = acc[i][j] + (p.bias ? p.bias[co] : 0.f);
            if (p.gdn_mode) {
                const float x = fm_load(p.gx, oy, ox, co);
                const float s = sqrtf(v);
                v = (p.gdn_mode == 1) ? x / s : x * s;
            } else if (p.act_channels == 0 || co < p.act_channels) {
                v = act_apply(p.act, v);
            }
            if (p.gate.data) v *= fm_load(p.gate, oy, ox, co);
            if (p.res.data) v += fm_load(p.res, oy, ox, co);
            v = post_apply(p.post, v);
            if (p.out_scale) v *= p.out_scale[co];
            fm_store(p.out, oy, ox, co, v);
        }
    }
}

int launch(const SimtParams &p, cudaStream_t st) {
    dim3 grid(ceil_div(p.mh, 8) * ceil_div(p.mw, 8), ceil_div(p.cout, BN), p.nphase);
    conv_simt_kernel<<<grid, NT, 0, st>>>(p);
    AIVC_CHECK_LAUNCH("conv_simt_kernel");
    return 0;
}

}  // namespace

// Geometry of one (phase of a) convolution as a tap list.
//   conv : out(oy,ox) = sum_{ky,kx} in(clamp(oy*s + ky - k/2), ...) w[ky][kx]
//   tconv: out(2m+py) gets in(m + (py + pad - ky)/2) w[ky] for ky with (py + pad - ky) even,
//          zero outside the input (PyTorch ConvTranspose2d, pad = (k+1)/2 - 1, output_padding 1)
int conv_simt_run(const aivc_conv_op *op, cudaStream_t st) {
    g_aivc_kernel_class = AIVC_KC_SIMT;
    const int k = op->k;
    if (k * k > MAX_TAPS) AIVC_FAIL("conv_simt: kernel size %d unsupported", k);
    const bool has_gdn = (op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN);
    if (has_gdn && !op->scratch) AIVC_FAIL("conv_simt: GDN needs a scratch buffer");
    if (has_gdn && op->act_channels) AIVC_FAIL("conv_simt: GDN with act_channels unsupported");

    SimtParams p;
    memset(&p, 0, sizeof(p));
    p.in = to_dev(op->in);
    p.cin = op->in.c;
    p.cout = op->out.c;
    p.w = (const float *)op->weight;
    p.bias = op->bias;
    p.act_channels = op->act_channels;

    FMap tmp;   // pre-GDN activations, fp32, no border
    if (has_gdn) {
        tmp = to_dev(op->out);
        tmp.data = op->scratch; tmp.c_off = 0; tmp.c_stride = op->out.c; tmp.pad = 0;
        tmp.pitch = op->out.w; tmp.rows = op->out.h; tmp.dtype = AIVC_F32;
        p.out = tmp;
        p.act = AIVC_ACT_NONE;
        p.post = AIVC_POST_NONE;
    } else {
        p.out = to_dev(op->out);
        p.act = op->act;
        p.post = op->post;
        p.out_scale = op->out_scale;
        if (op->residual.data) p.res = to_dev(op->residual);
        if (op->gate.data) p.gate = to_dev(op->gate);
    }

    if (op->kind == 0) {
        p.clamp = 1;
        p.in_step = op->stride;
        p.out_step = 1;
        p.mh = op->out.h; p.mw = op->out.w;
        p.nphase = 1;
        p.ph[0].ntaps = k * k;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
                const int t = ky * k + kx;
                p.ph[0].dy[t] = (signed char)(ky - k / 2);
                p.ph[0].dx[t] = (signed char)(kx - k / 2);
                p.ph[0].widx[t] = (unsigned char)t;
            }
        if (launch(p, st)) return 1;
    } else {
        const int pad = (k + 1) / 2 - 1;
        p.clamp = 0;
        p.in_step = 1;
        p.out_step = 2;
        p.mh = op->in.h; p.mw = op->in.w;
        p.nphase = 4;                              // the four output parities in one launch
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                SimtParams::Phase &ph = p.ph[py * 2 + px];
                int n = 0;
                for (int ky = 0; ky < k; ++ky) {
                    if ((py + pad - ky) & 1) continue;
                    for (int kx = 0; kx < k; ++kx) {
                        if ((px + pad - kx) & 1) continue;
                        ph.dy[n] = (signed char)((py + pad - ky) / 2);
                        ph.dx[n] = (signed char)((px + pad - kx) / 2);
                        ph.widx[n] = (unsigned char)(ky * k + kx);
                        ++n;
                    }
                }
                ph.ntaps = n;
                ph.out_py = py; ph.out_px = px;
            }
        if (launch(p, st)) return 1;
    }

    if (has_gdn) {
        // norm = beta + gamma . x^2 as a 1x1 conv over the squared scratch, then x */ sqrt(norm)
        SimtParams g;
        memset(&g, 0, sizeof(g));
        g.in = tmp; g.gx = tmp;
        g.out = to_dev(op->out);
        g.cin = op->out.c; g.cout = op->out.c;
        g.w = (const float *)op->gdn_gamma;      // [j][i] (transposed), see aivc_b200.h
        g.bias = op->gdn_beta;
        g.in_square = 1;
        g.gdn_mode = (op->act == AIVC_ACT_GDN) ? 1 : 2;
        g.post = op->post;
        g.out_scale = op->out_scale;
        if (op->residual.data) g.res = to_dev(op->residual);
        if (op->gate.data) g.gate = to_dev(op->gate);
        g.clamp = 1; g.in_step = 1; g.out_step = 1;
        g.mh = op->out.h; g.mw = op->out.w;
        g.nphase = 1;
        g.ph[0].ntaps = 1; g.ph[0].dy[0] = 0; g.ph[0].dx[0] = 0; g.ph[0].widx[0] = 0;
        if (launch(g, st)) return 1;
    }
    return 0;
}
