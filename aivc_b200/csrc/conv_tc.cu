// tcgen05 implicit-GEMM convolution / transposed convolution for sm_100a.
//
//   D[128 pixels x N channels] (fp32, TMEM) += A[128 x BK] (bf16, smem) . B[N x BK]^T (bf16, smem)
//
// * A tiles are *not* gathered: the producing kernel wrote the activation as bordered NHWC, so
//   tap (ky,kx) of a TH x TW pixel tile is one TMA box load at a shifted origin -- a 3-D map
//   for stride 1 and transposed-conv phases, a 5-D (c, x parity, x/2, y parity, y/2) map for
//   stride 2.  Out-of-range rows/columns of partial tiles (and the zero padding of transposed
//   convs) come from TMA's zero fill.
// * B tiles come from weights packed [tap][cout][cin] (K-major).
// * warp 0 = TMA producer, warp 1 = TMEM allocator + tcgen05.mma issuer, warps 2-5 = epilogue
//   (tcgen05.ld 32 lanes x 16 columns at a time): bias, LeakyReLU/ReLU/sigmoid, gate, residual,
//   post activation, per-channel gain, bf16/fp32 store incl. the replicate border.
// * GDN/IGDN is chained on the TMEM-resident tile: the epilogue squares the tile into a
//   swizzled smem A operand, a second tcgen05.mma multiplies it with gamma (smem B operand,
//   TMA-loaded) into a second TMEM region, and the final pass reads both accumulators.
//
// Replaces CustomConvLayer / UpscalingLayer / GDN forward
// (layers/misc/custom_conv_layers.py:129-253, layers/misc/misc_layers.py:113-154).
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tcgen;
extern int g_aivc_kernel_class;

namespace {

constexpr int MAX_TAPS = 25;
constexpr int NTHREADS = 192;
constexpr int MAX_STAGES = 8;

struct Phase {
    int ntaps, out_py, out_px, _pad;
    signed char ax[MAX_TAPS], ay[MAX_TAPS];      // origin offsets (already include border / parity shift)
    unsigned char qx[MAX_TAPS], qy[MAX_TAPS];    // parities for the 5-D (stride-2) map
    unsigned char widx[MAX_TAPS];
};

struct TcParams {
    FMap out, res, gate;
    const float *bias, *gdn_beta, *out_scale;
    int cout, kchunks;
    int act, post, act_channels, gdn;            // gdn: 0 none, 1 divide, 2 multiply
    int tw, th, tiles_x;
    int mh, mw, out_step, mode;                  // mode 0: 3-D map, 1: 5-D map
    int tmem_cols, kg;                           // kg: columns per swizzle atom of the GDN operands
    int nstages, xsq_off;                        // pipeline depth; byte offset of the x^2 operand tile
    int x3, a_lo, b_lo;                          // split-bf16 operands (AIVC_ENGINE_TC_X3): every (tap, chunk) runs as
                                                 // hi.Whi, lo.Whi, hi.Wlo; channel coordinates of the lo halves
    uint32_t stage_bytes, a_bytes, b_bytes, item_bytes;
    Phase ph[4];
};

// ------------------------------------------------------------------------------------ kernel
template <int BK>
__global__ void __launch_bounds__(NTHREADS, 2) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmG,
                                                           const __grid_constant__ TcParams p) {
    constexpr int ROWB = BK * 2;
    constexpr int IPS = 64 / BK;                // (tap, chunk) items per pipeline stage
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc, bar_gamma, bar_xsq, bar_norm;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[256], sscale[256], sbeta[128];

    uint8_t *tiles = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int NSTAGES = p.nstages;
    uint8_t *xsq_tile = tiles + p.xsq_off;            // [chunks][128][kg] bf16; may alias the (drained) stage ring
    uint8_t *gam_tile = tiles + (size_t)NSTAGES * p.stage_bytes + (p.xsq_off ? 128 * p.cout * 2 * (p.x3 ? 2 : 1) : 0);   // [chunks][N][kg]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Phase &ph = p.ph[blockIdx.z];
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x % p.tiles_x;
    const int my0 = tile_y * p.th, mx0 = tile_x * p.tw;
    const int N = p.cout;
    const int kv = p.x3 ? 3 * p.kchunks : p.kchunks;           // virtual chunks per tap
    const int total_it = ph.ntaps * kv;

    if (tid == 0) {
        for (int s = 0; s < NSTAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_acc, 1);
        mbar_init(&bar_gamma, 1);
        mbar_init(&bar_xsq, 128);
        mbar_init(&bar_norm, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    stage_vec(sbias, p.bias, N, 0.f, tid, NTHREADS);
    stage_vec(sscale, p.out_scale, N, 1.f, tid, NTHREADS);
    if (p.gdn) stage_vec(sbeta, p.gdn_beta, N, 0.f, tid, NTHREADS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait_prior_grid();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (p.gdn) {
                const int chunks = (p.x3 ? 2 : 1) * N / p.kg;           // x3: gamma is [N][hi N | lo N]
                mbar_expect_tx(&bar_gamma, (uint32_t)(chunks * p.kg * N * 2));
                for (int c = 0; c < chunks; ++c)
                    tma_load_2d(gam_tile + (size_t)c * N * p.kg * 2, &tmG, &bar_gamma, c * p.kg, 0);
            }
            // A stage holds IPS = 64 / BK (tap, chunk) items, so it always carries four MMAs: with
            // 16- or 32-channel chunks a per-item barrier round trip would cost more than its MMA
            int s = 0, t = 0, kc = 0;
            uint32_t par = 0;                                  // ring slot / phase, advanced incrementally
            for (int done = 0; done < total_it; done += IPS) {
                const int n = min(IPS, total_it - done);
                mbar_wait(&bar_empty[s], par ^ 1u);
                mbar_expect_tx(&bar_full[s], (uint32_t)n * (p.a_bytes + p.b_bytes));
                uint8_t *dst = tiles + (size_t)s * p.stage_bytes;
                for (int g = 0; g < n; ++g, dst += p.item_bytes) {
                    int ca = kc * BK, cb = kc * BK;
                    if (p.x3) {                                 // virtual chunk -> (part, real chunk)
                        const int part = kc / p.kchunks, j = kc - part * p.kchunks;
                        ca = j * BK + (part == 1 ? p.a_lo : 0);
                        cb = j * BK + (part == 2 ? p.b_lo : 0);
                    }
                    if (p.mode == 0)
                        tma_load_3d(dst, &tmA, &bar_full[s], ca, mx0 + ph.ax[t], my0 + ph.ay[t]);
                    else
                        tma_load_5d(dst, &tmA, &bar_full[s], ca, ph.qx[t], mx0 + ph.ax[t], ph.qy[t],
                                    my0 + ph.ay[t]);
                    tma_load_3d(dst + p.a_bytes, &tmB, &bar_full[s], cb, 0, ph.widx[t]);
                    if (++kc == kv) { kc = 0; ++t; }
                }
                if (++s == NSTAGES) { s = 0; par ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            int s = 0;
            uint32_t par = 0;
            for (int done = 0; done < total_it; done += IPS) {
                const int n = min(IPS, total_it - done);
                mbar_wait(&bar_full[s], par);
                tc_fence_after();
                uint32_t a_addr = smem_u32(tiles + (size_t)s * p.stage_bytes);
                for (int g = 0; g < n; ++g, a_addr += p.item_bytes) {
                    const uint64_t adesc = make_desc(a_addr, ROWB);
                    const uint64_t bdesc = make_desc(a_addr + p.a_bytes, ROWB);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk)
                        umma_bf16(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                  (done > 0 || g > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&bar_empty[s]);
                if (++s == NSTAGES) { s = 0; par ^= 1u; }
            }
            umma_commit(&bar_acc);
            if (p.gdn) {
                // norm[128 x N] = xsq[128 x N] . gamma[N x N]^T   (second accumulator at column N)
                mbar_wait(&bar_gamma, 0);
                mbar_wait(&bar_xsq, 0);
                tc_fence_after();
                const int rowb = p.kg * 2;
                const int nch = N / p.kg;                      // chunks of one half (hi or lo) of x^2 / gamma
                for (int part = 0; part < (p.x3 ? 3 : 1); ++part)        // hi.Ghi, lo.Ghi, hi.Glo
                    for (int k16 = 0; k16 < N / 16; ++k16) {
                        const int chunk = (k16 * 16) / p.kg, inner = (k16 * 16) % p.kg;
                        const int ca = chunk + (part == 1 ? nch : 0), cb = chunk + (part == 2 ? nch : 0);
                        const uint64_t ad = make_desc(smem_u32(xsq_tile + (size_t)ca * 128 * rowb) + inner * 2, rowb);
                        const uint64_t bd = make_desc(smem_u32(gam_tile + (size_t)cb * N * rowb) + inner * 2, rowb);
                        umma_bf16(tmem_base + (uint32_t)N, ad, bd, idesc, (part | k16) ? 1u : 0u);
                    }
                umma_commit(&bar_norm);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;               // tile row = output pixel
        const int my = my0 + row / p.tw, mx = mx0 + row % p.tw;
        const int oy = my * p.out_step + ph.out_py, ox = mx * p.out_step + ph.out_px;
        const bool valid = (my < p.mh) && (mx < p.mw) && (oy < p.out.h) && (ox < p.out.w);
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const EpiCtx ctx = make_epi(p.out, p.res, p.gate, p.post, p.act_channels, p.out_scale != nullptr, p.x3 != 0);

        mbar_wait(&bar_acc, 0);
        tc_fence_after();

        if (p.gdn) {
            // pass 1: (acc + bias)^2 -> bf16 A operand in smem (K-major, swizzled like TMA would)
            const int rowb = p.kg * 2;
            const int sw_bits = rowb == 128 ? 3 : (rowb == 64 ? 2 : 1);
            for (int j0 = 0; j0 < N; j0 += 16) {
                float v[16];
                tmem_ld16(tlane + (uint32_t)j0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = v[i] + sbias[j0 + i];
                    v[i] = a * a;
                }
                uint32_t w[8], wl[8];
                split16(v, w, wl);                              // (the lo half is only stored in x3 mode)
                const int chunk = j0 / p.kg, inner = j0 % p.kg;
                uint8_t *cbase = xsq_tile + (size_t)chunk * 128 * rowb;
#pragma unroll
                for (int hsel = 0; hsel < 2; ++hsel) {
                    uint32_t off = (uint32_t)(row * rowb + inner * 2 + hsel * 16);
                    off ^= ((off >> 7) & ((1u << sw_bits) - 1u)) << 4;
                    *reinterpret_cast<uint4 *>(cbase + off) =
                        make_uint4(w[4 * hsel], w[4 * hsel + 1], w[4 * hsel + 2], w[4 * hsel + 3]);
                    if (p.x3)
                        *reinterpret_cast<uint4 *>(cbase + (size_t)(N / p.kg) * 128 * rowb + off) =
                            make_uint4(wl[4 * hsel], wl[4 * hsel + 1], wl[4 * hsel + 2], wl[4 * hsel + 3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&bar_xsq);
            mbar_wait(&bar_norm, 0);
            tc_fence_after();
        }

        if (p.gdn) {
            const bool interior = p.out.pad == 0 || (oy > 0 && oy < p.out.h - 1 && ox > 0 && ox < p.out.w - 1);
            const size_t out_elem = valid ? fm_index(p.out, oy, ox, 0) : 0;
#pragma unroll 1
            for (int j0 = 0; j0 < N; j0 += 16) {
                float v[16], nrm[16];
                tmem_ld16(tlane + (uint32_t)j0, v);
                tmem_ld16(tlane + (uint32_t)(N + j0), nrm);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x = v[i] + sbias[j0 + i];
                    const float t = nrm[i] + sbeta[j0 + i];
                    if (p.x3) {                                 // fp32-faithful mode: as conv3x3_tc_gdn_kernel, MUFU.RSQ +
                        float rs = rsqrtf(t);                   // one Newton step = t^-1/2 to ~1 ulp
                        const float ht = 0.5f * t;
                        rs = fmaf(rs, fmaf(-ht * rs, rs, 0.5f), rs);
                        v[i] = (p.gdn == 1) ? x * rs : x * (t * rs);
                    } else {
                        const float rs = rsqrtf(t);             // MUFU.RSQ: 2^-22 relative, far below bf16
                        v[i] = (p.gdn == 1) ? x * rs : x * (t * rs);
                    }
                }
                if (valid) epi_tail16(v, ctx, sscale, oy, ox, j0, interior, out_elem);
            }
        } else {
            epi_row_dispatch(p.act, tlane, N, sbias, sscale, ctx, oy, ox, valid);
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------ host side
void pick_tile(int mh, int mw, int *tw, int *th) {
    int best = 1 << 30, btw = 16;
    for (int w = 128; w >= 8; w >>= 1) {       // prefer wide tiles on ties
        const int h = 128 / w;
        const int tiles = ceil_div(mw, w) * ceil_div(mh, h);
        if (tiles < best) { best = tiles; btw = w; }
    }
    *tw = btw;
    *th = 128 / btw;
}

template <int BK>
int launch_bk(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &g, const TcParams &p, dim3 grid,
              size_t smem, cudaStream_t st) {
    if (smem_attr_once((const void *)conv_tc_kernel<BK>, 220 * 1024)) return 1;
    AIVC_CHECK_CUDA(launch_pdl(conv_tc_kernel<BK>, grid, dim3(NTHREADS), smem, st, a, b, g, p));
    AIVC_CHECK_LAUNCH("conv_tc_kernel");
    return 0;
}

}  // namespace

int conv_tc3_run(const aivc_conv_op *op, cudaStream_t st);
int conv_tc1_run(const aivc_conv_op *op, cudaStream_t st);
int tconv_tc3_run(const aivc_conv_op *op, cudaStream_t st);

int conv_tc_run(const aivc_conv_op *op, cudaStream_t st) {
    if (op->engine != AIVC_ENGINE_TC_X3) {                   // (plain bf16 only)
        const int r = conv_tc1_run(op, st);                  // persistent 1x1 kernel
        if (r >= 0) return r;
    }
    {
        const int r = tconv_tc3_run(op, st);                 // persistent transposed 3x3 kernel
        if (r >= 0) return r;
    }
    {
        const int r = conv_tc3_run(op, st);                  // persistent 3x3 kernels
        if (r >= 0) return r;
    }
    g_aivc_kernel_class = AIVC_KC_TC_GENERIC;
    const int k = op->k, cin = op->in.c, cout = op->out.c;
    const bool x3 = op->engine == AIVC_ENGINE_TC_X3;
    if (x3 ? op->in.dtype != AIVC_BF16X2 : op->in.dtype != AIVC_BF16)
        AIVC_FAIL("conv_tc: input feature map must be %s", x3 ? "split bf16 (engine TC_X3)" : "bf16");
    if (cin % 16 || cout % 16 || cout > 256) AIVC_FAIL("conv_tc: cin %d / cout %d not tileable", cin, cout);
    if (k * k > MAX_TAPS) AIVC_FAIL("conv_tc: kernel size %d unsupported", k);
    if (op->in.c_off % 8 || op->in.c_stride % 8) AIVC_FAIL("conv_tc: input channel view must be 16-byte aligned");
    const int gdn = op->act == AIVC_ACT_GDN ? 1 : (op->act == AIVC_ACT_IGDN ? 2 : 0);
    if (gdn && !(cout == 16 || cout == 32 || cout == 64 || cout == 128))
        AIVC_FAIL("conv_tc: fused GDN supports 16/32/64/128 channels, got %d", cout);
    if (gdn && !op->bias) AIVC_FAIL("conv_tc: fused GDN expects a bias");
    const int BK = (cin % 64 == 0) ? 64 : ((cin % 32 == 0) ? 32 : 16);
    const int rowb = BK * 2;

    TcParams p;
    memset(&p, 0, sizeof(p));
    p.out = to_dev(op->out);
    if (op->residual.data) p.res = to_dev(op->residual);
    if (op->gate.data) p.gate = to_dev(op->gate);
    p.bias = op->bias; p.gdn_beta = op->gdn_beta; p.out_scale = op->out_scale;
    p.cout = cout; p.kchunks = cin / BK;
    p.act = gdn ? AIVC_ACT_NONE : op->act; p.post = op->post; p.act_channels = op->act_channels; p.gdn = gdn;
    p.x3 = x3 ? 1 : 0; p.a_lo = op->in.c_stride / 2; p.b_lo = cin;
    if (x3 && (p.a_lo % 8)) AIVC_FAIL("conv_tc: split-bf16 input needs 16-byte aligned halves");
    p.kg = cout < 64 ? cout : 64;
    int need_cols = gdn ? 2 * cout : cout;
    p.tmem_cols = 32;
    while (p.tmem_cols < need_cols) p.tmem_cols <<= 1;
    p.a_bytes = 128u * rowb;
    p.b_bytes = (uint32_t)cout * rowb;
    p.item_bytes = p.a_bytes + ((p.b_bytes + 1023u) & ~1023u);
    p.stage_bytes = p.item_bytes * (uint32_t)(64 / BK);

    const aivc_fmap &in = op->in;
    const size_t pix_b = (size_t)in.c_stride * 2, row_b = (size_t)in.pitch * pix_b;
    // channel extent of the activation map: a split-bf16 view also reaches its lo half at + c_stride / 2
    const int cin_ext = x3 ? p.a_lo + cin : cin;
    CUtensorMap tmA, tmB, tmG;
    memset(&tmG, 0, sizeof(tmG));
    int nphase = 1;
    if (op->kind == 0) {
        const int half = k / 2;
        if (k > 1 && in.pad < half) AIVC_FAIL("conv_tc: input border %d < %d", in.pad, half);
        p.mh = op->out.h; p.mw = op->out.w; p.out_step = 1;
        pick_tile(p.mh, p.mw, &p.tw, &p.th);
        Phase &ph = p.ph[0];
        ph.ntaps = k * k;
        void *base = (char *)in.data + (size_t)in.c_off * 2;
        if (op->stride == 1) {
            p.mode = 0;
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) {
                    const int t = ky * k + kx;
                    ph.ax[t] = (signed char)(kx - half + in.pad);
                    ph.ay[t] = (signed char)(ky - half + in.pad);
                    ph.widx[t] = (unsigned char)t;
                }
            cuuint64_t dims[3] = {(cuuint64_t)cin_ext, (cuuint64_t)(in.w + 2 * in.pad), (cuuint64_t)(in.h + 2 * in.pad)};
            cuuint64_t strides[2] = {pix_b, row_b};
            cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)p.tw, (cuuint32_t)p.th};
            if (encode_map(&tmA, base, 3, dims, strides, box, rowb, "A/3d")) return 1;
        } else {
            p.mode = 1;
            if ((in.pitch & 1) || (in.rows & 1)) AIVC_FAIL("conv_tc: stride-2 input needs even pitch/rows");
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) {
                    const int t = ky * k + kx;
                    const int ex = kx - half + in.pad, ey = ky - half + in.pad;
                    ph.ax[t] = (signed char)(ex >> 1); ph.qx[t] = (unsigned char)(ex & 1);
                    ph.ay[t] = (signed char)(ey >> 1); ph.qy[t] = (unsigned char)(ey & 1);
                    ph.widx[t] = (unsigned char)t;
                }
            cuuint64_t dims[5] = {(cuuint64_t)cin_ext, 2, (cuuint64_t)(in.pitch / 2), 2, (cuuint64_t)(in.rows / 2)};
            cuuint64_t strides[4] = {pix_b, 2 * pix_b, row_b, 2 * row_b};
            cuuint32_t box[5] = {(cuuint32_t)BK, 1, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th};
            if (encode_map(&tmA, base, 5, dims, strides, box, rowb, "A/5d")) return 1;
        }
    } else {
        // transposed conv: 4 output phases, zero padding = TMA fill outside the *interior*
        const int pad = (k + 1) / 2 - 1;
        p.mode = 0; p.mh = in.h; p.mw = in.w; p.out_step = 2;
        pick_tile(p.mh, p.mw, &p.tw, &p.th);
        nphase = 4;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                Phase &ph = p.ph[py * 2 + px];
                ph.out_py = py; ph.out_px = px;
                int n = 0;
                for (int ky = 0; ky < k; ++ky) {
                    if ((py + pad - ky) & 1) continue;
                    for (int kx = 0; kx < k; ++kx) {
                        if ((px + pad - kx) & 1) continue;
                        ph.ay[n] = (signed char)((py + pad - ky) / 2);
                        ph.ax[n] = (signed char)((px + pad - kx) / 2);
                        ph.widx[n] = (unsigned char)(ky * k + kx);
                        ++n;
                    }
                }
                ph.ntaps = n;
            }
        void *base = (char *)in.data + ((size_t)in.pad * in.pitch + in.pad) * pix_b + (size_t)in.c_off * 2;
        cuuint64_t dims[3] = {(cuuint64_t)cin_ext, (cuuint64_t)in.w, (cuuint64_t)in.h};
        cuuint64_t strides[2] = {pix_b, row_b};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)p.tw, (cuuint32_t)p.th};
        if (encode_map(&tmA, base, 3, dims, strides, box, rowb, "A/tconv")) return 1;
    }
    p.tiles_x = ceil_div(p.mw, p.tw);
    const int tiles_y = ceil_div(p.mh, p.th);
    {
        const int wcin = x3 ? 2 * cin : cin;                   // x3: [tap][cout][hi cin | lo cin]
        cuuint64_t dims[3] = {(cuuint64_t)wcin, (cuuint64_t)cout, (cuuint64_t)(k * k)};
        cuuint64_t strides[2] = {(cuuint64_t)wcin * 2, (cuuint64_t)wcin * cout * 2};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)cout, 1};
        if (encode_map(&tmB, (void *)op->weight, 3, dims, strides, box, rowb, "B")) return 1;
    }
    // Pipeline depth.  Two CTAs per SM (<= ~110 KB each) let one CTA's epilogue hide behind the
    // other's main loop; grids that cannot fill the chip twice get one deep pipeline instead.
    const int nctas = p.tiles_x * tiles_y * nphase;
    const size_t gam_bytes = gdn ? (size_t)cout * cout * 2 * (x3 ? 2 : 1) : 0;
    const size_t xsq_bytes = gdn ? (size_t)128 * cout * 2 * (x3 ? 2 : 1) : 0;
    (void)nctas;
    const size_t budget = (x3 && gdn) ? 200 * 1024 : 110 * 1024;   // default: two CTAs per SM
    int nst = (int)((budget - 1024 - gam_bytes) / p.stage_bytes);
    if (nst > MAX_STAGES) nst = MAX_STAGES;
    if (BK == 64 && nst > 3 && budget < 150 * 1024) nst = 3;   // measured: deeper 16 KB-stage rings only add L2 pressure
    if (nst < 2) nst = 2;
    p.nstages = nst;
    p.xsq_off = 0;                                   // x^2 tile aliases the stage ring ...
    size_t smem = 1024 + (size_t)nst * p.stage_bytes + gam_bytes;
    if (gdn && (size_t)nst * p.stage_bytes < xsq_bytes) {          // ... unless the ring is too small
        p.xsq_off = nst * (int)p.stage_bytes;
        smem += xsq_bytes;
    }
    if (gdn) {
        const int gk = x3 ? 2 * cout : cout;                   // x3: gamma rows are [hi | lo]
        cuuint64_t dims[2] = {(cuuint64_t)gk, (cuuint64_t)cout};
        cuuint64_t strides[1] = {(cuuint64_t)gk * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.kg, (cuuint32_t)cout};
        if (encode_map(&tmG, (void *)op->gdn_gamma, 2, dims, strides, box, p.kg * 2, "gamma")) return 1;
    }
    if (smem > 220 * 1024) AIVC_FAIL("conv_tc: %zu bytes of shared memory needed", smem);
    dim3 grid(p.tiles_x * tiles_y, 1, nphase);
    if (BK == 64) return launch_bk<64>(tmA, tmB, tmG, p, grid, smem, st);
    if (BK == 32) return launch_bk<32>(tmA, tmB, tmG, p, grid, smem, st);
    return launch_bk<16>(tmA, tmB, tmG, p, grid, smem, st);
}
