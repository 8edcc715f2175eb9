/* aivc_b200 -- C ABI of the B200-native AIVC hot path (libaivc_b200.so).
 *
 * The reference (Orange-OpenSource/AIVC) has no FFI: its hot path is Python classes
 * calling torch.nn.functional.  This header is the boundary a maintainer binds with
 * ctypes; each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes, no torch types; device pointers unless noted "host".
 *   - every device entry point is asynchronous on `stream` (a cudaStream_t passed as
 *     void*), performs no allocation and no hidden synchronisation; the caller owns all
 *     buffers, including outputs and scratch.
 *   - return value 0 = ok, non-zero = error; aivc_last_error() gives the (thread-local)
 *     message.  Nothing throws or exits across the ABI.
 *   - feature maps are NHWC ("pixel-major") with an optional replicate-filled border so
 *     that 3x3/5x5 taps of the next convolution are plain in-bounds TMA box loads.
 */
#ifndef AIVC_B200_H
#define AIVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIVC_ABI_VERSION 1

/* ---- data types / enums ------------------------------------------------------------ */
enum { AIVC_F32 = 0, AIVC_BF16 = 1,
       AIVC_F16 = 2, /* only the GEMM form of the output transposed conv (kind 2 input): pixel-domain partial
                      * sums need the 11-bit mantissa */
       AIVC_BF16X2 = 3 /* split bf16 (precision mode bf16x3): a pixel holds c_stride bf16 elements, element
                        * c_off + ch is the leading 8 bits ("hi") of channel ch and element
                        * c_stride/2 + c_off + ch the next 8 ("lo" = bf16(v - hi)); value = hi + lo.
                        * A tensor-core stage reads both halves as K-major operands and issues
                        * hi.Whi + lo.Whi + hi.Wlo, which restores fp32-grade products (2^-17) on tcgen05. */ };

/* activation applied right after bias (custom_conv_layers.py:155-177, attention.py:82) */
enum {
    AIVC_ACT_NONE = 0,
    AIVC_ACT_LEAKY = 1,    /* LeakyReLU(0.01) */
    AIVC_ACT_RELU = 2,
    AIVC_ACT_SIGMOID = 3,
    AIVC_ACT_GDN = 4,      /* x * (beta + gamma.x^2)^-1/2   misc_layers.py:113-154 */
    AIVC_ACT_IGDN = 5      /* x * (beta + gamma.x^2)^+1/2 */
};

/* applied last, after gate/residual */
enum {
    AIVC_POST_NONE = 0,
    AIVC_POST_LEAKY = 1,       /* AttentionResBlock: leaky(x + f(x))  attention.py:41 */
    AIVC_POST_RELU = 2,        /* ResBlock: relu(x + f(x))            custom_conv_layers.py:126 */
    AIVC_POST_ROUND_CLAMP = 3  /* Quantizer at eval + AC range clamp  misc_layers.py:167 */
};

enum { AIVC_ENGINE_SIMT = 0, AIVC_ENGINE_TC = 1,
       AIVC_ENGINE_TC_X3 = 2 /* tensor cores on split-bf16 operands (AIVC_BF16X2 input, weights packed
                              * [tap][cout][hi cin | lo cin]): hi.Whi + lo.Whi + hi.Wlo per K chunk, fp32
                              * accumulation in TMEM -- fp32-grade results (the "bf16x3" precision mode) */ };

/* A view on an NHWC feature map living in a (possibly wider, possibly bordered) buffer.
 * element (y, x, ch) of the view is at
 *   data[ ((y + pad) * pitch + (x + pad)) * c_stride + c_off + ch ]
 * `pad` border pixels on every side hold replicated edge values (written by producers). */
typedef struct {
    void *data;
    int32_t h, w;        /* logical size */
    int32_t c;           /* channels in this view */
    int32_t c_off;       /* first channel of the view inside a pixel */
    int32_t c_stride;    /* channels allocated per pixel */
    int32_t pad;         /* border width */
    int32_t pitch;       /* allocated pixels per row  (>= w + 2*pad) */
    int32_t rows;        /* allocated rows            (>= h + 2*pad) */
    int32_t dtype;       /* AIVC_F32 | AIVC_BF16 (| AIVC_F16, see above) */
    int32_t _r;
} aivc_fmap;

/* One fused convolution stage.
 *   kind 0: replicate-pad(k/2) + Conv2d(k, stride)      CustomConvLayer  custom_conv_layers.py:129-180
 *           (also the bare 1x1 Conv2d's of ChengResBlock/attention: pad 0)
 *   kind 1: ConvTranspose2d(k, stride 2, pad (k+1)/2-1, output_padding 1)
 *                                                        UpscalingLayer   custom_conv_layers.py:183-253
 *   kind 2: col2im of a transposed conv computed as GEMM: in holds k*k*cout channels
 *           P[(ky,kx,co)] per INPUT pixel, out[2iy-pad+ky][2ix-pad+kx][co] = act(bias + sum P)
 *   kind 3: space-to-depth repack (no arithmetic): out is the half-resolution map whose pixel (by, bx)
 *           holds the 2x2 block of `in` pixels as channels (dy*2+dx)*in.c + c, pixels outside `in`
 *           replicated from its edge; out.h = ceil(in.h/2), out.c = 4*in.c, both bf16.  It turns the 5x5
 *           stride-2 pixel-domain conv (first layer of g_a / g_a_ref) into a 3x3 stride-1 conv over 64
 *           channels that the persistent tensor-core kernel runs.
 *   out = post( act(conv(in) + bias) * gate + residual ) * out_scale
 */
typedef struct {
    int32_t kind, k, stride, engine;
    aivc_fmap in, out;
    const void *weight;        /* packed by aivc_pack_conv_weight for the chosen engine */
    const float *bias;         /* [cout] or NULL */
    int32_t act, post;
    const float *gdn_beta;     /* [cout], reparametrised */
    const void *gdn_gamma;     /* [cout][cout] fp32 (SIMT) or bf16 K-major (TC) */
    aivc_fmap residual;        /* .data == NULL -> none */
    aivc_fmap gate;            /* .data == NULL -> none */
    const float *out_scale;    /* [cout] or NULL: GainMatrix 'enc' folded in (gain_matrix.py:122-124) */
    void *scratch;             /* SIMT GDN needs cout*h*w floats; else NULL */
    int32_t act_channels;      /* `act` applies to output channels [0, act_channels); 0 = all */
    int32_t flags;             /* AIVC_OP_* : two-lane execution inside aivc_conv2d_fused_seq */
    double alg_flops;          /* algorithmic FLOPs of the REFERENCE op(s) this stage stands for (roofline
                                * accounting); 0 = derive from this stage's own geometry */
} aivc_conv_op;

/* Independent branches of a block (SimplifiedAttention's trunk and attention paths) run on two
 * CUDA streams: lane 0 = the caller's stream, lane 1 = an internal side stream. */
enum {
    AIVC_OP_LANE1 = 1,   /* run this stage on the side stream */
    AIVC_OP_FORK = 2,    /* before it: side stream waits for everything queued on the caller's stream */
    AIVC_OP_JOIN = 4,    /* before it: caller's stream waits for everything queued on the side stream */
    AIVC_OP_IN_EXACT = 8 /* split-bf16 input whose lo halves are all zero (8-bit level units straight from the pixel
                          * buffers): the lo.Whi third of the MMAs is skipped (persistent 3x3 kernels) */
};

/* ---- library ----------------------------------------------------------------------- */
int aivc_abi_version(void);
const char *aivc_last_error(void);

/* ---- instrumentation (bench.py): kernels launched so far by this process, and optional
 * CUDA-event timing of every convolution stage, summed per engine.
 * aivc_profile_read: out[0..5] = tc_ms, tc_flops, tc_stages, simt_ms, simt_flops, simt_stages */
unsigned long long aivc_launch_count(void);
int aivc_profile_enable(int on);
int aivc_profile_read(double *out);
/* one CSV line per recorded stage (engine, geometry, flops, ms, kernel class) */
int aivc_profile_dump(const char *path);
/* kernel classes of the recorded stages (which kernel a stage ran on) */
enum {
    AIVC_KC_SIMT = 0,        /* conv_simt_kernel (exact fp32) */
    AIVC_KC_TC_GENERIC = 1,  /* conv_tc_kernel: any k / stride / transposed phases, 128-pixel tiles */
    AIVC_KC_TC3 = 2,         /* conv3x3_tc_kernel: persistent 3x3 stride-1 (32x8 or 16x8 pixel tiles) */
    AIVC_KC_TC3_GDN = 3,     /* conv3x3_tc_gdn_kernel: persistent 3x3 + GDN / IGDN */
    AIVC_KC_TC1 = 4,         /* conv1x1_tc_kernel: persistent 1x1 with TMA-staged gate / residual */
    AIVC_KC_COL2IM = 5,      /* col2im_tconv_kernel */
    AIVC_KC_S2D = 6,         /* space_to_depth_kernel */
    AIVC_KC_TCONV3 = 7,      /* tconv3x3_tc_kernel: persistent transposed 3x3 stride 2 */
    AIVC_KC_COUNT = 8
};
/* per kernel class k < n: out[3k] = ms, out[3k+1] = algorithmic flops, out[3k+2] = stages */
int aivc_profile_read_classes(double *out, int n);
/* Host-only introspection (tests): the work items the persistent 3x3 kernel processes for a split-bf16 h x w layer
 * on sm_count SMs -- whole 32x8-pixel tiles for the full waves, 16-row half tiles for the rest -- as (y0, x0, rows)
 * triples in item order (at most cap of them are written); returns the number of items.  No reference counterpart. */
int aivc_debug_tc3_tiling(int h, int w, int sm_count, int *tiles, int cap);

/* ---- convolution stack ------------------------------------------------------------- */
/* Re-layout a PyTorch weight for an engine.  src: Conv2d [cout][cin][k][k] or
 * ConvTranspose2d [cin][cout][k][k] fp32 (device).  dst: SIMT [k*k][cin_pad][cout_pad] fp32,
 * TC [k*k][cout_pad][cin_pad] bf16.  Weight input channel ci lands at buffer channel
 * ci + cin_off (the layer reads a slice of a wider, zero-padded pixel); everything else is 0.
 * `scale` multiplies every weight (1/255 when the input buffer holds 8-bit levels). */
int aivc_pack_conv_weight(const float *src, void *dst, int kind, int k, int cin, int cout,
                          int engine, int cin_pad, int cout_pad, int cin_off, float scale,
                          void *stream);
size_t aivc_packed_weight_bytes(int k, int engine, int cin_pad, int cout_pad);

int aivc_conv2d_fused(const aivc_conv_op *op, void *stream);
/* A transform as ONE CUDA graph: aivc_plan_graph_create captures the n stages exactly as aivc_conv2d_fused_seq would
 * enqueue them (programmatic dependent launches, fork / join of the two-lane attention branches) on an internal stream
 * -- nothing executes -- and instantiates the graph; aivc_plan_graph_launch replays it on `stream` (one launch call,
 * no tensor-map encodes, no host gaps between the kernels).  The ops' pointers are baked in: re-create the graph when
 * a buffer or parameter pointer changes.  Run the stages once through aivc_conv2d_fused_seq before capturing (one-time
 * function attributes are set on first use).  Kernel launches inside a graph are not counted by aivc_launch_count. */
int aivc_plan_graph_create(const aivc_conv_op *ops, int n, void **graph_exec);
int aivc_plan_graph_launch(void *graph_exec, void *stream);
int aivc_plan_graph_destroy(void *graph_exec);
/* run n stages back to back on one stream (one FFI crossing per transform) */
int aivc_conv2d_fused_seq(const aivc_conv_op *ops, int n, void *stream);

/* ---- layout bridges at the nn.Module boundary (NCHW fp32 <-> bordered NHWC) --------- */
int aivc_nchw_to_fmap(const float *src, const aivc_fmap *dst, void *stream);
int aivc_fmap_to_nchw(const aivc_fmap *src, float *dst, void *stream);
int aivc_fill_border(const aivc_fmap *m, void *stream);
/* dst = src with dst's dtype, channel placement and replicate border (same h, w, c) */
int aivc_fmap_copy(const aivc_fmap *src, const aivc_fmap *dst, void *stream);

/* ---- pixel ends -------------------------------------------------------------------- */
/* InputLayer (ae_layers.py:27-35): planar 4:2:0 -> 3 channels of `dst` (Y, nearest-x2 U, V).
 * planes are uint8 levels when u8 != 0, else fp32 in [0,1].  levels != 0 stores 8-bit level
 * units (0..255, exact in bf16; the consumer's weights carry the 1/255), else [0,1] units. */
int aivc_yuv420_to_fmap(const void *y, const void *u, const void *v, int u8, int levels,
                        const aivc_fmap *dst, void *stream);

/* The same for the tensor-core engines' 16-channel level-unit pixel buffers (bf16, or split bf16 with zero lo halves), fused: up to three uint8 4:2:0 frames
 * (frame to code, previous and next reference; a NULL luma pointer = the all-zero frame, decode.py:710-714)
 * -> channels 0..8 of `dst` (whole pixel, border replicas included, channels 9..15 zero) in one launch;
 * `dst2` (optional) receives frame 0 alone in channels 0..2 (CodecNet input [code | pred]). */
int aivc_yuv420_pack16(const void *y0, const void *u0, const void *v0, const void *y1, const void *u1,
                       const void *v1, const void *y2, const void *u2, const void *v2, const aivc_fmap *dst,
                       const aivc_fmap *dst2, void *stream);

/* MOFNetDecoder post-processing + motion compensation + alpha split
 * (decode.py:729-739, 524-536; optical_flow.py:14-55).  mof: 6 channels (alpha, beta,
 * v_prev xy, v_next xy).  frame_is_p: beta := 1, v_next := 0.
 * pred = alpha * x_warp (3 ch, same units as prev/next), skip = (1 - alpha) * x_warp (3 ch,
 * always [0,1] units; levels != 0 says prev/next hold 8-bit level units).
 * aux: NULL, or fp32 [5][h][w] = alpha, beta, x_warp (3 planes, [0,1] units): the tensors the (missing)
 * encoder returns as net_out['alpha' | 'beta' | 'warping'] for logging (loss_function.py:179-204). */
int aivc_warp_blend(const aivc_fmap *mof, const aivc_fmap *prev, const aivc_fmap *next,
                    int frame_is_p, int levels, const aivc_fmap *pred, const aivc_fmap *skip,
                    float *aux, void *stream);
/* stand-alone motion compensation for the drop-in MotionCompensation module:
 * all tensors NCHW fp32, beta [3][h][w], flows [2][h][w]. */
int aivc_warp_blend_nchw(const float *prev, const float *next, const float *v_prev,
                         const float *v_next, const float *beta, float *out, int h, int w,
                         void *stream);

/* x = codec (+ skip); OutputLayer (ae_layers.py:42-56) + crop / U,V replicate-pad
 * (decode.py:557-571) + cast_before_png_saving (img_processing.py:68-73).
 * Writes uint8 planes y[h][w], u,v[ceil(h/2)][ceil(w/2)] and (optionally) the 4:4:4
 * fp32 reference feature map the next frames read (InputLayer of the result). */
int aivc_finalize_frame(const aivc_fmap *codec, const aivc_fmap *skip, uint8_t *y, uint8_t *u,
                        uint8_t *v, const aivc_fmap *ref444, void *stream);

/* ---- hyperprior / quantisation ----------------------------------------------------- */
/* PdfParamParameterizer (misc_layers.py:180-269): hs [2C] -> mu, sigma NCHW fp32. */
int aivc_mu_sigma_nchw(const float *hs, float *mu, float *sigma, int c, int hw, void *stream);

/* Encoder side, one latent (replaces PdfParamParameterizer + Quantizer + get_y_cdf +
 * torchac's float->int16 conversion; misc_layers.py:167,180-269, bitstream.py:127-154,241-255).
 *   y      : C channels, already multiplied by the encoder gain
 *   hs     : 2C channels (mu | log-variance), at least h x w
 *   q      : int16 NCHW [C][h][w]            = clamp(round(y - mu), -256, 255)
 *   bounds : uint32 NCHW, c_low | c_high<<16 = 16-bit CDF bounds of q under Laplace(0, sigma/sqrt 2)
 *   nz     : int32 [C], set to 1 where the channel has a non-zero symbol (zero it first)
 *   yhat   : C channels = (q + mu) * dec_gain   (input of g_s; decode.py:867-885)
 *   rate   : NULL, or fp32 [C][h][w]: the encoder-side rate ESTIMATE of every symbol in bits,
 *            -log2 clamp(cdf(q + 1/2) - cdf(q - 1/2), 2^-16, 1) for the Laplace of scale sigma/sqrt(2)
 *            (ParametricPdf.forward with zero_mu, pdf_estimator.py:27-70; EntropyCoder.forward,
 *            entropy_coder.py:25-30) -- what the reference logs as `*_rate_y` (loss_function.py:158,186)
 */
int aivc_quantize_latent(const aivc_fmap *y, const aivc_fmap *hs, const float *dec_gain,
                         int16_t *q, uint32_t *bounds, int32_t *nz, const aivc_fmap *yhat,
                         float *rate, void *stream);
/* ParametricPdf.forward (pdf_estimator.py:27-70), one mixture component, contiguous fp32 tensors of n elements:
 * out (+)= cdf(y + 1/2) - cdf(y - 1/2) for Laplace(mu, sigma/sqrt(2)) (family 0) or Normal(mu, sigma) (family 1);
 * mu == NULL: zero mean (zero_mu / the '*_mu' families); accumulate != 0 adds to `out` (mixtures). */
int aivc_pdf_prob(const float *y, const float *mu, const float *sigma, int family, int accumulate, float *out,
                  size_t n, void *stream);
/* Decoder side: scale b = sigma/sqrt(2) per symbol, fp32 NCHW, for the host range decoder. */
int aivc_laplace_scale(const aivc_fmap *hs, int c, float *b, void *stream);
/* Same, plus win[8] per symbol (uint16, NCHW symbol order, 16-byte aligned): the integer CDF entries
 * 253..260 = bounds of the symbols -3..+3, evaluated with the arithmetic of aivc_laplace_cdf_int_host,
 * so the host decoder (aivc_rc_decode_laplace_win) searches a 16-byte table for almost every symbol
 * instead of evaluating the Laplace CDF (bitstream.py:127-184 builds 514 entries per symbol). */
int aivc_laplace_window(const aivc_fmap *hs, int c, float *b, uint16_t *win, void *stream);
/* Decoder side: yhat = (q + mu) * dec_gain from host-decoded symbols. */
int aivc_dequantize_latent(const int16_t *q, const aivc_fmap *hs, const float *dec_gain,
                           const aivc_fmap *yhat, void *stream);
/* z symbols <-> feature map (values already rounded by the h_a epilogue). */
int aivc_fmap_to_i16(const aivc_fmap *src, int16_t *dst, void *stream);
int aivc_i16_to_fmap(const int16_t *src, const aivc_fmap *dst, void *stream);

/* ---- encoder-side metrics --------------------------------------- */
/* MSE / PSNR over Y, U, V (loss_function.py:415-435, 234) and plane-size-weighted MS-SSIM (loss_function.py:
 * 438-470; func_util/ms_ssim.py:37-150: five scales, 11x11 Gaussian window, val_range 1) of two frames given
 * as uint8 4:2:0 planes on the device.  out (device): mse, psnr, ms_ssim, ms_ssim_db.  Deterministic
 * (fixed-order reductions); agrees with the reference's fp32 evaluation to ~1e-6, not bit for bit. */
size_t aivc_frame_metrics_scratch_bytes(int h, int w);
int aivc_frame_metrics(const uint8_t *ya, const uint8_t *ua, const uint8_t *va, const uint8_t *yb,
                       const uint8_t *ub, const uint8_t *vb, int h, int w, void *scratch,
                       size_t scratch_bytes, float *out, void *stream);

/* ---- range coder (host; replaces torchac as called at bitstream.py:281,454,482) ------ */
/* All pointers are HOST memory.  Output capacity must be >= aivc_rc_bound(n). */
size_t aivc_rc_bound(size_t n_symbols);
/* symbols given by their 16-bit CDF bounds (c_low | c_high << 16), coded in array order */
int aivc_rc_encode_bounds(const uint32_t *bounds, size_t n, uint8_t *out, size_t cap, size_t *out_len);
/* 'pmf' mode: per-channel table [c][514] uint16, symbols int16 in [-256, 255], NCHW order */
int aivc_rc_encode_table(const uint16_t *table, const int16_t *sym, int c, size_t hw, uint8_t *out,
                         size_t cap, size_t *out_len);
int aivc_rc_decode_table(const uint16_t *table, const uint8_t *in, size_t in_len, int c, size_t hw,
                         int16_t *sym);
/* 'laplace' mode: per-symbol scale b (from aivc_laplace_scale), symbols out in [-256, 255] */
int aivc_rc_decode_laplace(const float *b, const uint8_t *in, size_t in_len, size_t n, int16_t *sym);
int aivc_rc_decode_laplace_win(const float *b, const uint16_t *win, const uint8_t *in, size_t in_len,
                               size_t n, int16_t *sym);
/* host evaluation of the integer Laplace CDF (same arithmetic as the device) */
uint32_t aivc_laplace_cdf_int_host(float b, int i);
float aivc_sigma_from_logvar_host(float v);

#ifdef __cplusplus
}
#endif
#endif /* AIVC_B200_H */
