// Persistent tcgen05 kernel for the layer that carries most of AIVC's FLOPs: 3x3, stride 1,
// replicate-padded convolution with Cin a multiple of 64 (ChengResBlock / ResBlock /
// attention trunks; custom_conv_layers.py:40-56, 112-126).
//
// The generic kernel (conv_tc.cu) reloads a 128-pixel A tile for each of the 9 taps and streams
// the whole 9*Cin x Cout weight matrix per 128 pixels: it is bound by L2->SM bandwidth.  Here
//   * a CTA owns a 32 x 8 pixel tile = two 128-row accumulators, so every weight slice that
//     reaches shared memory feeds twice as many MMAs;
//   * ALL nine taps share ONE activation box per 64-channel chunk: a TMA box of (32+2) rows x
//     (8+2) pixels lands as 340 rows of 128 B, and tap (ky,kx) of sub-tile j is the UMMA
//     descriptor start address + ((ky + 16 j) * 10 + kx) rows with a stride of 10 rows between
//     8-row groups.  (Measured on B200: the 128B swizzle is applied to the absolute shared-memory
//     address, so a 128 B-aligned, non-1024 B-aligned start needs no descriptor base offset.)
//     Activation L2 traffic drops 6x against per-tap loads;
//   * barriers are per kernel ROW (3 taps = 24 MMAs): the single MMA-issuing thread spends ~450
//     cycles per barrier round trip, which per-tap barriers (8 MMAs = 512 cycles) could not hide;
//   * the CTA is persistent: accumulators are double-buffered in TMEM (2 x 256 columns), the
//     eight epilogue warps drain tile i while the MMA thread works on tile i + 1.
// Per 256 pixels: A 2 x 43 KB + B 18 x 16 KB = 375 KB of L2 reads (generic kernel: 1152 KB).
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tcgen;

namespace {

constexpr int NTHREADS = 576;       // TMA warp, MMA warp, 16 epilogue warps: 4 teams = (sub-tile j, channel half h) x 4 lane quarters
// The warp scheduler favours HIGHER warp ids, so the two latency-critical single-thread roles get the
// highest ids and are never starved by the ALU-heavy epilogue warps of their sub-partition.
constexpr int TMA_WARP = 16, MMA_WARP = 17;
constexpr int STAGE_BYTES = 128 * 64;   // one team's store staging tile: 128 rows x 32 channels bf16
constexpr int TILE_H = 32, TILE_W = 8;
constexpr int PATCH_W = TILE_W + 2;                                             // 10 pixels
constexpr uint32_t PATCH_BYTES = (TILE_H + 2) * PATCH_W * 128;                   // 43520 B per 64-channel chunk
constexpr uint32_t A_SLOT = (PATCH_BYTES + 1023) & ~1023u;                       // 44032 B
constexpr int NA = 2;               // activation patches in flight
constexpr int NG_MAX = 4;           // weight groups (3 taps = one kernel row) in flight

struct Tc3Params {
    FMap out, res, gate;
    const float *bias, *out_scale;
    int cout, kchunks;
    int act, post, act_channels;
    int tiles_x, ntiles;
    int in_pad;
    uint32_t b_bytes, b_slot;         // one tap's weight slice (cout x 64 ch) and its 1 KB-rounded slot
    int ng;                           // weight-group ring depth
    int tma_store;                    // epilogue stores through smem + TMA (bf16, 64-channel multiples)
    int dbg;                          // AIVC_TC3_DBG bits: timing experiments only (wrong results)
};

template <int ACT>
__device__ __forceinline__ void epilogue_team(const Tc3Params &p, const CUtensorMap *tmO, uint8_t *stage,
                                              const float *sbias, const float *sscale, uint64_t *acc_full,
                                              uint64_t *acc_empty, uint32_t tmem_base, int warp, int lane) {
    const int quarter = warp & 3, team = warp >> 2, j = team >> 1, h = team & 1;
    const int row = quarter * 32 + lane;
    const bool leader = ((warp & 3) == 0) && lane == 0;
    const int N = p.cout;
    const int c_lo = h * 64, c_hi = min(N, c_lo + 64);
    EpiCtx ctx = make_epi(p.out, p.res, p.gate, p.post, p.act_channels, p.out_scale != nullptr);
    ctx.dbg = p.dbg;
    uint8_t *my_stage = stage + team * STAGE_BYTES;
    uint32_t it = 0;
    bool store_pending = false;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        const int y0 = (tile / p.tiles_x) * TILE_H, x0 = (tile % p.tiles_x) * TILE_W;
        const int oy = y0 + 16 * j + row / TILE_W, ox = x0 + row % TILE_W;
        const bool valid = (oy < p.out.h) && (ox < p.out.w);
        mbar_wait(&acc_full[buf], (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256u + (uint32_t)(j * 128);
        if (c_lo < N && !(p.dbg & 2)) {
            if (p.tma_store) {
                const bool interior = ctx.out.pad == 0 || (oy > 0 && oy < p.out.h - 1 && ox > 0 && ox < p.out.w - 1);
                for (int ch0 = c_lo; ch0 < c_hi; ch0 += 32) {
                    if (store_pending) {                       // staging tile still being read by the last store?
                        if (leader) tma_store_wait_read();
                        named_bar_sync(1 + team, 128);
                    }
                    epi_row_staged32<ACT>(tl, ch0, sbias, sscale, ctx, oy, ox, valid, interior, my_stage, row);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    named_bar_sync(1 + team, 128);
                    if (leader) {
                        tma_store_3d(tmO, my_stage, ch0, x0, y0 + 16 * j);
                        tma_store_commit();
                    }
                    store_pending = true;
                }
            } else {
                epi_row<ACT>(tl, c_hi, sbias, sscale, ctx, oy, ox, valid, c_lo);
            }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
    }
    if (store_pending && leader) tma_store_wait_read();        // smem must outlive the last store's read
}

__global__ void __launch_bounds__(NTHREADS, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmO,
                                                                 const Tc3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[NA], a_empty[NA], g_full[NG_MAX], g_empty[NG_MAX], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[128], sscale[128];

    uint8_t *a_ring = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *g_ring = a_ring + NA * A_SLOT;
    const uint32_t g_slot = 3u * p.b_slot;
    uint8_t *stage = g_ring + (size_t)p.ng * g_slot;          // 4 teams x 8 KB

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.cout;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NG_MAX; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 512); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    stage_vec(sbias, p.bias, N, 0.f, tid, NTHREADS);          // weights/bias never depend on the prior grid
    stage_vec(sscale, p.out_scale, N, 1.f, tid, NTHREADS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait_prior_grid();                                    // activations / residuals below do

    if (warp == TMA_WARP) {
        // ===================== TMA producer =====================
        // per tile and 64-channel chunk: ONE activation patch (34 x 10 pixels) that serves all nine
        // taps, and three weight groups (one kernel row = 3 taps each), every group on one barrier
        if (lane == 0) {
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0;          // ring slot / phase, advanced incrementally
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int y0 = (tile / p.tiles_x) * TILE_H, x0 = (tile % p.tiles_x) * TILE_W;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&a_empty[sa], pa ^ 1u);
                    if (p.dbg & 8) mbar_arrive(&a_full[sa]);
                    else {
                        mbar_expect_tx(&a_full[sa], PATCH_BYTES);
                        tma_load_3d(a_ring + sa * A_SLOT, &tmA, &a_full[sa], kc * 64, x0 - 1 + p.in_pad,
                                    y0 - 1 + p.in_pad);
                    }
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                    for (int ky = 0; ky < 3; ++ky) {
                        mbar_wait(&g_empty[sg], pg ^ 1u);
                        if (p.dbg & 16) mbar_arrive(&g_full[sg]);
                        else {
                            mbar_expect_tx(&g_full[sg], 3u * p.b_bytes);
                            uint8_t *dst = g_ring + sg * g_slot;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx)
                                tma_load_3d(dst + kx * p.b_slot, &tmB, &g_full[sg], kc * 64, 0, ky * 3 + kx);
                        }
                        if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer =====================
        // 24 MMAs (3 taps x 2 sub-tiles x 4 k-steps) per barrier wait: the issuing thread's
        // bookkeeping must stay well below the 1536 tensor cycles they take
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            // descriptor template: K-major, 128B swizzle, 8-row groups PATCH_W rows apart (the
            // swizzle is a function of the absolute smem address, so any 128 B-aligned start works)
            const uint64_t a_tmpl = (1ull << 16) | ((uint64_t)((PATCH_W * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1u;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);     // epilogue drained this buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * 256u;
                uint32_t accum = 0;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&a_full[sa], pa);
                    const uint32_t a_addr = smem_u32(a_ring + sa * A_SLOT);
                    for (int ky = 0; ky < 3; ++ky) {
                        mbar_wait(&g_full[sg], pg);
                        tc_fence_after();
                        const uint32_t g_addr = smem_u32(g_ring + sg * g_slot);
                        if (!(p.dbg & 4)) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint64_t bdesc = make_desc(g_addr + kx * p.b_slot, 128);
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint32_t start = a_addr + (uint32_t)((ky + 16 * j) * PATCH_W + kx) * 128u;
                                    const uint64_t adesc = a_tmpl | (uint64_t)((start >> 4) & 0x3FFF);
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk) {
                                        umma_bf16(acc + (uint32_t)(j * 128), adesc + (uint64_t)(kk * 2),
                                                  bdesc + (uint64_t)(kk * 2), idesc, accum | (uint32_t)(kx | kk));
                                    }
                                }
                            }
                        }
                        accum = 1;
                        umma_commit(&g_empty[sg]);
                        if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                    }
                    umma_commit(&a_empty[sa]);
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===================== epilogue (warps 0..15) =====================
        switch (p.act) {
            case AIVC_ACT_LEAKY: epilogue_team<AIVC_ACT_LEAKY>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane); break;
            case AIVC_ACT_RELU: epilogue_team<AIVC_ACT_RELU>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane); break;
            case AIVC_ACT_SIGMOID: epilogue_team<AIVC_ACT_SIGMOID>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane); break;
            default: epilogue_team<AIVC_ACT_NONE>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane); break;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                     : "memory");
    }
}

}  // namespace

// Returns -1 when the stage does not fit this kernel (caller falls through to the generic one).
int conv_tc3_run(const aivc_conv_op *op, cudaStream_t st) {
    const int cin = op->in.c, cout = op->out.c;
    if (op->kind != 0 || op->k != 3 || op->stride != 1) return -1;
    if (cin % 64 || cout % 16 || cout > 128) return -1;
    if (op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN) return -1;
    if (op->in.dtype != AIVC_BF16 || op->in.pad < 1 || op->in.c_off % 8 || op->in.c_stride % 8) return -1;
    const int tiles_x = ceil_div(op->out.w, TILE_W), tiles_y = ceil_div(op->out.h, TILE_H);
    const int ntiles = tiles_x * tiles_y;
    if (ntiles < 296) return -1;            // < 2 tiles per SM: the 128-pixel-tile kernel fills the chip better

    Tc3Params p;
    memset(&p, 0, sizeof(p));
    p.out = to_dev(op->out);
    if (op->residual.data) p.res = to_dev(op->residual);
    if (op->gate.data) p.gate = to_dev(op->gate);
    p.bias = op->bias; p.out_scale = op->out_scale;
    p.cout = cout; p.kchunks = cin / 64;
    p.act = op->act; p.post = op->post; p.act_channels = op->act_channels;
    p.tiles_x = tiles_x; p.ntiles = ntiles; p.in_pad = op->in.pad;
    p.b_bytes = (uint32_t)cout * 128u;
    p.b_slot = (p.b_bytes + 1023u) & ~1023u;
    { const char *e = getenv("AIVC_TC3_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.ng = NG_MAX;
    while (1024 + (size_t)NA * A_SLOT + (size_t)p.ng * 3 * p.b_slot + 4 * STAGE_BYTES > 219 * 1024 && p.ng > 1) --p.ng;
    if (p.ng < 2) return -1;

    const aivc_fmap &in = op->in;
    const size_t pix_b = (size_t)in.c_stride * 2, row_b = (size_t)in.pitch * pix_b;
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)(in.w + 2 * in.pad), (cuuint64_t)(in.h + 2 * in.pad)};
        cuuint64_t strides[2] = {pix_b, row_b};
        cuuint32_t box[3] = {64, PATCH_W, TILE_H + 2};
        if (encode_map(&tmA, (char *)in.data + (size_t)in.c_off * 2, 3, dims, strides, box, 128, "A/3x3")) return 1;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)cout, 1};
        if (encode_map(&tmB, (void *)op->weight, 3, dims, strides, box, 128, "B/3x3")) return 1;
    }
    CUtensorMap tmO;
    memset(&tmO, 0, sizeof(tmO));
    const aivc_fmap &o = op->out;
    p.tma_store = (o.dtype == AIVC_BF16 && o.c_off % 8 == 0 && o.c_stride % 8 == 0 && cout % 32 == 0 &&
                   getenv("AIVC_TC3_NO_TMA_STORE") == nullptr) ? 1 : 0;
    if (p.tma_store) {
        const size_t opix = (size_t)o.c_stride * 2, orow = (size_t)o.pitch * opix;
        cuuint64_t dims[3] = {(cuuint64_t)cout, (cuuint64_t)o.w, (cuuint64_t)o.h};     // interior only: OOB rows/cols are clipped
        cuuint64_t strides[2] = {opix, orow};
        cuuint32_t box[3] = {32, TILE_W, 16};
        void *base = (char *)o.data + ((size_t)o.pad * o.pitch + o.pad) * opix + (size_t)o.c_off * 2;
        if (encode_map(&tmO, base, 3, dims, strides, box, 64, "O/3x3")) return 1;
    }
    const size_t smem = 1024 + (size_t)NA * A_SLOT + (size_t)p.ng * 3 * p.b_slot + 4 * STAGE_BYTES;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        AIVC_CHECK_CUDA(cudaGetDevice(&dev));
        AIVC_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    AIVC_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         220 * 1024));
    const int grid = ntiles < sm_count ? ntiles : sm_count;
    AIVC_CHECK_CUDA(launch_pdl(conv3x3_tc_kernel, dim3(grid), dim3(NTHREADS), smem, st, tmA, tmB, tmO, p));
    AIVC_CHECK_LAUNCH("conv3x3_tc_kernel");
    return 0;
}
