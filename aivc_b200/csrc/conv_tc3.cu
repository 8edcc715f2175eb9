// Persistent tcgen05 kernel for the layer that carries most of AIVC's FLOPs: 3x3, stride 1,
// replicate-padded convolution with Cin a multiple of 64 (ChengResBlock / ResBlock /
// attention trunks; custom_conv_layers.py:40-56, 112-126).
//
// The generic kernel (conv_tc.cu) reloads a 128-pixel A tile for each of the 9 taps and streams
// the whole 9*Cin x Cout weight matrix per 128 pixels: it is bound by L2->SM bandwidth.  Here
//   * a CTA owns a 32 x 8 pixel tile = two 128-row accumulators, so every weight slice that
//     reaches shared memory feeds twice as many MMAs;
//   * the three taps of one kernel column share ONE activation box: for (channel chunk, kx) a
//     single TMA box of (32 + 2) rows x 8 pixels x 64 channels lands as 34 swizzle atoms, and
//     tap ky of sub-tile j is just the descriptor start address + (ky + 16 j) atoms -- always
//     1024-byte aligned, so no descriptor tricks.  Activation traffic drops 2.6x;
//   * the CTA is persistent: accumulators are double-buffered in TMEM (2 x 256 columns), the
//     eight epilogue warps drain tile i while the MMA thread works on tile i + 1.
// Per 256 pixels: A 6 x 34 KB + B 18 x 16 KB = 492 KB of L2 reads (generic kernel: 1152 KB).
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tcgen;

namespace {

constexpr int NTHREADS = 320;       // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int NA = 3;                 // activation-unit ring
constexpr int NB = 6;                 // weight-slice ring
constexpr int TILE_H = 32, TILE_W = 8;
constexpr uint32_t A_UNIT = (TILE_H + 2) * TILE_W * 128;      // 34816 B

struct Tc3Params {
    FMap out, res, gate;
    const float *bias, *out_scale;
    int cout, kchunks;
    int act, post, act_channels;
    int tiles_x, ntiles;
    int in_pad;
    uint32_t b_bytes, b_slot;         // weight slice bytes (cout * 128) and its 1 KB-rounded slot
    int dbg;                          // AIVC_TC3_DBG bits: timing experiments only (wrong results)
};

__global__ void __launch_bounds__(NTHREADS, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const Tc3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[NA], a_empty[NA], b_full[NB], b_empty[NB], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[128], sscale[128];

    uint8_t *a_ring = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *b_ring = a_ring + NA * A_UNIT;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.cout;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t ia = 0, ib = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int y0 = (tile / p.tiles_x) * TILE_H, x0 = (tile % p.tiles_x) * TILE_W;
                for (int kc = 0; kc < p.kchunks; ++kc)
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t sa = ia % NA;
                        mbar_wait(&a_empty[sa], ((ia / NA) & 1u) ^ 1u);
                        if (p.dbg & 8) mbar_arrive(&a_full[sa]);
                        else {
                        mbar_expect_tx(&a_full[sa], A_UNIT);
                        tma_load_3d(a_ring + sa * A_UNIT, &tmA, &a_full[sa], kc * 64,
                                    x0 + kx - 1 + p.in_pad, y0 - 1 + p.in_pad);
                        }
                        ++ia;
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint32_t sb = ib % NB;
                            mbar_wait(&b_empty[sb], ((ib / NB) & 1u) ^ 1u);
                            if (p.dbg & 16) mbar_arrive(&b_full[sb]);
                            else {
                            mbar_expect_tx(&b_full[sb], p.b_bytes);
                            tma_load_3d(b_ring + sb * p.b_slot, &tmB, &b_full[sb], kc * 64, 0, ky * 3 + kx);
                            }
                            ++ib;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            uint32_t ia = 0, ib = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1u;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);     // epilogue drained this buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * 256u;
                uint32_t first = 1;
                for (int kc = 0; kc < p.kchunks; ++kc)
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t sa = ia % NA;
                        mbar_wait(&a_full[sa], (ia / NA) & 1u);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(a_ring + sa * A_UNIT);
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint32_t sb = ib % NB;
                            mbar_wait(&b_full[sb], (ib / NB) & 1u);
                            tc_fence_after();
                            const uint64_t bdesc = make_desc(smem_u32(b_ring + sb * p.b_slot), 128);
                            if (!(p.dbg & 4))
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t adesc = make_desc(a_addr + (uint32_t)(ky + 16 * j) * 1024u, 128);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_bf16(acc + (uint32_t)(j * 128), adesc + (uint64_t)(kk * 2),
                                              bdesc + (uint64_t)(kk * 2), idesc, (first && kk == 0) ? 0u : 1u);
                            }
                            first = 0;
                            umma_commit(&b_empty[sb]);
                            ++ib;
                        }
                        umma_commit(&a_empty[sa]);
                        ++ia;
                    }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // warp w reads TMEM lanes 32*(w%4).. of sub-tile j = (w-2)/4
        const int quarter = warp & 3, j = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const EpiCtx ctx = make_epi(p.out, p.res, p.gate, p.post, p.act_channels, p.out_scale != nullptr);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u;
            const int y0 = (tile / p.tiles_x) * TILE_H, x0 = (tile % p.tiles_x) * TILE_W;
            const int oy = y0 + 16 * j + row / TILE_W, ox = x0 + row % TILE_W;
            const bool valid = (oy < p.out.h) && (ox < p.out.w);
            mbar_wait(&acc_full[buf], (it >> 1) & 1u);
            tc_fence_after();
            const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256u + (uint32_t)(j * 128);
            if (!(p.dbg & 2)) epi_row_dispatch(p.act, tl, N, sbias, sscale, ctx, oy, ox, valid);
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
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

    const aivc_fmap &in = op->in;
    const size_t pix_b = (size_t)in.c_stride * 2, row_b = (size_t)in.pitch * pix_b;
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)(in.w + 2 * in.pad), (cuuint64_t)(in.h + 2 * in.pad)};
        cuuint64_t strides[2] = {pix_b, row_b};
        cuuint32_t box[3] = {64, TILE_W, TILE_H + 2};
        if (encode_map(&tmA, (char *)in.data + (size_t)in.c_off * 2, 3, dims, strides, box, 128, "A/3x3")) return 1;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)cout, 1};
        if (encode_map(&tmB, (void *)op->weight, 3, dims, strides, box, 128, "B/3x3")) return 1;
    }
    const size_t smem = 1024 + (size_t)NA * A_UNIT + (size_t)NB * p.b_slot;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        AIVC_CHECK_CUDA(cudaGetDevice(&dev));
        AIVC_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    AIVC_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         220 * 1024));
    const int grid = ntiles < sm_count ? ntiles : sm_count;
    AIVC_CHECK_CUDA(launch_pdl(conv3x3_tc_kernel, dim3(grid), dim3(NTHREADS), smem, st, tmA, tmB, p));
    AIVC_CHECK_LAUNCH("conv3x3_tc_kernel");
    return 0;
}
