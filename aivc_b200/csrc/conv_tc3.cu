// Persistent tcgen05 kernels built around ONE shared activation patch per tile:
//   conv3x3_tc_kernel<SUB,RES>   3x3 stride-1 conv (+ residual)            -- this header
//   conv3x3_tc_gdn_kernel        3x3 stride-1 conv + GDN / IGDN, norm GEMM with its A operand in TMEM
//   tconv3x3_tc_kernel           transposed 3x3 stride-2 conv, four output phases per patch
//
// conv3x3_tc_kernel: the layer that carries most of AIVC's FLOPs: 3x3, stride 1, replicate-padded
// convolution with Cin a multiple of 64 (ChengResBlock / ResBlock / attention trunks;
// custom_conv_layers.py:40-56, 112-126).
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
//     sixteen epilogue warps drain tile i while the MMA thread works on tile i + 1.
// Per 256 pixels: A 2 x 43 KB + B 18 x 16 KB = 375 KB of L2 reads (generic kernel: 1152 KB).
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tcgen;
extern int g_aivc_kernel_class;

namespace {

// SUB = 128-row accumulators (sub-tiles) per CTA tile.  SUB = 2: 32 x 8 pixel tiles, 18 warps, one CTA per
// SM (large layers).  SUB = 1: 16 x 8 pixel tiles, 10 warps, <= 112 KB of shared memory and 256 TMEM
// columns so that TWO CTAs share an SM, optionally with the output channels split over two work items
// (small layers: 68 x 120 latents give 75 tiles -> 150 work items for 148 SMs).
// The warp scheduler favours HIGHER warp ids, so the two latency-critical single-thread roles get the
// highest ids and are never starved by the ALU-heavy epilogue warps of their sub-partition.
template <int SUB>
struct Cfg {
    static constexpr int EPI_WARPS = 8 * SUB;               // teams = (sub-tile j, channel half h) x 4 lane quarters
    static constexpr int NTHREADS = 32 * (EPI_WARPS + 2);
    static constexpr int TMA_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
    static constexpr int TILE_H = 16 * SUB;
    static constexpr uint32_t PATCH_BYTES = (TILE_H + 2) * 10 * 128;          // per 64-channel chunk
    static constexpr uint32_t A_SLOT = (PATCH_BYTES + 1023) & ~1023u;
    static constexpr uint32_t TMEM_COLS = 256 * SUB;          // two accumulator buffers of SUB x 128 columns
};
constexpr int STAGE_BYTES = 128 * 64;   // one team's store staging tile: 128 rows x 32 channels bf16
constexpr int TILE_W = 8;
constexpr int PATCH_W = TILE_W + 2;                                             // 10 pixels
constexpr int NA = 2;               // activation patches in flight
constexpr int NG_MAX = 6;           // weight groups (3 taps = one kernel row) in flight

struct Tc3Params {
    FMap out, res, gate;
    const float *bias, *out_scale;
    int cout, kchunks;
    int ncta, nsplit;                 // output channels per work item; work items per pixel tile
    int act, post;
    int tiles_x, nitems;
    int nfull;                        // pixel tiles [0, nfull) are whole tiles; SUB = 2 only: the tiles behind them are
                                      // 16-row HALF tiles (one accumulator), see item_tile
    int in_pad;
    uint32_t b_bytes, b_slot;         // one tap's weight slice (ncta x 64 ch) and its 1 KB-rounded slot
    int ng;                           // weight-group ring depth
    int tma_store;                    // epilogue stores through smem + TMA: 1 = bf16 (32-channel multiples), 2 = split
                                      // bf16 (16-channel chunks, hi and lo tiles; o_lo = channel coordinate of lo)
    int o_lo;
    int skip_lo;                      // split-bf16 input with all-zero lo halves (AIVC_OP_IN_EXACT): no lo.Whi part
    int nsteps;                       // conv3x3_tc_kernel: patch steps per tile (patch_step)
    int x3, a_lo, b_lo;               // split-bf16 operands (AIVC_ENGINE_TC_X3): hi.Whi + lo.Whi + hi.Wlo; channel
                                      // coordinates of the lo halves of the activations / weights
};

// conv3x3_tc_kernel and conv3x3_tc_gdn_kernel walk the contraction patch by patch: step s loads ONE activation patch (channel coordinate ca)
// and runs it against one or two weight sets (channel coordinates cb[0..n)).  Split bf16: the hi patch of a 64-channel
// chunk meets Whi and Wlo, the lo patch meets Whi -- four patch loads per 128 input channels where the plain order
// of virtual chunks (hi.Whi, lo.Whi, hi.Wlo over all channels) needs six, and twice the tensor time per hi patch to fetch the next one behind.
__device__ __forceinline__ int patch_step(const Tc3Params &p, int s, int &ca, int (&cb)[2]) {
    if (!p.x3) {
        ca = cb[0] = cb[1] = s * 64;
        return 1;
    }
    const int per = p.skip_lo ? 1 : 2, j = s / per;
    const bool lo = s - j * per == 1;
    ca = j * 64 + (lo ? p.a_lo : 0);
    cb[0] = j * 64;
    cb[1] = j * 64 + p.b_lo;
    return lo ? 1 : 2;
}

// Work item -> pixel tile of conv3x3_tc_kernel; returns the number of 128-row accumulators (sub-tiles) in use.
// Whole tiles are numbered row-major over the image in bands of tile_h rows.  A layer whose whole tiles fill a
// fraction of the last wave (270x480: 510 tiles of 32x8 on 148 SMs = 3.45 waves, paid as 4) is cut differently by
// the host: nfull = a multiple of the grid size whole tiles, and the rest of the image as 16-row half tiles -- the
// remaining columns of the band the whole tiles end in, then the 16-row bands below it -- so that the last wave
// costs half a tile per SM (3.5 tile times instead of 4).  Items are dealt round-robin, nfull % grid == 0.
__host__ __device__ __forceinline__ int item_tile(const Tc3Params &p, int item, int tile_h, int &y0, int &x0, int &n0) {
    const int tile = item / p.nsplit;
    n0 = (item - tile * p.nsplit) * p.ncta;
    if (tile < p.nfull) {
        y0 = (tile / p.tiles_x) * tile_h;
        x0 = (tile % p.tiles_x) * TILE_W;
        return tile_h >> 4;
    }
    int k = tile - p.nfull;
    const int fb = p.nfull / p.tiles_x, xf = p.nfull - fb * p.tiles_x;
    const int part = xf ? 2 * (p.tiles_x - xf) : 0;            // half tiles that complete band fb
    int urow, col;
    if (k < part) {
        col = xf + (k >> 1);
        urow = 2 * fb + (k & 1);
    } else {
        k -= part;
        urow = 2 * fb + (xf ? 2 : 0) + k / p.tiles_x;
        col = k % p.tiles_x;
    }
    y0 = urow * 16;
    x0 = col * TILE_W;
    return 1;
}

// 16 channels of this thread's pixel as packed bf16 (two 16-byte loads)
__device__ __forceinline__ void ldg_bf16x16(const __nv_bfloat16 *ptr, uint4 &r0, uint4 &r1) {
    const uint4 *q = reinterpret_cast<const uint4 *>(ptr);
    r0 = q[0];
    r1 = q[1];
}
__device__ __forceinline__ void add_bf16x16(float *v, const uint4 &r0, const uint4 &r1) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint4 &r = h ? r1 : r0;
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[8 * h + 2 * q] += __uint_as_float(w[q] << 16);
            v[8 * h + 2 * q + 1] += __uint_as_float(w[q] & 0xFFFF0000u);
        }
    }
}

// border replicas of an edge pixel (rare: edge tiles only), kept out of line so that its addressing
// does not cost the hot path registers
__device__ __forceinline__ void border_store_bf16x16(const Tc3Params *p, int ch, int oy, int ox, uint4 lo, uint4 hi) {
    const FMap &m = p->out;
    const int pd = m.pad;
    const int y0 = (oy == 0) ? 0 : oy + pd, y1 = (oy == m.h - 1) ? oy + 2 * pd : oy + pd;
    const int x0 = (ox == 0) ? 0 : ox + pd, x1 = (ox == m.w - 1) ? ox + 2 * pd : ox + pd;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) {
            if (yy == oy + pd && xx == ox + pd) continue;       // the pixel itself goes out with the TMA store
            uint4 *q = reinterpret_cast<uint4 *>((__nv_bfloat16 *)m.data + ((size_t)yy * m.pitch + xx) * m.c_stride +
                                                 m.c_off + ch);
            q[0] = lo;
            q[1] = hi;
        }
}

// One 32-channel chunk of an accumulator row -> bias, activation, residual, post, scale -> bf16 ->
// 64-byte row of the team's swizzled (64B mode) staging tile.  `rc`: residual of the chunk, already in
// registers `rr` (prefetched while the tensor pipe was still working on this tile).
template <int ACT, bool USE_RES>
__device__ __forceinline__ void chunk32_to_stage(const Tc3Params &p, uint32_t taddr, int ch0, int n0,
                                                 const float *sbias, const float *sscale, uint4 (&rr)[4],
                                                 const __nv_bfloat16 *res_next, int oy, int ox, bool edge,
                                                 uint8_t *stage, int row) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int j0 = ch0 + cc * 16;
        float v[16];
        tmem_ld16(taddr + (uint32_t)j0, v);
        epi_bias_act16<ACT>(v, sbias, n0 + j0, 0);
        if (USE_RES) {
            add_bf16x16(v, rr[2 * cc], rr[2 * cc + 1]);
            // the registers just consumed take the same 16 channels of the team's NEXT 32-channel chunk
            if (res_next) ldg_bf16x16(res_next + cc * 16, rr[2 * cc], rr[2 * cc + 1]);
        }
        post_apply16(p.post, v);
        if (p.out_scale) {
            const float4 *s4 = reinterpret_cast<const float4 *>(sscale + n0 + j0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 sc = s4[q];
                v[4 * q] *= sc.x; v[4 * q + 1] *= sc.y; v[4 * q + 2] *= sc.z; v[4 * q + 3] *= sc.w;
            }
        }
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&b2);
        }
        const uint4 lo = make_uint4(w[0], w[1], w[2], w[3]), hi = make_uint4(w[4], w[5], w[6], w[7]);
        uint32_t off = (uint32_t)(row * 64 + cc * 32);
        off ^= ((off >> 7) & 3u) << 4;                          // 64B swizzle, as the TMA store expects
        *reinterpret_cast<uint4 *>(stage + off) = lo;
        *reinterpret_cast<uint4 *>(stage + (off ^ 16u)) = hi;
        if (edge) border_store_bf16x16(&p, n0 + j0, oy, ox, lo, hi);
    }
}

// Split-bf16 form: one 16-channel chunk of an accumulator row -> bias, activation, residual (loaded in place), post,
// scale -> hi and lo bf16 -> two 32-byte rows of the team's staging tile ([128 rows x 32 B hi | 128 rows x 32 B lo],
// 32B swizzle), which leave as two TMA stores.  Per-thread global stores of a row's 32-byte pieces (32 lines per warp
// instruction) cost this MMA-bound kernel 18 % (270x480: 91 us against 75 us with the stores removed): they crowd
// the memory pipe the single MMA-issuing and TMA-issuing threads poll their barriers through.
template <int ACT>
__device__ __forceinline__ void chunk16_x3_to_stage(const Tc3Params &p, const EpiCtx &ctx, uint32_t taddr, int j0, int n0,
                                                    const float *sbias, const float *sscale, int oy, int ox, bool valid,
                                                    bool edge, uint8_t *stage, int row) {
    float v[16];
    tmem_ld16(taddr + (uint32_t)j0, v);
    epi_bias_act16<ACT>(v, sbias, n0 + j0, 0, true);
    if (p.res.data && valid) {
        float r[16];
        load16(p.res, ctx.res_vec, oy, ox, n0 + j0, 16, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
    }
    post_apply16(p.post, v);
    if (p.out_scale) {
        const float4 *s4 = reinterpret_cast<const float4 *>(sscale + n0 + j0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 sc = s4[q];
            v[4 * q] *= sc.x; v[4 * q + 1] *= sc.y; v[4 * q + 2] *= sc.z; v[4 * q + 3] *= sc.w;
        }
    }
    uint32_t w[8], l[8];
    split16(v, w, l);
    uint32_t off = (uint32_t)(row * 32);
    off ^= ((off >> 7) & 1u) << 4;                              // 32B swizzle, as the TMA store expects
    *reinterpret_cast<uint4 *>(stage + off) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4 *>(stage + (off ^ 16u)) = make_uint4(w[4], w[5], w[6], w[7]);
    *reinterpret_cast<uint4 *>(stage + 4096 + off) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4 *>(stage + 4096 + (off ^ 16u)) = make_uint4(l[4], l[5], l[6], l[7]);
    if (edge) store16(p.out, ctx.out_vec, oy, ox, n0 + j0, 16, v);       // border replicas (and the pixel once more)
}

template <int SUB, int ACT, bool RES>
__device__ __forceinline__ void epilogue_team(const Tc3Params &p, const CUtensorMap *tmO, uint8_t *stage,
                                              const float *sbias, const float *sscale, uint64_t *acc_full,
                                              uint64_t *acc_empty, uint32_t tmem_base, int warp, int lane,
                                              int first, int stride, int nunits) {
    constexpr int TILE_H = Cfg<SUB>::TILE_H;
    const int quarter = warp & 3, team = warp >> 2, j = team >> 1, h = team & 1;
    const int row = quarter * 32 + lane;
    const bool leader = ((warp & 3) == 0) && lane == 0;
    const int N = p.ncta;
    const int half = (N == 64) ? 32 : 64;                   // channels per team
    const int c_lo = h * half, c_hi = min(N, c_lo + half);
    // bf16 residuals with 16-byte aligned channel groups are prefetched into registers one 32-channel
    // chunk ahead -- the first chunk before the accumulator is even complete -- so their L2 latency
    // (about 1 us under load) never sits on the epilogue's critical path
    // (RES kernels are only launched with TMA stores and a residual the host found prefetchable)
    constexpr bool res_fast = RES;
    EpiCtx ctx = make_epi(p.out, p.res, p.gate, p.post, 0, p.out_scale != nullptr, p.x3 != 0);
    uint8_t *my_stage = stage + team * STAGE_BYTES;
    uint32_t it = 0;
    bool store_pending = false;
    for (int u = first; u < nunits; u += stride, ++it) {
        const uint32_t buf = it & 1u;
        int y0, x0, n0;
        const bool active = j < item_tile(p, u, TILE_H, y0, x0, n0);      // (half tiles: the j = 1 teams only hand the buffer back)
        const int oy = y0 + 16 * j + row / TILE_W, ox = x0 + row % TILE_W;
        const bool valid = active && (oy < p.out.h) && (ox < p.out.w);
        const bool use_res = res_fast && valid;
        ctx.ch0 = n0;                                             // (channels [n0, n0 + N) of out / res / gate)
        const __nv_bfloat16 *res_px = use_res ? (const __nv_bfloat16 *)p.res.data + fm_index(p.res, oy, ox, n0) : nullptr;
        uint4 rr[4];
        if (res_fast) {
#pragma unroll
            for (int i = 0; i < 4; ++i) rr[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (p.tma_store == 2 && p.res.data && valid && c_lo < N) {
            // split-bf16 residual rows (loaded in place, 16 channels at a time, chunk16_x3_to_stage): this thread's
            // hi and lo lines are asked into L2 while the tensor pipe still works on the tile, so that the four
            // dependent loads behind the wait are L2 hits (the residual was written two layers ago: HBM by now)
            const __nv_bfloat16 *r = (const __nv_bfloat16 *)p.res.data + fm_index(p.res, oy, ox, n0 + c_lo);
            prefetch_l2(r);
            if (p.res.dtype == AIVC_BF16X2) prefetch_l2(r + (p.res.c_stride >> 1));
        }
        if (use_res && c_lo < N) {                              // first 32 channels: issued before the wait
            ldg_bf16x16(res_px + c_lo, rr[0], rr[1]);
            ldg_bf16x16(res_px + c_lo + 16, rr[2], rr[3]);
        }
        mbar_wait(&acc_full[buf], (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)(128 * SUB) + (uint32_t)(j * 128);
        if (c_lo < N && active && p.tma_store == 2) {
            const bool edge = valid && p.out.pad != 0 && (oy == 0 || oy == p.out.h - 1 || ox == 0 || ox == p.out.w - 1);
            for (int ch0 = c_lo; ch0 < c_hi; ch0 += 16) {
                if (store_pending) {                           // staging tile still being read by the last stores?
                    if (leader) tma_store_wait_read();
                    named_bar_sync(1 + team, 128);
                }
                chunk16_x3_to_stage<ACT>(p, ctx, tl, ch0, n0, sbias, sscale, oy, ox, valid, edge, my_stage, row);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(1 + team, 128);
                if (leader) {
                    tma_store_3d(tmO, my_stage, n0 + ch0, x0, y0 + 16 * j);
                    tma_store_3d(tmO, my_stage + 4096, p.o_lo + n0 + ch0, x0, y0 + 16 * j);
                    tma_store_commit();
                }
                store_pending = true;
            }
        } else if (c_lo < N && active) {
            if (p.tma_store) {
                const bool edge = valid && p.out.pad != 0 && (oy == 0 || oy == p.out.h - 1 || ox == 0 || ox == p.out.w - 1);
                for (int ch0 = c_lo; ch0 < c_hi; ch0 += 32) {
                    if (store_pending) {                       // staging tile still being read by the last store?
                        if (leader) tma_store_wait_read();
                        named_bar_sync(1 + team, 128);
                    }
                    // (warp-uniform choice: tcgen05.ld inside is .sync.aligned; rows outside the image add zeros)
                    if (RES)
                        chunk32_to_stage<ACT, true>(p, tl, ch0, n0, sbias, sscale, rr,
                                                    (use_res && ch0 + 32 < c_hi) ? res_px + ch0 + 32 : nullptr, oy, ox, edge, my_stage, row);
                    else
                        chunk32_to_stage<ACT, false>(p, tl, ch0, n0, sbias, sscale, rr, nullptr, oy, ox, edge, my_stage, row);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    named_bar_sync(1 + team, 128);
                    if (leader) {
                        tma_store_3d(tmO, my_stage, n0 + ch0, x0, y0 + 16 * j);
                        tma_store_commit();
                    }
                    store_pending = true;
                }
            } else {
                epi_row<ACT>(tl, c_hi, sbias + n0, sscale + n0, ctx, oy, ox, valid, c_lo);
            }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
    }
    if (store_pending && leader) tma_store_wait_read();        // smem must outlive the last store's read
}

template <int SUB, bool RES>
__global__ void __launch_bounds__(Cfg<SUB>::NTHREADS, SUB == 1 ? 2 : 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ Tc3Params p) {
    using K = Cfg<SUB>;
    constexpr int NTHREADS = K::NTHREADS, TILE_H = K::TILE_H;
    constexpr uint32_t A_SLOT = K::A_SLOT, PATCH_BYTES = K::PATCH_BYTES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[NA], a_empty[NA], g_full[NG_MAX], g_empty[NG_MAX], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[128], sscale[128];

    uint8_t *a_ring = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *g_ring = a_ring + NA * A_SLOT;
    const uint32_t g_slot = 3u * p.b_slot;
    uint8_t *stage = g_ring + (size_t)p.ng * g_slot;          // TEAMS x 8 KB

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.ncta;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NG_MAX; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 32 * K::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == K::MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)),
                     "r"(K::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    stage_vec(sbias, p.bias, p.cout, 0.f, tid, NTHREADS);     // weights/bias never depend on the prior grid
    stage_vec(sscale, p.out_scale, p.cout, 1.f, tid, NTHREADS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait_prior_grid();                                    // activations / residuals below do

    if (warp == K::TMA_WARP) {
        // ===================== TMA producer =====================
        // per work item and 64-channel chunk: ONE activation patch ((TILE_H+2) x 10 pixels) that serves
        // all nine taps, and three weight groups (one kernel row = 3 taps each), every group on one barrier
        if (lane == 0) {
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0;          // ring slot / phase, advanced incrementally
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
                int y0, x0, n0;
                item_tile(p, item, TILE_H, y0, x0, n0);
                for (int st = 0; st < p.nsteps; ++st) {
                    int ca, cb[2];
                    const int nsets = patch_step(p, st, ca, cb);
                    mbar_wait(&a_empty[sa], pa ^ 1u);
                    mbar_expect_tx(&a_full[sa], PATCH_BYTES);
                    tma_load_3d(a_ring + sa * A_SLOT, &tmA, &a_full[sa], ca, x0 - 1 + p.in_pad, y0 - 1 + p.in_pad);
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                    for (int g = 0; g < 3 * nsets; ++g) {
                        const int ky = g >= 3 ? g - 3 : g;
                        mbar_wait(&g_empty[sg], pg ^ 1u);
                        mbar_expect_tx(&g_full[sg], 3u * p.b_bytes);
                        uint8_t *dst = g_ring + sg * g_slot;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
                            tma_load_3d(dst + kx * p.b_slot, &tmB, &g_full[sg], cb[g >= 3], n0, ky * 3 + kx);
                        if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == K::MMA_WARP) {
        // ===================== MMA issuer =====================
        // 12 SUB MMAs (3 taps x SUB sub-tiles x 4 k-steps) per barrier wait: the issuing thread's
        // bookkeeping must stay well below the tensor cycles they take
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            // descriptor template: K-major, 128B swizzle, 8-row groups PATCH_W rows apart (the
            // swizzle is a function of the absolute smem address, so any 128 B-aligned start works)
            const uint64_t a_tmpl = (1ull << 16) | ((uint64_t)((PATCH_W * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0, it = 0;
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
                const uint32_t buf = it & 1u;
                int y0_, x0_, n0_;
                const int nsub = item_tile(p, item, TILE_H, y0_, x0_, n0_);
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);     // epilogue drained this buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * (uint32_t)(128 * SUB);
                uint32_t accum = 0;
                for (int st = 0; st < p.nsteps; ++st) {
                    int ca_, cb_[2];
                    const int ngroups = 3 * patch_step(p, st, ca_, cb_);
                    mbar_wait(&a_full[sa], pa);
                    const uint32_t a_addr = smem_u32(a_ring + sa * A_SLOT);
                    for (int g = 0; g < ngroups; ++g) {
                        const int ky = g >= 3 ? g - 3 : g;
                        mbar_wait(&g_full[sg], pg);
                        tc_fence_after();
                        const uint32_t g_addr = smem_u32(g_ring + sg * g_slot);
                        {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint64_t bdesc = make_desc(g_addr + kx * p.b_slot, 128);
#pragma unroll
                                for (int j = 0; j < SUB; ++j) {
                                    if (SUB > 1 && j >= nsub) break;
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
        // ===================== epilogue (warps 0 .. 8 SUB - 1) =====================
        switch (p.act) {
            case AIVC_ACT_LEAKY: epilogue_team<SUB, AIVC_ACT_LEAKY, RES>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane, blockIdx.x, gridDim.x, p.nitems); break;
            case AIVC_ACT_RELU: epilogue_team<SUB, AIVC_ACT_RELU, RES>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane, blockIdx.x, gridDim.x, p.nitems); break;
            case AIVC_ACT_SIGMOID: epilogue_team<SUB, AIVC_ACT_SIGMOID, RES>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane, blockIdx.x, gridDim.x, p.nitems); break;
            default: epilogue_team<SUB, AIVC_ACT_NONE, RES>(p, &tmO, stage, sbias, sscale, acc_full, acc_empty, tmem_base, warp, lane, blockIdx.x, gridDim.x, p.nitems); break;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == K::MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS)
                     : "memory");
    }
}

template <int SUB, bool RES>
int launch_tc3(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &o, const Tc3Params &p, int grid,
               size_t smem, cudaStream_t st) {
    if (smem_attr_once((const void *)conv3x3_tc_kernel<SUB, RES>, 220 * 1024)) return 1;
    AIVC_CHECK_CUDA(launch_pdl(conv3x3_tc_kernel<SUB, RES>, dim3(grid), dim3(Cfg<SUB>::NTHREADS), smem, st, a, b, o, p));
    AIVC_CHECK_LAUNCH("conv3x3_tc_kernel");
    return 0;
}

// =====================================================================================================
// 3x3 stride-1 convolution + GDN / IGDN (misc_layers.py:113-154) in ONE persistent kernel, built on the
// 16 x 8 pixel tile so that TMEM holds everything the normalisation needs:
//     columns [0,256)   acc[buf] = conv(x)                       (main loop, as above; double buffered)
//     columns [256,384) norm     = (acc+bias)^2 . gamma^T         (8 MMAs per tile)
//     columns [384,448) (acc+bias)^2 as packed bf16: the A operand of the norm GEMM, read from TMEM
//     columns [448,512) split-bf16 mode only: the lo half of (acc+bias)^2 (norm = hi.Ghi + lo.Ghi + hi.Glo,
//                       gamma resident as [hi | lo], 2 N^2 bf16)
// Epilogue pass 1 (16 warps, one pixel x N/4 channels per thread) reads acc and writes (acc+bias)^2 back to
// TMEM (tcgen05.st); the MMA thread multiplies it with gamma (resident in shared memory for the whole
// kernel) between two weight groups of the NEXT tile's main loop, as soon as the operand is there; pass 2
// reads acc and norm, applies x * (beta+norm)^-+1/2, residual, post activation, gain, and stores through
// per-team staging tiles + TMA.
struct GdnBars {
    uint64_t a_full[NA], a_empty[NA], g_full[NG_MAX], g_empty[NG_MAX];
    uint64_t acc_full[2], acc_empty[2], norm_full, norm_empty, xsq_full, xsq_empty, gamma_full;
};

constexpr int GDN_EPI_WARPS = 16, GDN_THREADS = 32 * (GDN_EPI_WARPS + 2), GDN_TMA_WARP = 16, GDN_MMA_WARP = 17;

template <bool RES>
__global__ void __launch_bounds__(GDN_THREADS, 1)
conv3x3_tc_gdn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmG,
                      const __grid_constant__ Tc3Params p, const float *gdn_beta, int inverse) {
    using K = Cfg<1>;
    constexpr int NTHREADS = GDN_THREADS, TILE_H = K::TILE_H;
    constexpr uint32_t A_SLOT = K::A_SLOT, PATCH_BYTES = K::PATCH_BYTES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ GdnBars bars;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[128], sscale[128], sbeta[128];

    const int N = p.cout, nk = N / 64;                        // norm GEMM: K = N in 64-channel chunks
    uint8_t *a_ring = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *g_ring = a_ring + NA * A_SLOT;
    const uint32_t g_slot = 3u * p.b_slot;
    uint8_t *stage = g_ring + (size_t)p.ng * g_slot;          // 4 teams x 8 KB (none in split-bf16 mode)
    uint8_t *gam = stage + (p.x3 ? 0 : 4 * STAGE_BYTES);      // [nk (x3: 2 nk)][N rows][128 B], 128B swizzle (TMA)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&bars.a_full[s], 1); mbar_init(&bars.a_empty[s], 1); }
        for (int s = 0; s < NG_MAX; ++s) { mbar_init(&bars.g_full[s], 1); mbar_init(&bars.g_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars.acc_full[s], 1); mbar_init(&bars.acc_empty[s], 32 * GDN_EPI_WARPS); }
        mbar_init(&bars.norm_full, 1); mbar_init(&bars.norm_empty, 32 * GDN_EPI_WARPS);
        mbar_init(&bars.xsq_full, 32 * GDN_EPI_WARPS); mbar_init(&bars.xsq_empty, 1); mbar_init(&bars.gamma_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == GDN_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    stage_vec(sbias, p.bias, N, 0.f, tid, NTHREADS);
    stage_vec(sscale, p.out_scale, N, 1.f, tid, NTHREADS);
    stage_vec(sbeta, gdn_beta, N, 0.f, tid, NTHREADS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait_prior_grid();

    if (warp == GDN_TMA_WARP) {
        if (lane == 0) {
            const int ngam = p.x3 ? 2 * nk : nk;                                 // gamma ([hi | lo]): once per CTA
            mbar_expect_tx(&bars.gamma_full, (uint32_t)(ngam * N * 128));
            for (int c = 0; c < ngam; ++c) tma_load_2d(gam + (size_t)c * N * 128, &tmG, &bars.gamma_full, c * 64, 0);
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0;
            for (int tile = blockIdx.x; tile < p.nitems; tile += gridDim.x) {
                const int y0 = (tile / p.tiles_x) * TILE_H, x0 = (tile % p.tiles_x) * TILE_W;
                for (int st = 0; st < p.nsteps; ++st) {                          // patch by patch, see patch_step
                    int ca, cb[2];
                    const int nsets = patch_step(p, st, ca, cb);
                    mbar_wait(&bars.a_empty[sa], pa ^ 1u);
                    mbar_expect_tx(&bars.a_full[sa], PATCH_BYTES);
                    tma_load_3d(a_ring + sa * A_SLOT, &tmA, &bars.a_full[sa], ca, x0 - 1 + p.in_pad, y0 - 1 + p.in_pad);
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                    for (int g = 0; g < 3 * nsets; ++g) {
                        const int ky = g >= 3 ? g - 3 : g;
                        mbar_wait(&bars.g_empty[sg], pg ^ 1u);
                        mbar_expect_tx(&bars.g_full[sg], 3u * p.b_bytes);
                        uint8_t *dst = g_ring + sg * g_slot;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
                            tma_load_3d(dst + kx * p.b_slot, &tmB, &bars.g_full[sg], cb[g >= 3], 0, ky * 3 + kx);
                        if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == GDN_MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            const uint64_t a_tmpl = (1ull << 16) | ((uint64_t)((PATCH_W * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
            // norm(k) = xsq of tile k (TMEM, written by the epilogue) . gamma^T (smem)
            auto issue_norm = [&](uint32_t k) {
                if (k == 0) mbar_wait(&bars.gamma_full, 0);
                mbar_wait(&bars.norm_empty, (k & 1u) ^ 1u);                      // pass 2 of tile k-1 has read norm
                mbar_wait(&bars.xsq_full, k & 1u);                               // pass 1 of tile k has written x^2
                tc_fence_after();
                for (int part = 0; part < (p.x3 ? 3 : 1); ++part)                // hi.Ghi, lo.Ghi, hi.Glo
                    for (int c = 0; c < nk; ++c) {
                        const uint64_t bd = make_desc(smem_u32(gam + (size_t)(c + (part == 2 ? nk : 0)) * N * 128), 128);
                        const uint32_t xa = tmem_base + (part == 1 ? 448u : 384u);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_ts(tmem_base + 256u, xa + (uint32_t)((c * 4 + kk) * 8),
                                         bd + (uint64_t)(kk * 2), idesc, (uint32_t)(part | c | kk));
                    }
                umma_commit(&bars.norm_full);
                umma_commit(&bars.xsq_empty);
            };
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.nitems; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1u;
                // norm(it-1) is slipped in between two weight groups of this tile as soon as the epilogue has
                // delivered x^2 -- never blocking on it while shared-memory slots wait to be consumed
                bool norm_pending = it > 0;
                mbar_wait(&bars.acc_empty[buf], ((it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * 128u;
                uint32_t accum = 0;
                for (int st = 0; st < p.nsteps; ++st) {
                    int ca_, cb_[2];
                    const int ngroups = 3 * patch_step(p, st, ca_, cb_);
                    mbar_wait(&bars.a_full[sa], pa);
                    const uint32_t a_addr = smem_u32(a_ring + sa * A_SLOT);
                    for (int g = 0; g < ngroups; ++g) {
                        const int ky = g >= 3 ? g - 3 : g;
                        mbar_wait(&bars.g_full[sg], pg);
                        tc_fence_after();
                        const uint32_t g_addr = smem_u32(g_ring + sg * g_slot);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint64_t bdesc = make_desc(g_addr + kx * p.b_slot, 128);
                            const uint32_t start = a_addr + (uint32_t)(ky * PATCH_W + kx) * 128u;
                            const uint64_t adesc = a_tmpl | (uint64_t)((start >> 4) & 0x3FFF);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16(acc, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                          accum | (uint32_t)(kx | kk));
                        }
                        accum = 1;
                        umma_commit(&bars.g_empty[sg]);
                        if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                        if (norm_pending && mbar_test(&bars.xsq_full, (it - 1) & 1u)) {
                            issue_norm(it - 1);
                            norm_pending = false;
                        }
                    }
                    umma_commit(&bars.a_empty[sa]);
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                }
                umma_commit(&bars.acc_full[buf]);
                if (norm_pending) issue_norm(it - 1);
            }
            if (it > 0) issue_norm(it - 1);
        }
    } else {
        // ===================== epilogue: 4 teams (channel quarters) x 4 lane quarters =====================
        // sixteen warps, nothing carried in registers between the passes: pass 2 reads acc again (TMEM reads
        // are cheap), so every thread has N/4 channels of one pixel and plenty of warps hide the latencies
        const int quarter = warp & 3, team = warp >> 2;
        const int row = quarter * 32 + lane;
        const bool leader = quarter == 0 && lane == 0;
        const int per = N / 4;                                // 32 (N = 128) or 16 (N = 64) channels per thread
        const int c_lo = team * per;
        const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint8_t *my_stage = stage + team * STAGE_BYTES;
        EpiCtx ctx = make_epi(p.out, p.res, p.gate, p.post, 0, p.out_scale != nullptr, p.x3 != 0);
        const bool staged = p.tma_store && per == 32;         // one 32-channel TMA store per team and tile
        bool store_pending = false;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.nitems; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u;
            const int y0 = (tile / p.tiles_x) * TILE_H, x0 = (tile % p.tiles_x) * TILE_W;
            const int oy = y0 + row / TILE_W, ox = x0 + row % TILE_W;
            const bool valid = (oy < p.out.h) && (ox < p.out.w);
            const bool use_res = RES && valid;
            const bool edge = valid && p.out.pad != 0 && (oy == 0 || oy == p.out.h - 1 || ox == 0 || ox == p.out.w - 1);
            const size_t out_elem = valid ? fm_index(p.out, oy, ox, 0) : 0;
            if (p.x3 && p.res.data && valid) {                 // split-bf16 residual rows (loaded in place in pass 2): into L2 now
                const __nv_bfloat16 *r = (const __nv_bfloat16 *)p.res.data + fm_index(p.res, oy, ox, c_lo);
                prefetch_l2(r);
                if (p.res.dtype == AIVC_BF16X2) prefetch_l2(r + (p.res.c_stride >> 1));
            }
            // ---- pass 1: (acc + bias)^2 -> packed bf16 in TMEM (A operand of the norm GEMM)
            mbar_wait(&bars.acc_full[buf], (it >> 1) & 1u);
            mbar_wait(&bars.xsq_empty, (it & 1u) ^ 1u);        // norm MMAs of the previous tile have read x^2
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c * 16 < per) {
                    const int j0 = c_lo + c * 16;
                    float v[16];
                    tmem_ld16(tl + buf * 128u + (uint32_t)j0, v);
                    uint32_t w[8];
                    if (p.x3) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b = *reinterpret_cast<const float4 *>(sbias + j0 + 4 * q);
                            const float x0 = v[4 * q] + b.x, x1 = v[4 * q + 1] + b.y, x2 = v[4 * q + 2] + b.z, x3 = v[4 * q + 3] + b.w;
                            v[4 * q] = x0 * x0; v[4 * q + 1] = x1 * x1; v[4 * q + 2] = x2 * x2; v[4 * q + 3] = x3 * x3;
                        }
                        uint32_t wl[8];
                        split16(v, w, wl);
                        tmem_st8(tl + 448u + (uint32_t)(j0 >> 1), wl);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b = *reinterpret_cast<const float4 *>(sbias + j0 + 4 * q);
                            const float x0 = v[4 * q] + b.x, x1 = v[4 * q + 1] + b.y, x2 = v[4 * q + 2] + b.z, x3 = v[4 * q + 3] + b.w;
                            const __nv_bfloat162 lo = __floats2bfloat162_rn(x0 * x0, x1 * x1);
                            const __nv_bfloat162 hi = __floats2bfloat162_rn(x2 * x2, x3 * x3);
                            w[2 * q] = *reinterpret_cast<const uint32_t *>(&lo);
                            w[2 * q + 1] = *reinterpret_cast<const uint32_t *>(&hi);
                        }
                    }
                    tmem_st8(tl + 384u + (uint32_t)(j0 >> 1), w);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars.xsq_full);
            // ---- pass 2
            uint4 rr[4];
            if (RES) {                  // residual of this thread's channels: in flight while the norm GEMM runs
#pragma unroll
                for (int i = 0; i < 4; ++i) rr[i] = make_uint4(0u, 0u, 0u, 0u);
                if (use_res) {
                    const __nv_bfloat16 *res_px = (const __nv_bfloat16 *)p.res.data + fm_index(p.res, oy, ox, c_lo);
                    ldg_bf16x16(res_px, rr[0], rr[1]);
                    if (per == 32) ldg_bf16x16(res_px + 16, rr[2], rr[3]);
                }
            }
            mbar_wait(&bars.norm_full, it & 1u);
            tc_fence_after();
            if (staged && store_pending) {                     // staging tile still being read by the last tile's store?
                if (leader) tma_store_wait_read();
                named_bar_sync(1 + team, 128);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c * 16 < per) {
                    const int j0 = c_lo + c * 16;
                    float v[16], nr[16];
                    tmem_ld16(tl + buf * 128u + (uint32_t)j0, v);
                    tmem_ld16(tl + 256u + (uint32_t)j0, nr);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bi = *reinterpret_cast<const float4 *>(sbias + j0 + 4 * q);
                        const float4 be = *reinterpret_cast<const float4 *>(sbeta + j0 + 4 * q);
                        const float bb[4] = {bi.x, bi.y, bi.z, bi.w}, ee[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float xx = v[4 * q + e] + bb[e];
                            const float t = nr[4 * q + e] + ee[e];
                            if (p.x3) {
                                // fp32-faithful mode: MUFU.RSQ + one Newton step = t^-1/2 to ~1 ulp (the operands
                                // carry 2^-17 already), a quarter of the instructions of IEEE sqrt + division,
                                // which made pass 2 the longest link of this kernel's chain.  t >= beta_bound^2 > 0.
                                float rs;
                                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(t));
                                const float ht = 0.5f * t;
                                rs = fmaf(rs, fmaf(-ht * rs, rs, 0.5f), rs);
                                v[4 * q + e] = inverse ? xx * (t * rs) : xx * rs;
                            } else {
                                float rs;                                       // MUFU.RSQ: 2^-22 relative, far below bf16
                                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(t));
                                v[4 * q + e] = inverse ? xx * (t * rs) : xx * rs;
                            }
                        }
                    }
                    if (staged) {
                        if (RES) add_bf16x16(v, rr[2 * c], rr[2 * c + 1]);
                        post_apply16(p.post, v);
                        if (p.out_scale) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] *= sscale[j0 + i];
                        }
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            w[i] = *reinterpret_cast<const uint32_t *>(&b2);
                        }
                        const uint4 lo = make_uint4(w[0], w[1], w[2], w[3]), hi = make_uint4(w[4], w[5], w[6], w[7]);
                        uint32_t off = (uint32_t)(row * 64 + c * 32);
                        off ^= ((off >> 7) & 3u) << 4;                  // 64B swizzle, as the TMA store expects
                        *reinterpret_cast<uint4 *>(my_stage + off) = lo;
                        *reinterpret_cast<uint4 *>(my_stage + (off ^ 16u)) = hi;
                        if (edge) border_store_bf16x16(&p, j0, oy, ox, lo, hi);
                    } else if (valid) {
                        // generic stores (fp32 / unaligned outputs, slow residuals, gates, 64-channel layers)
                        epi_tail16(v, ctx, sscale, oy, ox, j0, !edge, out_elem);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars.norm_empty);
            mbar_arrive(&bars.acc_empty[buf]);
            if (staged) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(1 + team, 128);
                if (leader) {
                    tma_store_3d(&tmO, my_stage, c_lo, x0, y0);
                    tma_store_commit();
                }
                store_pending = true;
            }
        }
        if (store_pending && leader) tma_store_wait_read();    // smem must outlive the last store's read
    }

    tc_fence_before();
    __syncthreads();
    if (warp == GDN_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <bool RES>
int launch_tc3_gdn(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &o, const CUtensorMap &g,
                   const Tc3Params &p, const float *beta, int inverse, int grid, size_t smem, cudaStream_t st) {
    if (smem_attr_once((const void *)conv3x3_tc_gdn_kernel<RES>, 225 * 1024)) return 1;
    AIVC_CHECK_CUDA(launch_pdl(conv3x3_tc_gdn_kernel<RES>, dim3(grid), dim3(GDN_THREADS), smem, st, a, b, o, g, p,
                               beta, inverse));
    AIVC_CHECK_LAUNCH("conv3x3_tc_gdn_kernel");
    return 0;
}

// =====================================================================================================
// Transposed 3x3 stride-2 convolution (UpscalingLayer(3, C, C), custom_conv_layers.py:183-253) as ONE
// persistent kernel.  Output pixel (2y+py, 2x+px) of phase (py,px) sums 1 / 2 / 2 / 4 taps of the
// input neighbourhood {y, y+1} x {x, x+1}, so all four phases of a 16 x 8 INPUT tile read the same
// 18 x 10 activation patch (zero fill beyond the image = the transposed conv's implicit padding) and
// every one of the nine weight slices is used exactly once per tile:
//     step A  phases (0,0),(0,1): taps (1,1) | (1,0) (1,2)                     -> TMEM regions 0, 1
//     step B  phases (1,0),(1,1): taps (0,1) (2,1) | (0,0) (0,2) (2,0) (2,2)   -> TMEM regions 2, 3
// The epilogue drains step A while the tensor pipe runs step B and step B during the next tile's step A.
// The generic kernel runs the four phases as separate CTAs, each re-loading a 128-pixel tile per tap.
// Output goes through 5-D TMA stores (c, x parity, x/2, y parity, y/2).
struct TapT { unsigned char widx, ph, ay, ax; };
__constant__ TapT c_tconv3_taps[9] = {
    {4, 0, 0, 0}, {3, 1, 0, 1}, {5, 1, 0, 0},             // step A: (ky,kx) = (1,1) | (1,0) (1,2)
    {1, 2, 1, 0}, {7, 2, 0, 0}, {0, 3, 1, 1},             // step B, group 1: (0,1) (2,1) | (0,0)
    {2, 3, 1, 0}, {6, 3, 0, 1}, {8, 3, 0, 0}};            // step B, group 2: (0,2) (2,0) (2,2)

constexpr int TC_NA = 4;                                  // activation patches in flight (two tiles x two chunks)

__device__ __forceinline__ void border_store_up_bf16x32(const Tc3Params *p, int ch, int oy, int ox, uint4 a, uint4 b,
                                                     uint4 c, uint4 d) {
    const FMap &m = p->out;
    const int pd = m.pad;
    const int y0 = (oy == 0) ? 0 : oy + pd, y1 = (oy == m.h - 1) ? oy + 2 * pd : oy + pd;
    const int x0 = (ox == 0) ? 0 : ox + pd, x1 = (ox == m.w - 1) ? ox + 2 * pd : ox + pd;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) {
            if (yy == oy + pd && xx == ox + pd) continue;       // the pixel itself goes out with the TMA store
            uint4 *q = reinterpret_cast<uint4 *>((__nv_bfloat16 *)m.data + ((size_t)yy * m.pitch + xx) * m.c_stride +
                                                 m.c_off + ch);
            q[0] = a; q[1] = b; q[2] = c; q[3] = d;
        }
}

// X3 (split-bf16 output): every 32-channel chunk leaves as TWO tiles through the same staging buffer -- the hi halves
// at channel j0, then the lo halves at channel (c_stride / 2) + j0 of the output pixel.
template <int ACT, bool X3>
__device__ __forceinline__ void tconv_epilogue(const Tc3Params &p, const CUtensorMap *tmO, uint8_t *stage,
                                               const float *sbias, uint64_t *acc_full, uint64_t *acc_empty,
                                               uint32_t tmem_base, int warp, int lane) {
    const int quarter = warp & 3, team = warp >> 2;           // team = 32-channel quarter of the 128 outputs
    const int row = quarter * 32 + lane;
    const bool leader = quarter == 0 && lane == 0;
    const int j0 = team * 32;
    uint8_t *my_stage = stage + team * STAGE_BYTES;
    bool store_pending = false;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.nitems; tile += gridDim.x, ++it) {
        const int y0 = (tile / p.tiles_x) * 16, x0 = (tile % p.tiles_x) * TILE_W;
        const int iy = y0 + row / TILE_W, ix = x0 + row % TILE_W;
        const bool valid = 2 * iy < p.out.h && 2 * ix < p.out.w;       // (input pixel inside the image)
        const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {
            mbar_wait(&acc_full[step], it & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
                const int ph = step * 2 + sub;                 // py = step, px = sub
                float v[32];
                {
                    uint32_t raw[32];
                    tmem_ld32_nowait(tl + (uint32_t)(ph * 128 + j0), raw);
                    if (store_pending) {                       // staging tile still being read by the last store?
                        if (leader) tma_store_wait_read();
                        named_bar_sync(1 + team, 128);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                }
                if (sub == 1) {                                // both regions of this step are in registers / stored
                    tc_fence_before();
                    mbar_arrive(&acc_empty[step]);
                }
                const int oy = 2 * iy + step, ox = 2 * ix + sub;
                const bool edge = valid && p.out.pad != 0 && (oy == 0 || oy == p.out.h - 1 || ox == 0 || ox == p.out.w - 1);
                if (j0 < p.cout) {
                    epi_bias_act16<ACT>(v, sbias, j0, 0, X3);
                    epi_bias_act16<ACT>(v + 16, sbias, j0 + 16, 0, X3);
                }
#pragma unroll 1
                for (int half = 0; half < (X3 ? 2 : 1); ++half) {
                    if (half == 1) {                           // the hi tile must have been read before lo overwrites it
                        if (leader) tma_store_wait_read();
                        named_bar_sync(1 + team, 128);
                    }
                    if (j0 < p.cout) {
                        uint4 o4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t w[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float a = v[8 * k + 2 * q], b = v[8 * k + 2 * q + 1];
                                const __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
                                if (X3 && half == 0) {         // keep the remainder for the lo pass
                                    const float2 hf = __bfloat1622float2(b2);
                                    v[8 * k + 2 * q] = a - hf.x; v[8 * k + 2 * q + 1] = b - hf.y;
                                }
                                w[q] = *reinterpret_cast<const uint32_t *>(&b2);
                            }
                            o4[k] = make_uint4(w[0], w[1], w[2], w[3]);
                            uint32_t off = (uint32_t)(row * 64 + k * 16);
                            off ^= ((off >> 7) & 3u) << 4;      // 64B swizzle, as the TMA store expects
                            *reinterpret_cast<uint4 *>(my_stage + off) = o4[k];
                        }
                        if (edge) border_store_up_bf16x32(&p, j0 + half * (p.out.c_stride >> 1), oy, ox, o4[0], o4[1], o4[2], o4[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    named_bar_sync(1 + team, 128);
                    if (leader && j0 < p.cout) {
                        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                                         (uint64_t)tmO),
                                     "r"(smem_u32(my_stage)), "r"(j0 + half * (p.out.c_stride >> 1)), "r"(sub), "r"(x0), "r"(step), "r"(y0)
                                     : "memory");
                        tma_store_commit();
                    }
                }
                store_pending = true;
            }
        }
    }
    if (store_pending && leader) tma_store_wait_read();
}

__global__ void __launch_bounds__(Cfg<2>::NTHREADS, 1)
tconv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ Tc3Params p) {
    using K1 = Cfg<1>;
    constexpr int NTHREADS = Cfg<2>::NTHREADS, TMA_WARP = 16, MMA_WARP = 17;
    constexpr uint32_t A_SLOT = K1::A_SLOT, PATCH_BYTES = K1::PATCH_BYTES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[TC_NA], a_empty[TC_NA], g_full[NG_MAX], g_empty[NG_MAX], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[128];

    uint8_t *a_ring = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *g_ring = a_ring + TC_NA * A_SLOT;
    const uint32_t g_slot = 3u * p.b_slot;
    uint8_t *stage = g_ring + (size_t)p.ng * g_slot;          // 4 teams x 8 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.cout;
    const int nparts = p.x3 ? 3 : 1, npatch = p.x3 ? 2 * p.kchunks : p.kchunks;      // (npatch <= TC_NA)

    if (tid == 0) {
        for (int s = 0; s < TC_NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NG_MAX; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 512); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    stage_vec(sbias, p.bias, N, 0.f, tid, NTHREADS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait_prior_grid();

    if (warp == TMA_WARP) {
        if (lane == 0) {
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0;
            for (int tile = blockIdx.x; tile < p.nitems; tile += gridDim.x) {
                const int y0 = (tile / p.tiles_x) * 16, x0 = (tile % p.tiles_x) * TILE_W;
                // all patches of the tile stay for the whole tile: the kchunks hi chunks, then (split-bf16) the lo chunks
                for (int pi = 0; pi < npatch; ++pi) {
                    const int kc = pi % p.kchunks;
                    mbar_wait(&a_empty[sa], pa ^ 1u);
                    mbar_expect_tx(&a_full[sa], PATCH_BYTES);
                    tma_load_3d(a_ring + sa * A_SLOT, &tmA, &a_full[sa], kc * 64 + (pi >= p.kchunks ? p.a_lo : 0), x0, y0);
                    if (++sa == TC_NA) { sa = 0; pa ^= 1u; }
                }
                for (int g = 0; g < 3; ++g)                    // group 0 = step A, groups 1, 2 = step B
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        for (int part = 0; part < nparts; ++part) {     // hi.Whi, lo.Whi, hi.Wlo
                            mbar_wait(&g_empty[sg], pg ^ 1u);
                            mbar_expect_tx(&g_full[sg], 3u * p.b_bytes);
                            uint8_t *dst = g_ring + sg * g_slot;
#pragma unroll
                            for (int t = 0; t < 3; ++t)
                                tma_load_3d(dst + t * p.b_slot, &tmB, &g_full[sg], kc * 64 + (part == 2 ? p.b_lo : 0), 0,
                                            c_tconv3_taps[g * 3 + t].widx);
                            if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                        }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            const uint64_t a_tmpl = (1ull << 16) | ((uint64_t)((PATCH_W * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint32_t sa = 0, pa = 0, sg = 0, pg = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.nitems; tile += gridDim.x, ++it) {
                uint32_t a_addr[TC_NA];
                const uint32_t sa0 = sa, pa0 = pa;
                for (int pi = 0; pi < npatch; ++pi) {
                    a_addr[pi] = smem_u32(a_ring + sa * A_SLOT);
                    if (++sa == TC_NA) { sa = 0; pa ^= 1u; }
                }
                uint32_t started = 0;                          // phases whose region already holds a partial sum
                uint32_t waited = 0;                           // patches of this tile whose arrival has been waited for
                for (int g = 0; g < 3; ++g) {
                    if (g == 0 || g == 1) {                    // a new step starts: its regions must be drained
                        mbar_wait(&acc_empty[g], (it & 1u) ^ 1u);
                        tc_fence_after();
                    }
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        for (int part = 0; part < nparts; ++part) {
                            const int pi = kc + (part == 1 ? p.kchunks : 0);      // hi patch, except for lo.Whi
                            if (!((waited >> pi) & 1u)) {      // first use of this patch
                                uint32_t s_ = sa0 + (uint32_t)pi, p_ = pa0;
                                if (s_ >= TC_NA) { s_ -= TC_NA; p_ ^= 1u; }
                                mbar_wait(&a_full[s_], p_);
                                waited |= 1u << pi;
                            }
                            mbar_wait(&g_full[sg], pg);
                            tc_fence_after();
                            const uint32_t g_addr = smem_u32(g_ring + sg * g_slot);
#pragma unroll
                            for (int t = 0; t < 3; ++t) {
                                const TapT tp = c_tconv3_taps[g * 3 + t];
                                const uint64_t bdesc = make_desc(g_addr + t * p.b_slot, 128);
                                const uint32_t start = a_addr[pi] + (uint32_t)(tp.ay * PATCH_W + tp.ax) * 128u;
                                const uint64_t adesc = a_tmpl | (uint64_t)((start >> 4) & 0x3FFF);
                                const uint32_t acc = tmem_base + (uint32_t)tp.ph * 128u;
                                const uint32_t had = (started >> tp.ph) & 1u;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_bf16(acc, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, had | (uint32_t)kk);
                                started |= 1u << tp.ph;
                            }
                            umma_commit(&g_empty[sg]);
                            if (++sg == (uint32_t)p.ng) { sg = 0; pg ^= 1u; }
                        }
                    if (g == 0) umma_commit(&acc_full[0]);     // step A complete
                }
                umma_commit(&acc_full[1]);                     // step B complete
                {                                              // the tile's patches are free
                    uint32_t s_ = sa0;
                    for (int pi = 0; pi < npatch; ++pi) {
                        umma_commit(&a_empty[s_]);
                        if (++s_ == TC_NA) s_ = 0;
                    }
                }
            }
        }
    } else {
        if (p.x3) {
            switch (p.act) {
                case AIVC_ACT_LEAKY: tconv_epilogue<AIVC_ACT_LEAKY, true>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
                case AIVC_ACT_RELU: tconv_epilogue<AIVC_ACT_RELU, true>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
                case AIVC_ACT_SIGMOID: tconv_epilogue<AIVC_ACT_SIGMOID, true>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
                default: tconv_epilogue<AIVC_ACT_NONE, true>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
            }
        } else {
            switch (p.act) {
                case AIVC_ACT_LEAKY: tconv_epilogue<AIVC_ACT_LEAKY, false>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
                case AIVC_ACT_RELU: tconv_epilogue<AIVC_ACT_RELU, false>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
                case AIVC_ACT_SIGMOID: tconv_epilogue<AIVC_ACT_SIGMOID, false>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
                default: tconv_epilogue<AIVC_ACT_NONE, false>(p, &tmO, stage, sbias, acc_full, acc_empty, tmem_base, warp, lane); break;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

static int sm_count_cached() {
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sm_count = 148;
    }
    return sm_count;
}

// Split-bf16 stages on 32-row tiles: whole tiles for as many full waves as the image holds, the rest as 16-row half
// tiles (item_tile) when that shortens the step.  Costs in half-tile units; a half tile is dearer than half a whole
// one (every weight group feeds half as many MMAs, so the two-deep group ring runs latency-bound: measured at 270x480,
// three whole tiles + one half tile per SM take 92.5 us against 96.5 us for four whole ones).
// p.tiles_x / p.nsplit = 1 set by the caller; sets p.nfull, p.nitems.
static void choose_tiling(int out_h, int tiles_x, int G, Tc3Params &p) {
    const int ntiles = tiles_x * ceil_div(out_h, 32);
    p.nfull = ntiles;
    p.nitems = ntiles;
    const int urows = ceil_div(out_h, 16), units = urows * tiles_x;
    const int q = units / (2 * G), nfull = q * G, nhalf = units - 2 * nfull;
    const double mixed = 2.0 * q + 1.5 * ceil_div(nhalf, G), whole = 2.0 * ceil_div(ntiles, G);
    if (q > 0 && nfull <= (urows / 2) * tiles_x && mixed < whole - 0.25) {
        p.nfull = nfull;
        p.nitems = nfull + nhalf;
    }
}

// Introspection for the host-side tests (no GPU needed): the work items conv3x3_tc_kernel<2, *> would process for a
// split-bf16 h x w layer on sm_count SMs, as (y0, x0, rows) triples in item order; returns the number of items.
extern "C" int aivc_debug_tc3_tiling(int h, int w, int sm_count, int *tiles, int cap) {
    Tc3Params p;
    memset(&p, 0, sizeof(p));
    p.tiles_x = ceil_div(w, TILE_W);
    p.nsplit = 1;
    p.ncta = 128;
    choose_tiling(h, p.tiles_x, sm_count, p);
    for (int i = 0; i < p.nitems && i < cap; ++i) {
        int y0, x0, n0;
        const int nsub = item_tile(p, i, 32, y0, x0, n0);
        tiles[3 * i] = y0; tiles[3 * i + 1] = x0; tiles[3 * i + 2] = 16 * nsub;
    }
    return p.nitems;
}

// Returns -1 when the stage does not fit this kernel (caller falls through to the generic one).
int conv_tc3_run(const aivc_conv_op *op, cudaStream_t st) {
    const int cin = op->in.c, cout = op->out.c;
    const bool x3 = op->engine == AIVC_ENGINE_TC_X3;         // split-bf16 operands, see Tc3Params
    if (op->kind != 0 || op->k != 3 || op->stride != 1) return -1;
    if (cin % 64 || cout % 16 || cout > 128) return -1;
    const bool gdn = op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN;
    if (gdn && (cout != 128 && cout != 64)) return -1;
    if (op->in.dtype != (x3 ? AIVC_BF16X2 : AIVC_BF16) || op->in.pad < 1 || op->in.c_off % 8 || op->in.c_stride % 16) return -1;
    if (op->act_channels) return -1;
    const int tiles_x = ceil_div(op->out.w, TILE_W);
    // big layers: 32-row tiles, one CTA per SM.  Fewer than two of those per SM: 16-row tiles, two CTAs
    // per SM and (wide layers) the output channels of a tile split over two work items.
    const int tiles32 = tiles_x * ceil_div(op->out.h, 32);
    int sub = tiles32 >= 296 ? 2 : 1;
    if (gdn) sub = 1;                                          // conv + GDN kernel: 16-row tiles, one CTA per SM
    else if (tiles32 >= 148 && tiles32 < 296 && !x3) return -1;     // one-and-a-bit waves of 32-row tiles: the generic
                                                               // 128-pixel-tile kernel fills the chip better (measured,
                                                               // plain bf16; split bf16: 16-row tiles, see below)
    const int tile_h = 16 * sub;
    const int ntiles = tiles_x * ceil_div(op->out.h, tile_h);

    Tc3Params p;
    memset(&p, 0, sizeof(p));
    p.out = to_dev(op->out);
    if (op->residual.data) p.res = to_dev(op->residual);
    if (op->gate.data) p.gate = to_dev(op->gate);
    p.bias = op->bias; p.out_scale = op->out_scale;
    p.cout = cout; p.kchunks = cin / 64;
    p.skip_lo = (x3 && (op->flags & AIVC_OP_IN_EXACT)) ? 1 : 0;
    p.x3 = x3 ? 1 : 0; p.a_lo = op->in.c_stride / 2; p.b_lo = cin;
    p.nsteps = (x3 && !p.skip_lo ? 2 : 1) * p.kchunks;
    // small layers (SUB = 1): output channels split over two work items, two CTAs per SM, shallow weight ring
    p.nsplit = (sub == 1 && !gdn && cout == 128 && ntiles < 296) ? 2 : 1;
    p.ncta = cout / p.nsplit;
    p.act = op->act; p.post = op->post;
    p.tiles_x = tiles_x; p.nitems = ntiles * p.nsplit; p.nfull = ntiles; p.in_pad = op->in.pad;
    if (sub == 2 && x3 && !gdn) choose_tiling(op->out.h, tiles_x, sm_count_cached(), p);
    p.b_bytes = (uint32_t)p.ncta * 128u;                       // weight rows one CTA stages per tap
    p.b_slot = (p.b_bytes + 1023u) & ~1023u;
    const size_t a_slot = sub == 1 ? Cfg<1>::A_SLOT : Cfg<2>::A_SLOT;
    // conv + GDN: gamma + two more staging tiles; split-bf16: gamma [hi | lo], no staging tiles (generic stores)
    const size_t gdn_bytes = gdn ? (x3 ? (size_t)cout * cout * 4 - 2 * STAGE_BYTES : (size_t)cout * cout * 2 + 2 * STAGE_BYTES) : 0;
    const size_t fixed = 1024 + (size_t)NA * a_slot + (size_t)2 * sub * STAGE_BYTES + gdn_bytes;
    // SUB = 1 aims at two CTAs per SM (<= 111 KB each) when two weight groups fit in that
    const size_t two_cta = 111 * 1024;
    const bool two_per_sm = sub == 1 && !gdn && fixed + 2 * 3 * (size_t)p.b_slot <= two_cta;
    const size_t budget = sub == 2 ? 219 * 1024 : (two_per_sm ? two_cta : (gdn ? 224 * 1024 : 165 * 1024));
    p.ng = NG_MAX;
    while (fixed + (size_t)p.ng * 3 * p.b_slot > budget && p.ng > 1) --p.ng;
    if (p.ng < 2) return -1;

    const aivc_fmap &in = op->in;
    const size_t pix_b = (size_t)in.c_stride * 2, row_b = (size_t)in.pitch * pix_b;
    CUtensorMap tmA, tmB;
    {   // (a split-bf16 view also reaches its lo half at + c_stride / 2)
        cuuint64_t dims[3] = {(cuuint64_t)(x3 ? p.a_lo + cin : cin), (cuuint64_t)(in.w + 2 * in.pad), (cuuint64_t)(in.h + 2 * in.pad)};
        cuuint64_t strides[2] = {pix_b, row_b};
        cuuint32_t box[3] = {64, PATCH_W, (cuuint32_t)(tile_h + 2)};
        if (encode_map(&tmA, (char *)in.data + (size_t)in.c_off * 2, 3, dims, strides, box, 128, "A/3x3")) return 1;
    }
    {
        const int wcin = x3 ? 2 * cin : cin;                   // x3: [tap][cout][hi cin | lo cin]
        cuuint64_t dims[3] = {(cuuint64_t)wcin, (cuuint64_t)cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)wcin * 2, (cuuint64_t)wcin * cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)p.ncta, 1};
        if (encode_map(&tmB, (void *)op->weight, 3, dims, strides, box, 128, "B/3x3")) return 1;
    }
    CUtensorMap tmO;
    memset(&tmO, 0, sizeof(tmO));
    const aivc_fmap &o = op->out;
    // staged epilogue (smem + TMA store, residual prefetched into registers): bf16 output in 32-channel
    // multiples, no gate, residual (if any) bf16 with 16-byte aligned channel groups.  Split-bf16 and fp32
    // outputs take the generic per-thread stores.
    const aivc_fmap &rs = op->residual;
    const bool res_ok = !rs.data || (rs.dtype == AIVC_BF16 && rs.c_off % 8 == 0 && rs.c_stride % 8 == 0 &&
                                     ((uintptr_t)rs.data & 15) == 0);
    p.tma_store = (!x3 && o.dtype == AIVC_BF16 && o.c_off % 8 == 0 && o.c_stride % 8 == 0 && p.ncta % 32 == 0 &&
                   ((uintptr_t)o.data & 15) == 0 && !op->gate.data && res_ok) ? 1 : 0;
    // (the 16-row-tile form of the kernel runs the small, latency-bound layers: per-thread stores there)
    if (x3 && !gdn && sub == 2 && o.dtype == AIVC_BF16X2 && o.c_off % 8 == 0 && o.c_stride % 16 == 0 && p.ncta % 32 == 0 &&
        ((uintptr_t)o.data & 15) == 0 && !op->gate.data && (!rs.data || rs.dtype == AIVC_BF16X2 || rs.dtype == AIVC_BF16)) {
        p.tma_store = 2;
        p.o_lo = o.c_stride / 2;
        const size_t opix = (size_t)o.c_stride * 2, orow = (size_t)o.pitch * opix;
        cuuint64_t dims[3] = {(cuuint64_t)(p.o_lo + cout), (cuuint64_t)o.w, (cuuint64_t)o.h};  // hi and lo halves of the interior
        cuuint64_t strides[2] = {opix, orow};
        cuuint32_t box[3] = {16, TILE_W, 16};
        void *base = (char *)o.data + ((size_t)o.pad * o.pitch + o.pad) * opix + (size_t)o.c_off * 2;
        if (encode_map(&tmO, base, 3, dims, strides, box, 32, "O/3x3 split")) return 1;
    } else if (p.tma_store) {
        const size_t opix = (size_t)o.c_stride * 2, orow = (size_t)o.pitch * opix;
        cuuint64_t dims[3] = {(cuuint64_t)cout, (cuuint64_t)o.w, (cuuint64_t)o.h};     // interior only: OOB rows/cols are clipped
        cuuint64_t strides[2] = {opix, orow};
        cuuint32_t box[3] = {32, TILE_W, 16};
        void *base = (char *)o.data + ((size_t)o.pad * o.pitch + o.pad) * opix + (size_t)o.c_off * 2;
        if (encode_map(&tmO, base, 3, dims, strides, box, 64, "O/3x3")) return 1;
    }
    const size_t smem = fixed + (size_t)p.ng * 3 * p.b_slot;
    const int sm_count = sm_count_cached();
    const bool res = p.tma_store == 1 && rs.data;            // residual prefetched into registers
    g_aivc_kernel_class = gdn ? AIVC_KC_TC3_GDN : AIVC_KC_TC3;
    if (gdn) {
        CUtensorMap tmG;
        const int gk = x3 ? 2 * cout : cout;                   // x3: gamma rows are [hi | lo]
        cuuint64_t dims[2] = {(cuuint64_t)gk, (cuuint64_t)cout};
        cuuint64_t strides[1] = {(cuuint64_t)gk * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)cout};
        if (encode_map(&tmG, (void *)op->gdn_gamma, 2, dims, strides, box, 128, "gamma/3x3")) return 1;
        const int grid = ntiles < sm_count ? ntiles : sm_count;
        const int inverse = op->act == AIVC_ACT_IGDN ? 1 : 0;
        return res ? launch_tc3_gdn<true>(tmA, tmB, tmO, tmG, p, op->gdn_beta, inverse, grid, smem, st)
                   : launch_tc3_gdn<false>(tmA, tmB, tmO, tmG, p, op->gdn_beta, inverse, grid, smem, st);
    }
    const int slots = sm_count * (two_per_sm ? 2 : 1);
    const int grid = p.nitems < slots ? p.nitems : slots;
    if (sub == 1) return res ? launch_tc3<1, true>(tmA, tmB, tmO, p, grid, smem, st) : launch_tc3<1, false>(tmA, tmB, tmO, p, grid, smem, st);
    return res ? launch_tc3<2, true>(tmA, tmB, tmO, p, grid, smem, st) : launch_tc3<2, false>(tmA, tmB, tmO, p, grid, smem, st);
}

// Transposed 3x3 stride-2 stage on the persistent kernel above; -1 = not eligible.
int tconv_tc3_run(const aivc_conv_op *op, cudaStream_t st) {
    const int cin = op->in.c, cout = op->out.c;
    const bool x3 = op->engine == AIVC_ENGINE_TC_X3;         // split-bf16 operands and output
    if (op->kind != 1 || op->k != 3 || op->stride != 2) return -1;
    if (cin % 64 || cin > 128 || cout % 32 || cout > 128) return -1;
    if (op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN || op->act_channels) return -1;
    if (op->residual.data || op->gate.data || op->out_scale || op->post != AIVC_POST_NONE) return -1;
    const aivc_fmap &in = op->in, &o = op->out;
    const int dt = x3 ? AIVC_BF16X2 : AIVC_BF16, al = x3 ? 16 : 8;
    if (in.dtype != dt || in.c_off % 8 || in.c_stride % al) return -1;
    if (o.dtype != dt || o.c_off % 8 || o.c_stride % al || ((uintptr_t)o.data & 15)) return -1;
    const int tiles_x = ceil_div(in.w, TILE_W), ntiles = tiles_x * ceil_div(in.h, 16);
    if (ntiles < 120) return -1;                            // tiny maps: the phase-parallel generic kernel fills the chip better

    Tc3Params p;
    memset(&p, 0, sizeof(p));
    p.out = to_dev(o);
    p.bias = op->bias;
    p.cout = cout; p.ncta = cout; p.nsplit = 1; p.kchunks = cin / 64;
    p.x3 = x3 ? 1 : 0; p.a_lo = in.c_stride / 2; p.b_lo = cin;
    p.act = op->act; p.post = op->post;
    p.tiles_x = tiles_x; p.nitems = ntiles;
    p.b_bytes = (uint32_t)cout * 128u;
    p.b_slot = (p.b_bytes + 1023u) & ~1023u;
    p.ng = 2;
    const size_t smem = 1024 + (size_t)TC_NA * Cfg<1>::A_SLOT + (size_t)p.ng * 3 * p.b_slot + 4 * STAGE_BYTES;

    CUtensorMap tmA, tmB, tmO;
    const size_t pix_b = (size_t)in.c_stride * 2, row_b = (size_t)in.pitch * pix_b;
    {   // interior only: rows / columns beyond the image read as zero (the transposed conv's padding)
        cuuint64_t dims[3] = {(cuuint64_t)(x3 ? p.a_lo + cin : cin), (cuuint64_t)in.w, (cuuint64_t)in.h};
        cuuint64_t strides[2] = {pix_b, row_b};
        cuuint32_t box[3] = {64, PATCH_W, 18};
        void *base = (char *)in.data + ((size_t)in.pad * in.pitch + in.pad) * pix_b + (size_t)in.c_off * 2;
        if (encode_map(&tmA, base, 3, dims, strides, box, 128, "A/tconv3")) return 1;
    }
    {
        const int wcin = x3 ? 2 * cin : cin;                   // x3: [tap][cout][hi cin | lo cin]
        cuuint64_t dims[3] = {(cuuint64_t)wcin, (cuuint64_t)cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)wcin * 2, (cuuint64_t)wcin * cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)cout, 1};
        if (encode_map(&tmB, (void *)op->weight, 3, dims, strides, box, 128, "B/tconv3")) return 1;
    }
    {   // output interior as (c, x parity, x/2, y parity, y/2): one phase of a tile is a {32, 1, 8, 1, 16} box
        const size_t opix = (size_t)o.c_stride * 2, orow = (size_t)o.pitch * opix;
        cuuint64_t dims[5] = {(cuuint64_t)(x3 ? o.c_stride / 2 + cout : cout), 2, (cuuint64_t)in.w, 2, (cuuint64_t)in.h};
        cuuint64_t strides[4] = {opix, 2 * opix, orow, 2 * orow};
        cuuint32_t box[5] = {32, 1, TILE_W, 1, 16};
        void *base = (char *)o.data + ((size_t)o.pad * o.pitch + o.pad) * opix + (size_t)o.c_off * 2;
        if (encode_map(&tmO, base, 5, dims, strides, box, 64, "O/tconv3")) return 1;
    }
    const int sm_count = sm_count_cached();
    g_aivc_kernel_class = AIVC_KC_TCONV3;
    if (smem_attr_once((const void *)tconv3x3_tc_kernel, 225 * 1024)) return 1;
    const int grid = ntiles < sm_count ? ntiles : sm_count;
    AIVC_CHECK_CUDA(launch_pdl(tconv3x3_tc_kernel, dim3(grid), dim3(Cfg<2>::NTHREADS), smem, st, tmA, tmB, tmO, p));
    AIVC_CHECK_LAUNCH("tconv3x3_tc_kernel");
    return 0;
}
