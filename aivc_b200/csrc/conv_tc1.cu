// Persistent tcgen05 kernel for 1x1 stride-1 convolutions (the gated tail of SimplifiedAttention,
// attention.py:90-97: sigmoid(conv1x1(a)) * trunk + x, and the 1x1 GEMM form of the narrow output
// transposed conv).  These stages are HBM-bound -- per pixel they read cin + (gate, residual) and write
// cout channels for 2*cin*cout FLOPs -- so everything moves as bulk TMA traffic:
//   * the weight matrix is loaded once per CTA and stays in shared memory;
//   * a stage slot holds, for one 128-pixel tile, the activation tile (UMMA A operand, 128B swizzle) and
//     the gate / residual tiles as 32-channel chunks (64B swizzle); two slots are in flight;
//   * the epilogue reads gate and residual from shared memory, writes the bf16 result IN PLACE over the
//     gate (or residual, or a dedicated output) chunk, and one thread per team hands the chunks to TMA
//     stores; accumulators are double-buffered in TMEM.
// The generic kernel (conv_tc.cu) pays a full CTA prologue, un-pipelined loads and scattered 16-byte
// global accesses for gate / residual / output per 128 pixels.
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tcgen;
extern int g_aivc_kernel_class;

namespace {

constexpr int EPI_WARPS = 8, NTHREADS = 32 * (EPI_WARPS + 2), TMA_WARP = 8, MMA_WARP = 9;
constexpr int NS = 2;                                         // tile slots in flight (= TMEM accumulator buffers)

struct Tc1Params {
    FMap out;
    const float *bias, *out_scale;
    int cin, cout, kchunks, nch;                              // nch = cout / 32 epilogue chunks
    int act, post;
    int tw, th, tiles_x, ntiles;
    int has_gate, has_res, out_f16, stride2;
    uint32_t a_bytes, c_bytes;                                // activation tile; one gate/res/out tile (all chunks)
    uint32_t slot_bytes, off_g, off_r, off_o;                 // slot layout (off_o may alias off_g / off_r)
    uint32_t b_bytes;
};

__device__ __forceinline__ void border_store_chunk32(const FMap *m, int ch, int oy, int ox, uint4 a, uint4 b, uint4 c,
                                                  uint4 d) {
    const int pd = m->pad;
    const int y0 = (oy == 0) ? 0 : oy + pd, y1 = (oy == m->h - 1) ? oy + 2 * pd : oy + pd;
    const int x0 = (ox == 0) ? 0 : ox + pd, x1 = (ox == m->w - 1) ? ox + 2 * pd : ox + pd;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) {
            if (yy == oy + pd && xx == ox + pd) continue;       // the pixel itself goes out with the TMA store
            uint4 *q = reinterpret_cast<uint4 *>((__nv_bfloat16 *)m->data + ((size_t)yy * m->pitch + xx) * m->c_stride +
                                                 m->c_off + ch);
            q[0] = a; q[1] = b; q[2] = c; q[3] = d;
        }
}

__device__ __forceinline__ void unpack8(const uint4 &u, float *f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        f[2 * q] = __uint_as_float(w[q] << 16);
        f[2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
    }
}

template <int ACT>
__device__ __forceinline__ void epilogue(const Tc1Params &p, const CUtensorMap *tmO, uint8_t *slots, const float *sbias,
                                         uint64_t *full, uint64_t *empty, uint64_t *acc_full,
                                         uint64_t *acc_empty, uint32_t tmem_base, int warp, int lane) {
    const int quarter = warp & 3, team = warp >> 2;
    const int row = quarter * 32 + lane;
    const bool leader = quarter == 0 && lane == 0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
        const int y0 = (tile / p.tiles_x) * p.th, x0 = (tile % p.tiles_x) * p.tw;
        const int oy = y0 + row / p.tw, ox = x0 + row % p.tw;
        const bool valid = oy < p.out.h && ox < p.out.w;
        const bool edge = valid && p.out.pad != 0 && (oy == 0 || oy == p.out.h - 1 || ox == 0 || ox == p.out.w - 1);
        uint8_t *slot = slots + (size_t)s * p.slot_bytes;
        mbar_wait(&full[s], ph);                              // gate / residual tiles have landed
        mbar_wait(&acc_full[s], ph);
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + s * 256u;
        for (int c = team; c < p.nch; c += 2) {
            const int j0 = c * 32;
            float v[32];
            {
                uint32_t raw[32];
                tmem_ld32_nowait(tl + (uint32_t)j0, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
            }
            epi_bias_act16<ACT>(v, sbias, j0, 0);
            epi_bias_act16<ACT>(v + 16, sbias, j0 + 16, 0);
            // this thread's 64-byte row of chunk c: four 16-byte units at swizzled positions
            uint32_t off[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t o = (uint32_t)(row * 64 + k * 16);
                off[k] = (uint32_t)c * (128u * 64u) + (o ^ (((o >> 7) & 3u) << 4));
            }
            if (p.has_gate) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float g[8];
                    unpack8(*reinterpret_cast<const uint4 *>(slot + p.off_g + off[k]), g);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[8 * k + i] *= g[i];
                }
            }
            if (p.has_res) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float r[8];
                    unpack8(*reinterpret_cast<const uint4 *>(slot + p.off_r + off[k]), r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[8 * k + i] += r[i];
                }
            }
            post_apply16(p.post, v);
            post_apply16(p.post, v + 16);
            if (p.out_scale) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= __ldg(p.out_scale + j0 + i);      // (rare: gains live in g_a's last stage)
            }
            uint4 o4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = pack16x2(v[8 * k + 2 * q], v[8 * k + 2 * q + 1], p.out_f16 != 0);
                o4[k] = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4 *>(slot + p.off_o + off[k]) = o4[k];
            }
            if (edge) border_store_chunk32(&p.out, j0, oy, ox, o4[0], o4[1], o4[2], o4[3]);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[s]);                           // accumulator drained
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(1 + team, 128);
        if (leader) {
            for (int c = team; c < p.nch; c += 2)
                tma_store_3d(tmO, slot + p.off_o + (size_t)c * (128 * 64), c * 32, x0, y0);
            tma_store_commit();
            tma_store_wait_read();                            // the slot may be refilled once the stores have read it
            mbar_arrive(&empty[s]);
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ Tc1Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[NS], empty[NS], acc_full[NS], acc_empty[NS], b_full;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[256];

    uint8_t *wsm = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // resident weights
    uint8_t *slots = wsm + ((p.b_bytes + 1023u) & ~1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.cout;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1); mbar_init(&empty[s], 2);                      // two team leaders release a slot
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 32 * EPI_WARPS);
        }
        mbar_init(&b_full, 1);
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
            mbar_expect_tx(&b_full, p.b_bytes);               // weights: once per CTA
            for (int kc = 0; kc < p.kchunks; ++kc)
                tma_load_3d(wsm + (size_t)kc * N * 128, &tmB, &b_full, kc * 64, 0, 0);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
                const int y0 = (tile / p.tiles_x) * p.th, x0 = (tile % p.tiles_x) * p.tw;
                uint8_t *slot = slots + (size_t)s * p.slot_bytes;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], p.a_bytes + (uint32_t)(p.has_gate + p.has_res) * p.c_bytes);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    if (p.stride2)                             // (c, x parity 0, x/2, y parity 0, y/2) view of the input
                        tma_load_5d(slot + (size_t)kc * 128 * 128, &tmA, &full[s], kc * 64, 0, x0, 0, y0);
                    else
                        tma_load_3d(slot + (size_t)kc * 128 * 128, &tmA, &full[s], kc * 64, x0, y0);
                }
                for (int c = 0; c < p.nch; ++c) {
                    if (p.has_gate) tma_load_3d(slot + p.off_g + (size_t)c * (128 * 64), &tmG, &full[s], c * 32, x0, y0);
                    if (p.has_res) tma_load_3d(slot + p.off_r + (size_t)c * (128 * 64), &tmR, &full[s], c * 32, x0, y0);
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            mbar_wait(&b_full, 0);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
                mbar_wait(&acc_empty[s], ph ^ 1u);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(slots + (size_t)s * p.slot_bytes);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    const uint64_t adesc = make_desc(a_addr + (uint32_t)kc * 128u * 128u, 128);
                    const uint64_t bdesc = make_desc(smem_u32(wsm + (size_t)kc * N * 128), 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + s * 256u, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                  (uint32_t)(kc | kk));
                }
                umma_commit(&acc_full[s]);
            }
        }
    } else {
        switch (p.act) {
            case AIVC_ACT_LEAKY: epilogue<AIVC_ACT_LEAKY>(p, &tmO, slots, sbias, full, empty, acc_full, acc_empty, tmem_base, warp, lane); break;
            case AIVC_ACT_RELU: epilogue<AIVC_ACT_RELU>(p, &tmO, slots, sbias, full, empty, acc_full, acc_empty, tmem_base, warp, lane); break;
            case AIVC_ACT_SIGMOID: epilogue<AIVC_ACT_SIGMOID>(p, &tmO, slots, sbias, full, empty, acc_full, acc_empty, tmem_base, warp, lane); break;
            default: epilogue<AIVC_ACT_NONE>(p, &tmO, slots, sbias, full, empty, acc_full, acc_empty, tmem_base, warp, lane); break;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

bool bf16_chunkable(const aivc_fmap &m, bool f16_ok = false) {
    return (m.dtype == AIVC_BF16 || (f16_ok && m.dtype == AIVC_F16)) && m.c_off % 8 == 0 && m.c_stride % 8 == 0 && ((uintptr_t)m.data & 15) == 0;
}

// interior view of a bordered map as a {channels, w, h} tensor with 32-channel x tw x th boxes
int chunk_map(CUtensorMap *tm, const aivc_fmap &m, int channels, int tw, int th, const char *what) {
    const size_t pix = (size_t)m.c_stride * 2, rowb = (size_t)m.pitch * pix;
    cuuint64_t dims[3] = {(cuuint64_t)channels, (cuuint64_t)m.w, (cuuint64_t)m.h};
    cuuint64_t strides[2] = {pix, rowb};
    cuuint32_t box[3] = {32, (cuuint32_t)tw, (cuuint32_t)th};
    void *base = (char *)m.data + ((size_t)m.pad * m.pitch + m.pad) * pix + (size_t)m.c_off * 2;
    return encode_map(tm, base, 3, dims, strides, box, 64, what);
}

}  // namespace

static void pick_tile1(int mh, int mw, int *tw, int *th) {      // 128-pixel tile shape with the fewest tiles
    int best = 1 << 30, btw = 16;
    for (int w = 128; w >= 8; w >>= 1) {
        const int tiles = ceil_div(mw, w) * ceil_div(mh, 128 / w);
        if (tiles < best) { best = tiles; btw = w; }
    }
    *tw = btw;
    *th = 128 / btw;
}

// Returns -1 when the stage does not fit this kernel (caller falls through to the generic one).
int conv_tc1_run(const aivc_conv_op *op, cudaStream_t st) {
    const int cin = op->in.c, cout = op->out.c;
    if (op->kind != 0 || op->k != 1 || (op->stride != 1 && op->stride != 2)) return -1;
    if (op->stride == 2 && ((op->in.pitch & 1) || (op->in.rows & 1))) return -1;
    if (cin % 64 || cin > 256 || cout % 32 || cout > 256) return -1;
    if (op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN || op->act_channels) return -1;
    if (!bf16_chunkable(op->in) || !bf16_chunkable(op->out, true)) return -1;
    const bool has_gate = op->gate.data != nullptr, has_res = op->residual.data != nullptr;
    if (has_gate && !bf16_chunkable(op->gate)) return -1;
    if (has_res && !bf16_chunkable(op->residual)) return -1;

    Tc1Params p;
    memset(&p, 0, sizeof(p));
    p.out = to_dev(op->out);
    p.bias = op->bias; p.out_scale = op->out_scale;
    p.cin = cin; p.cout = cout; p.kchunks = cin / 64; p.nch = cout / 32;
    p.act = op->act; p.post = op->post;
    p.has_gate = has_gate; p.has_res = has_res; p.out_f16 = op->out.dtype == AIVC_F16; p.stride2 = op->stride == 2;
    pick_tile1(op->out.h, op->out.w, &p.tw, &p.th);
    p.tiles_x = ceil_div(op->out.w, p.tw);
    p.ntiles = p.tiles_x * ceil_div(op->out.h, p.th);
    p.a_bytes = (uint32_t)p.kchunks * 128u * 128u;
    p.c_bytes = (uint32_t)p.nch * 128u * 64u;
    p.b_bytes = (uint32_t)p.kchunks * (uint32_t)cout * 128u;
    uint32_t off_b = p.a_bytes;
    if (has_gate) { p.off_g = off_b; off_b += p.c_bytes; }
    if (has_res) { p.off_r = off_b; off_b += p.c_bytes; }
    if (has_gate) p.off_o = p.off_g;                       // result written in place over the gate ...
    else if (has_res) p.off_o = p.off_r;                   // ... or the residual ...
    else { p.off_o = off_b; off_b += p.c_bytes; }          // ... or into its own tile
    p.slot_bytes = (off_b + 1023u) & ~1023u;
    const size_t smem = 1024 + ((p.b_bytes + 1023u) & ~1023u) + (size_t)NS * p.slot_bytes;
    if (smem > 225 * 1024 + 512) return -1;

    CUtensorMap tmA, tmB, tmG, tmR, tmO;
    memset(&tmG, 0, sizeof(tmG));
    memset(&tmR, 0, sizeof(tmR));
    {
        const aivc_fmap &in = op->in;
        const size_t pix = (size_t)in.c_stride * 2, rowb = (size_t)in.pitch * pix;
        void *base = (char *)in.data + ((size_t)in.pad * in.pitch + in.pad) * pix + (size_t)in.c_off * 2;
        if (op->stride == 1) {
            cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)in.w, (cuuint64_t)in.h};
            cuuint64_t strides[2] = {pix, rowb};
            cuuint32_t box[3] = {64, (cuuint32_t)p.tw, (cuuint32_t)p.th};
            if (encode_map(&tmA, base, 3, dims, strides, box, 128, "A/1x1")) return 1;
        } else {
            // stride 2 (ChengResBlock 'down' skip path): the even pixels of the even rows, as a 5-D view
            cuuint64_t dims[5] = {(cuuint64_t)cin, 2, (cuuint64_t)((in.w + 1) / 2), 2, (cuuint64_t)((in.h + 1) / 2)};
            cuuint64_t strides[4] = {pix, 2 * pix, rowb, 2 * rowb};
            cuuint32_t box[5] = {64, 1, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th};
            if (encode_map(&tmA, base, 5, dims, strides, box, 128, "A/1x1s2")) return 1;
        }
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, 1};
        cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)cout, 1};
        if (encode_map(&tmB, (void *)op->weight, 3, dims, strides, box, 128, "B/1x1")) return 1;
    }
    if (has_gate && chunk_map(&tmG, op->gate, cout, p.tw, p.th, "gate/1x1")) return 1;
    if (has_res && chunk_map(&tmR, op->residual, cout, p.tw, p.th, "res/1x1")) return 1;
    if (chunk_map(&tmO, op->out, cout, p.tw, p.th, "out/1x1")) return 1;

    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        AIVC_CHECK_CUDA(cudaGetDevice(&dev));
        AIVC_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (smem_attr_once((const void *)conv1x1_tc_kernel, 225 * 1024 + 512)) return 1;
    g_aivc_kernel_class = AIVC_KC_TC1;
    const int grid = p.ntiles < sm_count ? p.ntiles : sm_count;
    AIVC_CHECK_CUDA(launch_pdl(conv1x1_tc_kernel, dim3(grid), dim3(NTHREADS), smem, st, tmA, tmB, tmG, tmR, tmO, p));
    AIVC_CHECK_LAUNCH("conv1x1_tc_kernel");
    return 0;
}
