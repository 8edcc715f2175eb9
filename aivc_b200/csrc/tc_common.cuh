// Shared device helpers of the tcgen05 engines: PTX wrappers (mbarrier, TMA, tcgen05.mma /
// commit / ld, TMEM fences), UMMA descriptors, and the vectorised epilogue load/store.
#pragma once
#include <cuda.h>
#include <utility>
#include "common.cuh"

namespace tcgen {

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ uint32_t mbar_test(uint64_t *b, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    if (mbar_try_wait(addr, parity)) return;                  // fast path: no clock reads
    const long long t0 = clock64();
    while (!mbar_try_wait(addr, parity))
        if (clock64() - t0 > 4000000000LL) __trap();          // never hang the GPU
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 consecutive columns of this thread's TMEM lane in one instruction; the caller waits (tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 consecutive 32-bit columns of this thread's TMEM lane (e.g. 16 packed bf16 of an A operand)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem, 128 lanes x 16 bf16 packed in 8 columns] . B[smem descriptor]^T
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Programmatic dependent launch: let the next kernel's CTAs start their prologue while this grid
// drains, and wait for the previous grid (and its memory) before touching dependent data.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand descriptor: `rowbytes` = bytes of one row (= swizzle span: 128 / 64 / 32)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int rowbytes) {
    const uint64_t layout = rowbytes == 128 ? 2ull : (rowbytes == 64 ? 4ull : 6ull);
    const uint64_t sbo = (uint64_t)(8 * rowbytes) >> 4;       // 8-row core-matrix group stride
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
    // c = f32 (1<<4), a = b = bf16 (1<<7, 1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------------------------ epilogue IO
// two floats -> one 32-bit word of the map's 16-bit type (bf16, or fp16 for the pixel-domain partial sums)
__device__ __forceinline__ uint32_t pack16x2(float a, float b, bool f16) {
    if (f16) {
        const __half2 h2 = __floats2half2_rn(a, b);
        return *reinterpret_cast<const uint32_t *>(&h2);
    }
    const __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&b2);
}
struct VecInfo {
    bool vec;       // 16-channel groups are 16-byte aligned
};
__device__ __forceinline__ bool fmap_vec_ok(const FMap &m) {
    const int al = (m.dtype == AIVC_F32) ? 4 : 8;
    if (m.dtype == AIVC_BF16X2 && (m.c_stride >> 1) % 8) return false;      // lo half 16-byte aligned too
    return (m.c_off % al == 0) && (m.c_stride % al == 0) && (((uintptr_t)m.data & 15) == 0);
}

// 32-byte global accesses (sm_100: LDG / STG .256): a thread's 16 channels of bf16 in ONE instruction and one full
// sector, where two 16-byte accesses cost the load / store unit two passes over the warp's 32 lines
__device__ __forceinline__ bool aligned32(const void *p) { return ((uintptr_t)p & 31) == 0; }
__device__ __forceinline__ void stg256(void *p, const uint32_t *w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 16 packed 16-bit channels at p (16-byte aligned at least)
__device__ __forceinline__ void ldg_16ch(const void *p, uint4 &a, uint4 &b) {
    if (aligned32(p)) {
        ldg256(p, a, b);
    } else {
        a = reinterpret_cast<const uint4 *>(p)[0];
        b = reinterpret_cast<const uint4 *>(p)[1];
    }
}
__device__ __forceinline__ void stg_16ch(void *p, const uint32_t *w) {
    if (aligned32(p)) {
        stg256(p, w);
    } else {
        reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// 16 floats -> split bf16: hi[8] / lo[8] packed words (see AIVC_BF16X2)
__device__ __forceinline__ void split16(const float *v, uint32_t *hi, uint32_t *lo) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __bfloat1622float2(h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
        hi[i] = *reinterpret_cast<const uint32_t *>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t *>(&l2);
    }
}
__device__ __forceinline__ void add_packed_bf16x8(float *v, const uint4 &t) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[2 * j] += __uint_as_float(w[j] << 16);
        v[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
    }
}

__device__ __forceinline__ void load16(const FMap &m, bool vec, int y, int x, int ch0, int nvalid, float *v) {
    const size_t base = fm_index(m, y, x, ch0);
    if (vec && nvalid == 16) {
        if (m.dtype == AIVC_F32) {
            const float4 *p = reinterpret_cast<const float4 *>((const float *)m.data + base);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 t = p[i];
                v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
            }
        } else {
            uint4 p[2];
            ldg_16ch((const __nv_bfloat16 *)m.data + base, p[0], p[1]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint4 t = p[i];
                const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
                    v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
                }
            }
            if (m.dtype == AIVC_BF16X2) {                      // + lo half
                uint4 pl[2];
                ldg_16ch((const __nv_bfloat16 *)m.data + base + (m.c_stride >> 1), pl[0], pl[1]);
                add_packed_bf16x8(v, pl[0]);
                add_packed_bf16x8(v + 8, pl[1]);
            }
        }
    } else {
        for (int i = 0; i < 16; ++i) v[i] = (i < nvalid) ? fm_load(m, y, x, ch0 + i) : 0.f;
    }
}

__device__ __forceinline__ void store16(const FMap &m, bool vec, int y, int x, int ch0, int nvalid,
                                        const float *v) {
    const int p = m.pad;
    const int y0 = (y == 0) ? 0 : y + p, y1 = (y == m.h - 1) ? y + 2 * p : y + p;
    const int x0 = (x == 0) ? 0 : x + p, x1 = (x == m.w - 1) ? x + 2 * p : x + p;
    if (vec && nvalid == 16) {
        if (m.dtype == AIVC_F32) {
            for (int yy = y0; yy <= y1; ++yy)
                for (int xx = x0; xx <= x1; ++xx) {
                    float4 *q = reinterpret_cast<float4 *>(
                        (float *)m.data + ((size_t)yy * m.pitch + xx) * m.c_stride + m.c_off + ch0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
        } else if (m.dtype == AIVC_BF16X2) {
            uint32_t w[8], l[8];
            split16(v, w, l);
            for (int yy = y0; yy <= y1; ++yy)
                for (int xx = x0; xx <= x1; ++xx) {
                    uint4 *q = reinterpret_cast<uint4 *>(
                        (__nv_bfloat16 *)m.data + ((size_t)yy * m.pitch + xx) * m.c_stride + m.c_off + ch0);
                    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    uint4 *ql = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(q) + (m.c_stride >> 1));
                    ql[0] = make_uint4(l[0], l[1], l[2], l[3]);
                    ql[1] = make_uint4(l[4], l[5], l[6], l[7]);
                }
        } else {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack16x2(v[2 * i], v[2 * i + 1], m.dtype == AIVC_F16);
            for (int yy = y0; yy <= y1; ++yy)
                for (int xx = x0; xx <= x1; ++xx) {
                    uint4 *q = reinterpret_cast<uint4 *>(
                        (__nv_bfloat16 *)m.data + ((size_t)yy * m.pitch + xx) * m.c_stride + m.c_off + ch0);
                    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
        }
    } else {
        for (int yy = y0; yy <= y1; ++yy)
            for (int xx = x0; xx <= x1; ++xx)
                for (int i = 0; i < nvalid; ++i) fm_store_raw(m, yy, xx, ch0 + i, v[i]);
    }
}

// ------------------------------------------------------------------------------------ fused epilogue
// One accumulator row (= one output pixel, N channels) from TMEM to global memory:
//   out = post( act(acc + bias) * gate + residual ) * out_scale
// Everything that does not depend on the element (activation kind, presence of gate / residual /
// scale, vector alignment, border replication) is decided once per row, outside the channel loop;
// bias and scale are read from shared memory as broadcast float4.
// The maps are referenced, not copied: they live in the kernel's __grid_constant__ parameter block (constant bank), so
// an epilogue does not spend ~40 registers (or local-memory spills) on three FMap structs.  `ch0` is added to every
// channel index of out / res / gate (a work item that covers output channels [ch0, ch0 + N) of the stage).
struct EpiCtx {
    const FMap *outp, *resp, *gatep;
    int ch0;
    int post, act_channels;
    bool has_scale, out_vec, res_vec, gate_vec;
    bool precise;       // bf16x3 mode: accurate expf in the sigmoid (see epi_bias_act16)
};

__device__ __forceinline__ EpiCtx make_epi(const FMap &out, const FMap &res, const FMap &gate, int post,
                                           int act_channels, bool has_scale, bool precise = false) {
    EpiCtx c;
    c.outp = &out; c.resp = &res; c.gatep = &gate; c.ch0 = 0; c.post = post; c.act_channels = act_channels;
    c.has_scale = has_scale;
    c.precise = precise;
    c.out_vec = fmap_vec_ok(out);
    c.res_vec = res.data ? fmap_vec_ok(res) : false;
    c.gate_vec = gate.data ? fmap_vec_ok(gate) : false;
    return c;
}

// post activation of 16 values, the kind decided once (not per element)
__device__ __forceinline__ void post_apply16(int post, float *v) {
    if (post == AIVC_POST_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (post == AIVC_POST_LEAKY) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
    } else if (post == AIVC_POST_ROUND_CLAMP) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(rintf(v[i]), -256.f), 255.f);
    }
}

template <int ACT>
__device__ __forceinline__ void epi_bias_act16(float *v, const float *sbias, int j0, int act_channels,
                                               bool precise = false) {
    const float4 *b4 = reinterpret_cast<const float4 *>(sbias + j0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = b4[q];
        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
    }
    if (ACT == AIVC_ACT_SIGMOID && act_channels == 0) {
        // ex2.approx + rcp.approx (2^-22 relative) instead of expf + IEEE division; the split-bf16 mode keeps the
        // accurate expf (its operands carry 2^-17: the approximate reciprocal's 2 ulp are far below that, an IEEE
        // division's slow path is half of this epilogue's instructions)
        if (precise) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fdividef(1.f, 1.f + expf(-v[i]));
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fdividef(1.f, 1.f + __expf(-v[i]));
        }
    } else if (ACT != AIVC_ACT_NONE) {
        if (act_channels == 0 || j0 + 16 <= act_channels) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = act_apply(ACT, v[i]);
        } else if (j0 < act_channels) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (j0 + i < act_channels) v[i] = act_apply(ACT, v[i]);
        }
    }
}

// vectorised 16-channel store at a precomputed element offset (interior pixel, no replicas)
__device__ __forceinline__ void store16_at(const FMap &m, size_t elem, const float *v) {
    if (m.dtype == AIVC_F32) {
        float4 *q = reinterpret_cast<float4 *>((float *)m.data + elem);
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if (m.dtype == AIVC_BF16X2) {
        uint32_t w[8], l[8];
        split16(v, w, l);
        stg_16ch((__nv_bfloat16 *)m.data + elem, w);
        stg_16ch((__nv_bfloat16 *)m.data + elem + (m.c_stride >> 1), l);
    } else {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack16x2(v[2 * i], v[2 * i + 1], m.dtype == AIVC_F16);
        stg_16ch((__nv_bfloat16 *)m.data + elem, w);
    }
}

__device__ __forceinline__ void unpack_bf16x16(const uint4 *rb, float *r) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t w[4] = {rb[h].x, rb[h].y, rb[h].z, rb[h].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            r[8 * h + 2 * q] = __uint_as_float(w[q] << 16);
            r[8 * h + 2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
        }
    }
}

__device__ __forceinline__ void epi_tail16(float *v, const EpiCtx &c, const float *sscale, int oy, int ox,
                                           int j0, bool interior, size_t out_elem, const uint4 *rb = nullptr,
                                           const uint4 *gb = nullptr) {
    if (gb) {                       // gate row prefetched as packed bf16 (split maps: [2..3] = the lo halves)
        float g[16];
        unpack_bf16x16(gb, g);
        if (c.gatep->dtype == AIVC_BF16X2) { add_packed_bf16x8(g, gb[2]); add_packed_bf16x8(g + 8, gb[3]); }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= g[i];
    } else if (c.gatep->data) {
        float g[16];
        load16(*c.gatep, c.gate_vec, oy, ox, c.ch0 + j0, 16, g);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= g[i];
    }
    if (rb) {                       // residual row prefetched as packed bf16 (see res_prefetch)
        float r[16];
        unpack_bf16x16(rb, r);
        if (c.resp->dtype == AIVC_BF16X2) { add_packed_bf16x8(r, rb[2]); add_packed_bf16x8(r + 8, rb[3]); }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
    } else if (c.resp->data) {
        float r[16];
        load16(*c.resp, c.res_vec, oy, ox, c.ch0 + j0, 16, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
    }
    if (c.post != AIVC_POST_NONE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = post_apply(c.post, v[i]);
    }
    if (c.has_scale) {
        const float4 *s4 = reinterpret_cast<const float4 *>(sscale + j0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 s = s4[q];
            v[4 * q] *= s.x; v[4 * q + 1] *= s.y; v[4 * q + 2] *= s.z; v[4 * q + 3] *= s.w;
        }
    }
    if (interior && c.out_vec) store16_at(*c.outp, out_elem + j0, v);
    else store16(*c.outp, c.out_vec, oy, ox, c.ch0 + j0, 16, v);
}

template <bool X2 = false>
__device__ __forceinline__ void fetch16(const FMap &m, size_t elem, int j0, uint4 *rb) {
    ldg_16ch((const __nv_bfloat16 *)m.data + elem + j0, rb[0], rb[1]);
    if (X2 && m.dtype == AIVC_BF16X2)                          // lo halves of a split map
        ldg_16ch((const __nv_bfloat16 *)m.data + elem + j0 + (m.c_stride >> 1), rb[2], rb[3]);
}

// X2PIPE: also pipeline split-bf16 residual / gate rows (8 more registers per row in flight: only kernels with
// registers to spare -- the generic kernel -- ask for it; the 96-register persistent kernels load split rows in place)
template <int ACT, bool X2PIPE = false>
__device__ __forceinline__ void epi_row(uint32_t taddr, int N, const float *sbias, const float *sscale,
                                        const EpiCtx &c, int oy, int ox, bool valid, int c_begin = 0) {
    const bool interior = c.outp->pad == 0 || (oy > 0 && oy < c.outp->h - 1 && ox > 0 && ox < c.outp->w - 1);
    const size_t out_elem = valid ? fm_index(*c.outp, oy, ox, c.ch0) : 0;
    // bf16 residual / gate rows are fetched one chunk ahead, so the L2 latency of chunk j+1 hides behind
    // the TMEM load and arithmetic of chunk j
    constexpr int NB = X2PIPE ? 4 : 2;
    const bool pipe_res = valid && c.resp->data && c.res_vec && (c.resp->dtype == AIVC_BF16 || (X2PIPE && c.resp->dtype == AIVC_BF16X2));
    const bool pipe_gate = valid && c.gatep->data && c.gate_vec && (c.gatep->dtype == AIVC_BF16 || (X2PIPE && c.gatep->dtype == AIVC_BF16X2));
    const size_t res_elem = pipe_res ? fm_index(*c.resp, oy, ox, c.ch0) : 0;
    const size_t gate_elem = pipe_gate ? fm_index(*c.gatep, oy, ox, c.ch0) : 0;
    uint4 rb[NB], rn[NB], gb[NB], gn[NB];
    if (pipe_res) fetch16<X2PIPE>(*c.resp, res_elem, c_begin, rn);
    if (pipe_gate) fetch16<X2PIPE>(*c.gatep, gate_elem, c_begin, gn);
#pragma unroll 1
    for (int j0 = c_begin; j0 < N; j0 += 16) {
        if (pipe_res) {
#pragma unroll
            for (int i = 0; i < NB; ++i) rb[i] = rn[i];
            if (j0 + 16 < N) fetch16<X2PIPE>(*c.resp, res_elem, j0 + 16, rn);
        }
        if (pipe_gate) {
#pragma unroll
            for (int i = 0; i < NB; ++i) gb[i] = gn[i];
            if (j0 + 16 < N) fetch16<X2PIPE>(*c.gatep, gate_elem, j0 + 16, gn);
        }
        float v[16];
        tmem_ld16(taddr + (uint32_t)j0, v);
        epi_bias_act16<ACT>(v, sbias, j0, c.act_channels, c.precise);
        if (valid) epi_tail16(v, c, sscale, oy, ox, j0, interior, out_elem, pipe_res ? rb : nullptr, pipe_gate ? gb : nullptr);
    }
}

__device__ __forceinline__ void epi_row_dispatch(int act, uint32_t taddr, int N, const float *sbias,
                                                 const float *sscale, const EpiCtx &c, int oy, int ox,
                                                 bool valid, int c_begin = 0) {
    // channels [c_begin, N)
    switch (act) {
        case AIVC_ACT_LEAKY: epi_row<AIVC_ACT_LEAKY, true>(taddr, N, sbias, sscale, c, oy, ox, valid, c_begin); break;
        case AIVC_ACT_RELU: epi_row<AIVC_ACT_RELU, true>(taddr, N, sbias, sscale, c, oy, ox, valid, c_begin); break;
        case AIVC_ACT_SIGMOID: epi_row<AIVC_ACT_SIGMOID, true>(taddr, N, sbias, sscale, c, oy, ox, valid, c_begin); break;
        default: epi_row<AIVC_ACT_NONE, true>(taddr, N, sbias, sscale, c, oy, ox, valid, c_begin); break;
    }
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     (uint64_t)map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    // immediate barrier ids (1..4): a register id makes ptxas reserve all 16 hardware barriers
    switch (id) {
        case 1: asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); break;
        case 2: asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); break;
        case 3: asm volatile("bar.sync 3, %0;" ::"r"(nthreads) : "memory"); break;
        default: asm volatile("bar.sync 4, %0;" ::"r"(nthreads) : "memory"); break;
    }
}

// bias / scale staged once per CTA in shared memory (zeros / ones when absent)
__device__ __forceinline__ void stage_vec(float *dst, const float *src, int n, float fill, int tid, int nthreads) {
    for (int i = tid; i < n; i += nthreads) dst[i] = src ? src[i] : fill;
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

inline CUtensorMapSwizzle swizzle_for(int rowbytes) {
    return rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                           : (rowbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

inline int encode_map(CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
               const cuuint32_t *box, int rowbytes, const char *what) {
    EncodeTiledFn enc = get_encode();
    if (!enc) AIVC_FAIL("cuTensorMapEncodeTiled is not available from this driver");
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, ones,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(rowbytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) AIVC_FAIL("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}


// cudaFuncAttributeMaxDynamicSharedMemorySize sticks to (function, device): set it once, not per launch
int smem_attr_once(const void *kernel, int bytes);       // api.cu (mutex-protected table); 0 = ok

// launch with the programmatic-stream-serialization attribute (PDL)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace tcgen
