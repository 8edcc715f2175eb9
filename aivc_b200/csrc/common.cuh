// Shared device/host helpers for libaivc_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/aivc_b200.h"

// ---- error plumbing (thread-local message, int return codes; nothing throws) ----------
void aivc_set_error(const char *fmt, ...);
#define AIVC_FAIL(...)               \
    do {                             \
        aivc_set_error(__VA_ARGS__); \
        return 1;                    \
    } while (0)
#define AIVC_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) AIVC_FAIL("%s failed: %s", #expr, cudaGetErrorString(_e));  \
    } while (0)
extern unsigned long long g_aivc_launches;   // kernels launched by this library (this process)
#define AIVC_CHECK_LAUNCH(name)                                                            \
    do {                                                                                   \
        ++g_aivc_launches;                                                                 \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) AIVC_FAIL("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- feature-map addressing ---------------------------------------------------------
struct FMap {                 // device-side copy of aivc_fmap (same fields)
    void *data;
    int h, w, c, c_off, c_stride, pad, pitch, rows, dtype;
};

static inline FMap to_dev(const aivc_fmap &m) {
    FMap f;
    f.data = m.data; f.h = m.h; f.w = m.w; f.c = m.c; f.c_off = m.c_off; f.c_stride = m.c_stride;
    f.pad = m.pad; f.pitch = m.pitch; f.rows = m.rows; f.dtype = m.dtype;
    return f;
}

__device__ __forceinline__ size_t fm_index(const FMap &m, int y, int x, int ch) {
    return ((size_t)(y + m.pad) * m.pitch + (x + m.pad)) * m.c_stride + m.c_off + ch;
}

// Split-bf16 maps (AIVC_BF16X2, the bf16x3 precision mode): a pixel holds c_stride bf16 elements, the first
// half the leading 8 bits ("hi") of every channel and the second half the next 8 ("lo" = bf16(v - hi));
// the value is hi + lo (16 significant bits, fp32 exponent range).
__device__ __forceinline__ void bf16_split(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float fm_load(const FMap &m, int y, int x, int ch) {
    const size_t i = fm_index(m, y, x, ch);
    if (m.dtype == AIVC_F32) return ((const float *)m.data)[i];
    if (m.dtype == AIVC_F16) return __half2float(((const __half *)m.data)[i]);
    const __nv_bfloat16 *b = (const __nv_bfloat16 *)m.data;
    if (m.dtype == AIVC_BF16X2) return __bfloat162float(b[i]) + __bfloat162float(b[i + (m.c_stride >> 1)]);
    return __bfloat162float(b[i]);
}

__device__ __forceinline__ void fm_store_raw(const FMap &m, int yp, int xp, int ch, float v) {
    // (yp, xp) are coordinates in the padded buffer
    const size_t i = ((size_t)yp * m.pitch + xp) * m.c_stride + m.c_off + ch;
    if (m.dtype == AIVC_F32) ((float *)m.data)[i] = v;
    else if (m.dtype == AIVC_F16) ((__half *)m.data)[i] = __float2half_rn(v);
    else if (m.dtype == AIVC_BF16X2) {
        __nv_bfloat16 *b = (__nv_bfloat16 *)m.data;
        bf16_split(v, b[i], b[i + (m.c_stride >> 1)]);
    } else ((__nv_bfloat16 *)m.data)[i] = __float2bfloat16_rn(v);
}

// store element (y,x,ch) and its replicas in the border (edge pixels own their border copies)
__device__ __forceinline__ void fm_store(const FMap &m, int y, int x, int ch, float v) {
    const int p = m.pad;
    const int y0 = (y == 0) ? 0 : y + p, y1 = (y == m.h - 1) ? y + 2 * p : y + p;
    const int x0 = (x == 0) ? 0 : x + p, x1 = (x == m.w - 1) ? x + 2 * p : x + p;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) fm_store_raw(m, yy, xx, ch, v);
}

__device__ __forceinline__ float act_apply(int act, float v) {
    switch (act) {
        case AIVC_ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
        case AIVC_ACT_RELU: return fmaxf(v, 0.f);
        case AIVC_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ float post_apply(int post, float v) {
    switch (post) {
        case AIVC_POST_LEAKY: return v > 0.f ? v : 0.01f * v;
        case AIVC_POST_RELU: return fmaxf(v, 0.f);
        case AIVC_POST_ROUND_CLAMP: return fminf(fmaxf(rintf(v), -256.f), 255.f);
        default: return v;
    }
}

int validate_fmap(const aivc_fmap *m, const char *what);
