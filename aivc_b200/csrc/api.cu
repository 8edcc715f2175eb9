// C-ABI glue: error reporting, argument validation, engine dispatch.
#include <stdarg.h>
#include "common.cuh"
#include "laplace_cdf.h"

int conv_simt_run(const aivc_conv_op *op, cudaStream_t st);
int conv_tc_run(const aivc_conv_op *op, cudaStream_t st);
int col2im_tconv_run(const aivc_conv_op *op, cudaStream_t st);
int space_to_depth_run(const aivc_conv_op *op, cudaStream_t st);

static thread_local char g_err[512] = "";
unsigned long long g_aivc_launches = 0;

// ---- optional per-stage timing (bench.py's roofline leg): CUDA events around every conv stage
#include <vector>
struct StageRec { cudaEvent_t a, b; int engine; double flops; int kind, k, stride, cin, cout, h, w, act, kclass; };
int g_aivc_kernel_class = 0;     // set by the launchers: which kernel the last stage ran on (AIVC_KC_*)
static std::vector<StageRec> g_prof;
static bool g_prof_on = false;

static double stage_flops(const aivc_conv_op *op) {
    if (op->alg_flops > 0.0) return op->alg_flops;
    if (op->kind >= 2) return 0.0;              // col2im / space-to-depth: data movement only
    const double px = op->kind == 0 ? (double)op->out.h * op->out.w : (double)op->in.h * op->in.w;
    double f = 2.0 * op->k * op->k * op->in.c * op->out.c * px;
    if (op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN) f += 2.0 * op->out.c * op->out.c * (double)op->out.h * op->out.w;
    return f;
}

void aivc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int validate_fmap(const aivc_fmap *m, const char *what) {
    if (!m || !m->data) AIVC_FAIL("%s: null feature map", what);
    if (m->h <= 0 || m->w <= 0 || m->c <= 0) AIVC_FAIL("%s: empty feature map %dx%dx%d", what, m->h, m->w, m->c);
    if (m->c_off < 0 || m->c_off + m->c > m->c_stride) AIVC_FAIL("%s: channel view [%d,%d) exceeds pixel stride %d", what, m->c_off, m->c_off + m->c, m->c_stride);
    if (m->pad < 0 || m->pitch < m->w + 2 * m->pad || m->rows < m->h + 2 * m->pad) AIVC_FAIL("%s: pitch/rows smaller than padded size", what);
    if (m->dtype != AIVC_F32 && m->dtype != AIVC_BF16 && m->dtype != AIVC_F16 && m->dtype != AIVC_BF16X2) AIVC_FAIL("%s: unknown dtype %d", what, m->dtype);
    if (m->dtype == AIVC_BF16X2 && ((m->c_stride & 1) || m->c_off + m->c > m->c_stride / 2)) AIVC_FAIL("%s: split-bf16 view [%d,%d) exceeds half the pixel stride %d", what, m->c_off, m->c_off + m->c, m->c_stride);
    return 0;
}

static int validate_conv(const aivc_conv_op *op) {
    if (!op) AIVC_FAIL("conv2d_fused: null op");
    if (validate_fmap(&op->in, "conv in") || validate_fmap(&op->out, "conv out")) return 1;
    if (!op->weight && op->kind < 2) AIVC_FAIL("conv2d_fused: null weight");
    if (op->k < 1 || (op->k & 1) == 0) AIVC_FAIL("conv2d_fused: kernel size %d must be odd", op->k);
    if (op->kind == 0) {
        if (op->stride != 1 && op->stride != 2) AIVC_FAIL("conv2d_fused: stride %d unsupported", op->stride);
        const int eh = (op->in.h + op->stride - 1) / op->stride, ew = (op->in.w + op->stride - 1) / op->stride;
        if (op->out.h != eh || op->out.w != ew) AIVC_FAIL("conv2d_fused: output %dx%d, expected %dx%d", op->out.h, op->out.w, eh, ew);
    } else if (op->kind == 1) {
        if (op->stride != 2) AIVC_FAIL("conv2d_fused: transposed conv needs stride 2");
        if (op->out.h != 2 * op->in.h || op->out.w != 2 * op->in.w) AIVC_FAIL("conv2d_fused: transposed output must be exactly 2x");
    } else if (op->kind == 2 || op->kind == 3) {
        return 0;                                   // col2im / space-to-depth: checked by their launchers
    } else {
        AIVC_FAIL("conv2d_fused: unknown kind %d", op->kind);
    }
    if ((op->act == AIVC_ACT_GDN || op->act == AIVC_ACT_IGDN) && (!op->gdn_beta || !op->gdn_gamma)) AIVC_FAIL("conv2d_fused: GDN without parameters");
    if (op->residual.data) {
        if (validate_fmap(&op->residual, "conv residual")) return 1;
        if (op->residual.h != op->out.h || op->residual.w != op->out.w || op->residual.c != op->out.c) AIVC_FAIL("conv2d_fused: residual shape mismatch");
    }
    if (op->gate.data) {
        if (validate_fmap(&op->gate, "conv gate")) return 1;
        if (op->gate.h != op->out.h || op->gate.w != op->out.w || op->gate.c != op->out.c) AIVC_FAIL("conv2d_fused: gate shape mismatch");
    }
    return 0;
}

extern "C" {

int aivc_abi_version(void) { return AIVC_ABI_VERSION; }
const char *aivc_last_error(void) { return g_err; }

int aivc_conv2d_fused(const aivc_conv_op *op, void *stream) {
    if (validate_conv(op)) return 1;
    if (op->kind < 2 && op->engine != AIVC_ENGINE_SIMT && op->engine != AIVC_ENGINE_TC && op->engine != AIVC_ENGINE_TC_X3) AIVC_FAIL("conv2d_fused: unknown engine %d", op->engine);
    StageRec r;
    if (g_prof_on) {
        AIVC_CHECK_CUDA(cudaEventCreate(&r.a));
        AIVC_CHECK_CUDA(cudaEventCreate(&r.b));
        r.engine = op->kind >= 2 ? AIVC_ENGINE_SIMT : op->engine;
        r.flops = stage_flops(op);
        r.kind = op->kind; r.k = op->k; r.stride = op->stride; r.cin = op->in.c; r.cout = op->out.c;
        r.h = op->out.h; r.w = op->out.w; r.act = op->act;
        AIVC_CHECK_CUDA(cudaEventRecord(r.a, (cudaStream_t)stream));
    }
    const int rc = op->kind == 3 ? space_to_depth_run(op, (cudaStream_t)stream)
                 : op->kind == 2 ? col2im_tconv_run(op, (cudaStream_t)stream)
                 : op->engine == AIVC_ENGINE_SIMT ? conv_simt_run(op, (cudaStream_t)stream)
                                                  : conv_tc_run(op, (cudaStream_t)stream);
    if (g_prof_on) {
        AIVC_CHECK_CUDA(cudaEventRecord(r.b, (cudaStream_t)stream));
        r.kclass = g_aivc_kernel_class;
        g_prof.push_back(r);
    }
    return rc;
}

unsigned long long aivc_launch_count(void) { return g_aivc_launches; }

int aivc_profile_enable(int on) {
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
    return 0;
}

// one CSV line per recorded stage: engine,kind,k,stride,cin,cout,out_h,out_w,act,flops,ms
int aivc_profile_dump(const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) AIVC_FAIL("profile_dump: cannot open %s", path);
    fprintf(f, "engine,kind,k,stride,cin,cout,out_h,out_w,act,flops,ms,kernel\n");
    for (auto &r : g_prof) {
        AIVC_CHECK_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        AIVC_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        fprintf(f, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%.0f,%.6f,%d\n", r.engine, r.kind, r.k, r.stride, r.cin, r.cout, r.h, r.w, r.act, r.flops, ms, r.kclass);
    }
    fclose(f);
    return 0;
}

// out[0..5] = tc_ms, tc_flops, tc_stages, simt_ms, simt_flops, simt_stages (device must be idle or
// the caller synchronised; this call synchronises on the last event of every stage)
int aivc_profile_read(double *out) {
    for (int i = 0; i < 6; ++i) out[i] = 0.0;
    for (auto &r : g_prof) {
        AIVC_CHECK_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        AIVC_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        const int o = r.engine != AIVC_ENGINE_SIMT ? 0 : 3;
        out[o] += ms; out[o + 1] += r.flops; out[o + 2] += 1.0;
    }
    return 0;
}

// per kernel class k < n: out[3k] = ms, out[3k+1] = algorithmic flops, out[3k+2] = launches
int aivc_profile_read_classes(double *out, int n) {
    for (int i = 0; i < 3 * n; ++i) out[i] = 0.0;
    for (auto &r : g_prof) {
        AIVC_CHECK_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        AIVC_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.kclass < 0 || r.kclass >= n) continue;
        out[3 * r.kclass] += ms; out[3 * r.kclass + 1] += r.flops; out[3 * r.kclass + 2] += 1.0;
    }
    return 0;
}

}  // extern "C"

#include <mutex>
namespace tcgen {
int smem_attr_once(const void *kernel, int bytes) {
    struct Entry { const void *k; unsigned devmask; };
    static Entry tab[64];
    static int n = 0;
    static std::mutex mu;
    int dev = 0;
    AIVC_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    Entry *e = nullptr;
    for (int i = 0; i < n; ++i)
        if (tab[i].k == kernel) { e = &tab[i]; break; }
    if (e && ((e->devmask >> dev) & 1u)) return 0;
    AIVC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (!e && n < 64) { e = &tab[n++]; e->k = kernel; e->devmask = 0; }
    if (e && dev < 32) e->devmask |= 1u << dev;
    return 0;
}
}  // namespace tcgen

extern "C" {

// Two-lane execution of a plan: the side stream + fork / join events belong to the CALLER'S stream, so that
// plans enqueued on different streams (two frames in flight, the encoder's shortcut transform next to the analysis
// chain) never funnel their side stages through one shared stream.  Small table per process, mutex-protected.
struct Lanes { int dev = -1; cudaStream_t main = nullptr, side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static Lanes g_lanes[64];
static int g_nlanes = 0;
static std::mutex g_lanes_mu;

static int lanes_for_stream(cudaStream_t main_s, Lanes **out) {
    int dev = 0;
    AIVC_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_lanes_mu);
    for (int i = 0; i < g_nlanes; ++i)
        if (g_lanes[i].dev == dev && g_lanes[i].main == main_s) { *out = &g_lanes[i]; return 0; }
    if (g_nlanes == 64) AIVC_FAIL("more than 64 (device, stream) pairs run two-lane plans");
    Lanes &l = g_lanes[g_nlanes];
    l.dev = dev; l.main = main_s;
    AIVC_CHECK_CUDA(cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking));
    AIVC_CHECK_CUDA(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
    AIVC_CHECK_CUDA(cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming));
    ++g_nlanes;
    *out = &l;
    return 0;
}

// `given`: side lane to use (graph capture brings its own, short-lived one); NULL: the caller stream's lane
static int fused_seq(const aivc_conv_op *ops, int n, cudaStream_t main_s, Lanes *given) {
    void *stream = (void *)main_s;
    Lanes *l = given;
    bool side_dirty = false;
    for (int i = 0; i < n; ++i) {
        const int flags = g_prof_on ? 0 : (ops[i].flags & (AIVC_OP_LANE1 | AIVC_OP_FORK | AIVC_OP_JOIN));   // per-stage timing needs kernels one at a time
        if (flags && !l && lanes_for_stream(main_s, &l)) return 1;
        if (flags & AIVC_OP_FORK) {
            AIVC_CHECK_CUDA(cudaEventRecord(l->fork, main_s));
            AIVC_CHECK_CUDA(cudaStreamWaitEvent(l->side, l->fork, 0));
        }
        if ((flags & AIVC_OP_JOIN) && side_dirty) {
            AIVC_CHECK_CUDA(cudaEventRecord(l->join, l->side));
            AIVC_CHECK_CUDA(cudaStreamWaitEvent(main_s, l->join, 0));
            side_dirty = false;
        }
        const bool on_side = (flags & AIVC_OP_LANE1) != 0;
        if (on_side) side_dirty = true;
        if (aivc_conv2d_fused(ops + i, on_side ? (void *)l->side : stream)) {
            char tmp[400];
            snprintf(tmp, sizeof(tmp), "%.399s", g_err);
            AIVC_FAIL("stage %d/%d: %s", i, n, tmp);
        }
    }
    if (side_dirty) {                              // never leave work the caller cannot see
        AIVC_CHECK_CUDA(cudaEventRecord(l->join, l->side));
        AIVC_CHECK_CUDA(cudaStreamWaitEvent(main_s, l->join, 0));
    }
    return 0;
}

int aivc_conv2d_fused_seq(const aivc_conv_op *ops, int n, void *stream) {
    return fused_seq(ops, n, (cudaStream_t)stream, nullptr);
}

// ---- a transform as ONE CUDA graph: the n kernel launches (with their programmatic-dependent-launch edges and the
// fork / join of the two-lane attention branches) are captured once; replaying them costs one cudaGraphLaunch instead
// of n launches and 3 n tensor-map encodes on the host, and the device schedules the nodes without host gaps.
int aivc_plan_graph_create(const aivc_conv_op *ops, int n, void **graph_exec) {
    if (!ops || n <= 0 || !graph_exec) AIVC_FAIL("plan_graph_create: bad arguments");
    if (g_prof_on) AIVC_FAIL("plan_graph_create: per-stage profiling is on (stages are timed one by one, not as a graph)");
    *graph_exec = nullptr;
    cudaStream_t cs = nullptr;
    Lanes tmp;
    AIVC_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    AIVC_CHECK_CUDA(cudaStreamCreateWithFlags(&tmp.side, cudaStreamNonBlocking));
    AIVC_CHECK_CUDA(cudaEventCreateWithFlags(&tmp.fork, cudaEventDisableTiming));
    AIVC_CHECK_CUDA(cudaEventCreateWithFlags(&tmp.join, cudaEventDisableTiming));
    int rc = 0;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        rc = fused_seq(ops, n, cs, &tmp);
        e = cudaStreamEndCapture(cs, &g);                       // (also ends a capture that failed half way)
    }
    if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&ex, g, 0);
    if (g) cudaGraphDestroy(g);
    cudaEventDestroy(tmp.fork); cudaEventDestroy(tmp.join);
    cudaStreamDestroy(tmp.side); cudaStreamDestroy(cs);
    if (rc) return 1;                                           // (message set by the failing stage)
    if (e != cudaSuccess) { cudaGetLastError(); AIVC_FAIL("plan_graph_create: %s", cudaGetErrorString(e)); }
    *graph_exec = (void *)ex;
    return 0;
}

int aivc_plan_graph_launch(void *graph_exec, void *stream) {
    if (!graph_exec) AIVC_FAIL("plan_graph_launch: null graph");
    AIVC_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return 0;
}

int aivc_plan_graph_destroy(void *graph_exec) {
    if (graph_exec) AIVC_CHECK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return 0;
}

uint32_t aivc_laplace_cdf_int_host(float b, int i) { return aivc_laplace_cdf_int(b, i); }
float aivc_sigma_from_logvar_host(float v) { return aivc_sigma_from_logvar(v); }

}  // extern "C"
