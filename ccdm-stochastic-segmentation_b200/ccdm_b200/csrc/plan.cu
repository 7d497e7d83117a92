// C ABI surface + the "plan": one reverse step (UNet forward + posterior + draw,
// reference diffusion_denoising.py:189-212) as a fixed sequence of fused kernel
// launches, captured once into a CUDA graph and replayed for every t.  Per-step
// values (timestep-embedding row, alpha_t, cumalpha_{t-1}, draw mode, Philox draw
// index) are read by the kernels from a device-resident step table indexed by a
// device-side counter that the head kernel advances, so the graph never needs to
// be re-captured or patched between steps.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace ccdm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// CCDM_PDL: 0 = off, 1 = every launch, 2 = only the launches of ops on small maps (<= kPdlSmallPixels pixels per
// sample: the latency-bound 8x8 .. 32x32 levels), selected per op by launch_any through g_pdl_hint.
static thread_local int g_pdl_hint = 0;
constexpr int kPdlSmallPixels = 1024;
static int pdl_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CCDM_PDL");
        v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
    }
    return v;
}
bool pdl_enabled() { return pdl_mode() == 1 || (pdl_mode() == 2 && g_pdl_hint); }

static int launch_any(const ccdm_op &op, cudaStream_t s) {
    g_pdl_hint = (op.kind == CCDM_OP_CONV || op.kind == CCDM_OP_ATTENTION) &&
                 (op.kind == CCDM_OP_ATTENTION ? op.Hin * op.Win : op.Hout * op.Wout) <= kPdlSmallPixels;
    switch (op.kind) {
        case CCDM_OP_INPUT_CONV:
        case CCDM_OP_CONV: return launch_conv(op, s);
        case CCDM_OP_ATTENTION: return launch_attention(op, s);
        case CCDM_OP_HEAD: return launch_head(op, s);
        case CCDM_OP_ENCODE_INPUT: return launch_encode_input(op, s);
        default: CCDM_FAIL(-2, "unknown op kind %d", op.kind);
    }
}

}  // namespace ccdm

using namespace ccdm;

struct ccdm_plan {
    std::vector<ccdm_op> ops;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

extern "C" int ccdm_abi_version(void) { return CCDM_ABI_VERSION; }

extern "C" const char *ccdm_last_error(void) { return g_err; }

extern "C" size_t ccdm_sizeof_op(void) { return sizeof(ccdm_op); }
extern "C" size_t ccdm_sizeof_step_entry(void) { return sizeof(ccdm_step_entry); }
namespace ccdm { size_t conv_part_floats(int B, int Hout, int Wout, int Cout); }
extern "C" size_t ccdm_conv_part_floats(int B, int Hout, int Wout, int Cout) { return ccdm::conv_part_floats(B, Hout, Wout, Cout); }

namespace ccdm { bool conv_uses_tc(const ccdm_op &op); size_t op_part_floats(const ccdm_op &op); int conv_tc_nt(int Cout, int taps, int x3); }
extern "C" int ccdm_conv_uses_tc(const ccdm_op *op) { return op && ccdm::conv_uses_tc(*op) ? 1 : 0; }
extern "C" int ccdm_conv_tc_nt(int Cout, int taps, int x3) { return ccdm::conv_tc_nt(Cout, taps, x3); }
namespace ccdm { int conv_tma_config(const ccdm_op &op, int32_t *out); bool conv_uses_tma(const ccdm_op &op); }
extern "C" int ccdm_conv_tc_config(const ccdm_op *op, int32_t *out16) {
    if (!op || !out16 || !ccdm::conv_uses_tma(*op)) return -1;
    return ccdm::conv_tma_config(*op, out16);
}
extern "C" int ccdm_conv_uses_tma(const ccdm_op *op) { return op && ccdm::conv_uses_tma(*op) ? 1 : 0; }
namespace ccdm { int conv_tma_stat_layout(const ccdm_op &op, int32_t *out5); }
extern "C" int ccdm_conv_stat_layout(const ccdm_op *op, int32_t *out5) {
    if (!op || !out5 || !ccdm::conv_uses_tc(*op)) return -1;
    return ccdm::conv_tma_stat_layout(*op, out5);
}
extern "C" size_t ccdm_op_part_floats(const ccdm_op *op) { return op ? ccdm::op_part_floats(*op) : 0; }

extern "C" int ccdm_check_device(void) {
    int dev = 0;
    CCDM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CCDM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) CCDM_FAIL(-4, "libccdm_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return 0;
}

extern "C" int ccdm_launch_op(const ccdm_op *op, void *stream) {
    if (!op) CCDM_FAIL(-1, "null op");
    return launch_any(*op, (cudaStream_t)stream);
}

extern "C" ccdm_plan *ccdm_plan_create(const ccdm_op *ops, int n_ops) {
    if (!ops || n_ops <= 0) {
        set_error("plan_create: empty program");
        return nullptr;
    }
    ccdm_plan *p = new ccdm_plan();
    p->ops.assign(ops, ops + n_ops);
    return p;
}

static void drop_graph(ccdm_plan *plan) {
    if (plan->exec) cudaGraphExecDestroy(plan->exec);
    if (plan->graph) cudaGraphDestroy(plan->graph);
    plan->exec = nullptr;
    plan->graph = nullptr;
}

extern "C" void ccdm_plan_destroy(ccdm_plan *plan) {
    if (!plan) return;
    drop_graph(plan);
    delete plan;
}

extern "C" int ccdm_plan_num_launches(const ccdm_plan *plan) { return plan ? int(plan->ops.size()) : 0; }

extern "C" int ccdm_plan_set_noise(ccdm_plan *plan, int noise_mode, uint64_t seed, int32_t sample0, const float *noise,
                                   float *noise_out) {
    if (!plan) CCDM_FAIL(-1, "null plan");
    bool changed = false;
    for (auto &op : plan->ops) {
        if (op.kind != CCDM_OP_HEAD) continue;
        if (op.noise_mode != noise_mode || op.seed != seed || op.sample0 != sample0 || op.noise != (uint64_t)noise ||
            op.noise_out != (uint64_t)noise_out)
            changed = true;
        op.noise_mode = noise_mode;
        op.seed = seed;
        op.sample0 = sample0;
        op.noise = (uint64_t)noise;
        op.noise_out = (uint64_t)noise_out;
    }
    if (changed) drop_graph(plan);
    return 0;
}

static int launch_all(ccdm_plan *plan, cudaStream_t s) {
    for (size_t i = 0; i < plan->ops.size(); ++i) {
        int rc = launch_any(plan->ops[i], s);
        if (rc != 0) {
            char msg[400];
            strncpy(msg, g_err, sizeof(msg) - 1);
            msg[sizeof(msg) - 1] = 0;
            CCDM_FAIL(rc, "op %d (kind %d): %s", int(i), plan->ops[i].kind, msg);
        }
    }
    return 0;
}

static int capture(ccdm_plan *plan, cudaStream_t s) {
    // Make sure every kernel's attributes are set outside capture by a dry launch
    // sequence is NOT needed: cudaFuncSetAttribute is legal during capture.
    CCDM_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = launch_all(plan, s);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (rc != 0) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) CCDM_FAIL(-100, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    plan->graph = g;
    CCDM_CUDA(cudaGraphInstantiate(&plan->exec, plan->graph, 0));
    return 0;
}

extern "C" int ccdm_plan_step(ccdm_plan *plan, int use_graph, void *stream) {
    if (!plan) CCDM_FAIL(-1, "null plan");
    cudaStream_t s = (cudaStream_t)stream;
    if (!use_graph) return launch_all(plan, s);
    if (!plan->exec) {
        int rc = capture(plan, s);
        if (rc != 0) return rc;
    }
    CCDM_CUDA(cudaGraphLaunch(plan->exec, s));
    return 0;
}

extern "C" int ccdm_plan_run(ccdm_plan *plan, int n_steps, int use_graph, void *stream) {
    for (int i = 0; i < n_steps; ++i) {
        int rc = ccdm_plan_step(plan, use_graph, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

// Per-op device time of one reverse step: every op is captured into its own one-node CUDA graph and
// replayed `iters` times between two events, so the figure is free of host launch overhead (tensor-map
// encoding, ctypes) even for launches of a few microseconds.  Inputs of small ops are L2-resident on
// the replays, as they are in the real chain where the producer has just written them.  The device step
// counter is restored afterwards (the head op advances it).  Measurement aid for bench.py, not on the
// product path.
extern "C" int ccdm_plan_profile(ccdm_plan *plan, int iters, float *ms_per_op, void *stream) {
    if (!plan || !ms_per_op || iters <= 0) CCDM_FAIL(-1, "plan_profile: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CCDM_CUDA(cudaEventCreate(&e0));
    CCDM_CUDA(cudaEventCreate(&e1));
    int rc = 0;
    int saved_step = 0;
    const int *step_ptr = nullptr;
    for (auto &op : plan->ops)
        if (op.step_ptr) step_ptr = (const int *)op.step_ptr;
    if (step_ptr) {
        CCDM_CUDA(cudaMemcpyAsync(&saved_step, step_ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        CCDM_CUDA(cudaStreamSynchronize(s));
    }
    for (size_t i = 0; i < plan->ops.size() && rc == 0; ++i) {
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ex = nullptr;
        CCDM_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        rc = launch_any(plan->ops[i], s);
        cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc != 0 || e != cudaSuccess) {
            if (g) cudaGraphDestroy(g);
            if (rc == 0) {
                set_error("plan_profile: capture of op %d failed: %s", int(i), cudaGetErrorString(e));
                rc = -100;
            }
            break;
        }
        CCDM_CUDA(cudaGraphInstantiate(&ex, g, 0));
        CCDM_CUDA(cudaGraphLaunch(ex, s));  // warm-up
        CCDM_CUDA(cudaEventRecord(e0, s));
        for (int k = 0; k < iters; ++k) CCDM_CUDA(cudaGraphLaunch(ex, s));
        CCDM_CUDA(cudaEventRecord(e1, s));
        CCDM_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CCDM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        ms_per_op[i] = ms / float(iters);
        cudaGraphExecDestroy(ex);
        cudaGraphDestroy(g);
        if (step_ptr) CCDM_CUDA(cudaMemcpyAsync((void *)step_ptr, &saved_step, sizeof(int), cudaMemcpyHostToDevice, s));
    }
    if (step_ptr) cudaStreamSynchronize(s);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

// Debug aid: nanosecond time stamps (%globaltimer) that CTA 0 of the most recent conv_tma launch took at
// its role milestones (conv_tma.cu, kTrace*).  Returns the number of slots copied.
namespace ccdm { int conv_tma_read_trace(unsigned long long *out, int n); }
extern "C" int ccdm_debug_conv_trace(unsigned long long *out, int n) { return ccdm::conv_tma_read_trace(out, n); }
