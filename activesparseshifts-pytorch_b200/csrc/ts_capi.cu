// ts_capi.cu -- the C ABI of include/torchshifts_b200.h: argument validation, host-side border
// logic, path selection (staged vs generic) and launch bookkeeping.  No torch types, no
// allocation, no synchronisation; every entry point only enqueues kernels on the caller's stream.
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ts_kernels.h"

namespace ts {

static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_forced_path{0};
static std::atomic<unsigned> g_tuning_epoch{0};     // bumped by ts_set_tuning: invalidates memoised plans
static thread_local int t_last_path = TS_PATH_NONE;
static thread_local char t_cuda_error[256] = "";

void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
void note_error(const char* text) { snprintf(t_cuda_error, sizeof(t_cuda_error), "%s", text); }

int check_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return TS_OK;
    snprintf(t_cuda_error, sizeof(t_cuda_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return TS_ERR_CUDA;
}

bool ensure_dynamic_smem(const void* func, size_t bytes) {
    // The attribute belongs to the (function, device) pair of the PROCESS: the memo is shared by all threads (autograd runs
    // backward kernels on its own threads -- a per-thread memo let one thread lower the limit another thread relied on)
    // and the limit only ever grows.
    struct Slot { const void* func; int dev; size_t bytes; };
    static std::mutex mu;
    static Slot seen[256];
    static int used = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(mu);
    int i = 0;
    for (; i < used; ++i)
        if (seen[i].func == func && seen[i].dev == dev) break;
    if (i < used && seen[i].bytes >= bytes) return true;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
    if (i == used) {
        if (used == 256) return true;         // table full: keep setting the attribute per launch for the rest
        ++used;
    }
    seen[i] = Slot{func, dev, bytes};
    return true;
}

static int cuda_fail(cudaError_t e) {
    snprintf(t_cuda_error, sizeof(t_cuda_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TS_ERR_NO_DEVICE : TS_ERR_CUDA;
}

static int sm_count(int* out) {
    static std::atomic<int> cached[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    if (dev >= 0 && dev < 64 && cached[dev].load() > 0) { *out = cached[dev].load(); return TS_OK; }
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    if (dev >= 0 && dev < 64) cached[dev].store(n);
    *out = n;
    return TS_OK;
}

static int make_geo(const ts_geometry* in, int padding, Geo* g) {
    if (!in) return TS_ERR_INVALID_ARGUMENT;
    if (in->dim < 1 || in->dim > 3) return TS_ERR_INVALID_ARGUMENT;
    if (padding < 0 || padding > 4) return TS_ERR_INVALID_ARGUMENT;
    if (in->N < 0 || in->C < 0) return TS_ERR_INVALID_ARGUMENT;
    g->dim = in->dim;
    g->pad = padding;
    g->N = in->N;
    g->C = in->C;
    long long ip = 1, op = 1;
    for (int a = 0; a < 3; ++a) {
        const long long S = a < in->dim ? in->size[a] : 1;
        const long long lb = a < in->dim ? in->lb[a] : 0;
        const long long rb = a < in->dim ? in->rb[a] : 1;
        if (S < 0 || lb < 0 || rb > S || rb < lb) return TS_ERR_INVALID_ARGUMENT;
        if (S >= 0x7fffffffLL) return TS_ERR_TOO_LARGE;
        g->S[a] = (int)S;
        g->lb[a] = (int)lb;
        g->OS[a] = (int)(rb - lb);
        ip *= S;
        op *= (rb - lb);
        if (ip >= 0x7fffffffLL) return TS_ERR_TOO_LARGE;
    }
    for (int a = 0; a < 5; ++a) g->xs[a] = (a < 2 + in->dim) ? in->x_stride[a] : 0;
    g->in_plane = ip;
    g->out_plane = op;
    return TS_OK;
}

static bool x_is_dense(const Geo& g) {
    long long expect = 1;
    for (int a = g.dim - 1; a >= 0; --a) {
        if (g.S[a] != 1 && g.xs[2 + a] != expect) return false;
        expect *= g.S[a];
    }
    if (g.C != 1 && g.xs[1] != expect) return false;
    expect *= g.C;
    if (g.N != 1 && g.xs[0] != expect) return false;
    return true;
}

static int elem_size(int dtype) {
    switch (dtype) {
    case TS_F32: return 4;
    case TS_F64: return 8;
    case TS_F16: case TS_BF16: return 2;
    }
    return 0;
}

}  // namespace ts

using namespace ts;

extern "C" {

int ts_abi_version(void) { return TS_ABI_VERSION; }
int ts_cuda_version(void) { return CUDART_VERSION; }

const char* ts_error_string(int status) {
    switch (status) {
    case TS_OK: return "ok";
    case TS_ERR_INVALID_ARGUMENT: return "invalid argument (dim must be 1..3, padding 0..4, sizes/borders consistent, pointers non-null)";
    case TS_ERR_UNSUPPORTED: return "unsupported request (no kernel for this element type / forced kernel path not applicable)";
    case TS_ERR_WORKSPACE: return "backward workspace missing, misaligned or smaller than ts_shift_backward_workspace_bytes()";
    case TS_ERR_TOO_LARGE: return "a single (n,c) plane must have fewer than 2^31 elements";
    case TS_ERR_BORDERS: return "borders produce a negative output dimension";
    case TS_ERR_CUDA: return "CUDA error (see ts_last_cuda_error())";
    case TS_ERR_NO_DEVICE: return "no CUDA device available: torchshifts-b200 has no CPU fallback";
    }
    return "unknown status";
}

const char* ts_last_cuda_error(void) { return t_cuda_error; }
int ts_last_kernel_path(void) { return t_last_path; }
int ts_set_kernel_path(int path) {
    if (path < 0 || path > 6 || path == TS_PATH_NHWC) return -1;
    return g_forced_path.exchange(path);
}
uint64_t ts_launch_count(void) { return (uint64_t)g_launches.load(); }

int ts_set_tuning(const char* spec) {
    if (!spec) return TS_ERR_INVALID_ARGUMENT;
    Tuning t = tuning();
    const char* p = spec;
    while (*p) {
        char key[32];
        int val = 0, n = 0;
        if (sscanf(p, " %31[a-z_]=%d%n", key, &val, &n) != 2) return TS_ERR_INVALID_ARGUMENT;
        if (!strcmp(key, "stages")) t.stages = val;
        else if (!strcmp(key, "stage_kb")) t.stage_kb = val;
        else if (!strcmp(key, "warps")) t.warps = val;
        else if (!strcmp(key, "ctas_per_sm")) t.ctas_per_sm = val;
        else if (!strcmp(key, "chunk_planes")) t.chunk_planes = val;
        else if (!strcmp(key, "tma_stages")) t.tma_stages = val;
        else if (!strcmp(key, "tma_ctas_per_sm")) t.tma_ctas_per_sm = val;
        else if (!strcmp(key, "tma_warps")) t.tma_warps = val;
        else if (!strcmp(key, "use_tma")) t.use_tma = val != 0;
        else if (!strcmp(key, "tma_stage_kb")) t.tma_stage_kb = val;
        else if (!strcmp(key, "nhwc_variant")) t.nhwc_variant = val;
        else if (!strcmp(key, "nhwc_ring_rows")) t.nhwc_ring_rows = val;
        else if (!strcmp(key, "nhwc_rows_warps")) t.nhwc_rows_warps = val;
        else if (!strcmp(key, "use_halo")) t.use_halo = val != 0;
        else if (!strcmp(key, "halo")) t.halo = val;
        else if (!strcmp(key, "halo_stages")) t.halo_stages = val;
        else if (!strcmp(key, "halo_warps")) t.halo_warps = val;
        else if (!strcmp(key, "halo_split")) t.halo_split = val != 0;
        else if (!strcmp(key, "halo_compact")) t.halo_compact = val != 0;
        else if (!strcmp(key, "halo_probe")) t.halo_probe = val;
        else if (!strcmp(key, "unit_order")) t.unit_order = val != 0;
        else if (!strcmp(key, "no_table")) t.no_table = val != 0;
        else if (!strcmp(key, "no_pdl")) t.no_pdl = val != 0;
        else if (!strcmp(key, "use_flat")) t.use_flat = val != 0;
        else if (!strcmp(key, "flat_variant")) t.flat_variant = val;
        else if (!strcmp(key, "flat_ctas")) t.flat_ctas = val;
        else if (!strcmp(key, "flat_stage_kb")) t.flat_stage_kb = val;
        else if (!strcmp(key, "flat_stages")) t.flat_stages = val;
        else if (!strcmp(key, "flat_warps")) t.flat_warps = val;
        else return TS_ERR_INVALID_ARGUMENT;
        p += n;
        while (*p == ',' || *p == ' ') ++p;
    }
    if (t.stages < 0 || t.stages > 8 || t.stage_kb < 0 || t.stage_kb > 220 || t.warps < 1 || t.warps > 31 ||
        t.ctas_per_sm < 1 || t.ctas_per_sm > 8 || t.chunk_planes < 0 || t.tma_stages < 0 || t.tma_stages > 32 ||
        t.tma_ctas_per_sm < 0 || t.tma_ctas_per_sm > 8 || t.tma_warps < 0 || t.tma_warps > 31 || t.tma_stage_kb < 0 ||
        t.tma_stage_kb > 220 || t.nhwc_variant < 0 || t.nhwc_variant > 3 || t.nhwc_ring_rows < 0 || t.nhwc_rows_warps < 0 || t.nhwc_rows_warps > 12 || t.halo < 0 || t.halo > 16 ||
        t.halo_stages < 0 || t.halo_stages > 32 || t.halo_warps < 0 || t.halo_warps > 15 || t.flat_ctas < 0 || t.flat_ctas > 8 ||
        t.flat_stage_kb < 0 || t.flat_stage_kb > 100 || t.flat_stages < 0 || t.flat_stages > 16 || t.flat_warps < 0 || t.flat_warps > 15)
        return TS_ERR_INVALID_ARGUMENT;
    tuning() = t;
    g_tuning_epoch.fetch_add(1);
    return TS_OK;
}

// ---- host logic ------------------------------------------------------------------------------
int ts_check_borders(int dim, const int64_t* sizes, const int64_t* user, int64_t lb[3], int64_t rb[3]) {
    if (dim < 1 || dim > 3 || !sizes || !lb || !rb) return TS_ERR_INVALID_ARGUMENT;
    for (int a = 0; a < 3; ++a) { lb[a] = 0; rb[a] = a < dim ? sizes[a] : 1; }
    if (!user) return TS_OK;
    for (int a = 0; a < dim; ++a) {
        const int size = (int)sizes[a];
        int r = size - (int)user[2 * a + 1];
        int l = (int)user[2 * a];
        if (r - l < 1) r = l + 1;
        if (l == size) { l = size - 1; r = l + 1; }
        if (r == 0) { l = 0; r = 1; }
        if (l < 0) l = 0;
        if (r > size) r = size;
        if (r - l < 0) return TS_ERR_BORDERS;
        lb[a] = l;
        rb[a] = r;
    }
    return TS_OK;
}

int ts_debug_remap(int padding, int len, int idx) { return len == 1 ? 0 : remap_literal(idx, len, padding); }

int ts_debug_remap_reduced(int padding, int len, int pos, int64_t shift, int plus) {
    const int s = reduce_shift((long long)shift, len, padding);
    return axis_index(pos - s + plus, len, padding);
}

void ts_debug_split_f32(int backward, int active, float w, int64_t* iw, float* dw) {
    long long i; float d;
    if (backward) split_backward<float>(w, active != 0, i, d); else split_forward<float>(w, active != 0, i, d);
    *iw = i; *dw = d;
}
void ts_debug_split_f64(int backward, int active, double w, int64_t* iw, double* dw) {
    long long i; double d;
    if (backward) split_backward<double>(w, active != 0, i, d); else split_forward<double>(w, active != 0, i, d);
    *iw = i; *dw = d;
}

// ---- device entry points ---------------------------------------------------------------------
int ts_shift_forward(const ts_geometry* gin, int dtype, int padding, int active, const void* x, const void* weights,
                     void* y, void* stream) {
    Geo g;
    int rc = make_geo(gin, padding, &g);
    if (rc != TS_OK) return rc;
    const int es = elem_size(dtype);
    if (!es) return TS_ERR_INVALID_ARGUMENT;
    if (g.N * g.C == 0 || g.out_plane == 0) return TS_OK;
    if (!x || !weights || !y) return TS_ERR_INVALID_ARGUMENT;
    int sms = 0;
    if ((rc = sm_count(&sms)) != TS_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int forced = g_forced_path.load();
    if ((forced == TS_PATH_NONE && tuning().use_tma) || forced == TS_PATH_TMA) {
        const TmaPlan tp = plan_tma(g, active ? 1 : 0, active, es, dtype, x_is_dense(g), 0ull, x, y, nullptr, sms);
        if (tp.ok) {
            t_last_path = TS_PATH_TMA;
            return active ? tma_active_forward(g, tp, x, weights, y, s) : tma_gather(g, tp, dtype, x, y, es, weights, 0, 0, s);
        }
    }
    if (forced == TS_PATH_TMA) return TS_ERR_UNSUPPORTED;
    if (active && ((forced == TS_PATH_NONE && tuning().use_halo) || forced == TS_PATH_HALO)) {
        const HaloPlan hp = plan_halo(g, 1, 1, dtype, x_is_dense(g), x, y, nullptr, sms, forced == TS_PATH_HALO);
        if (hp.ok) {
            t_last_path = TS_PATH_HALO;
            return halo_active_forward(g, hp, x, weights, y, s);
        }
    }
    if (forced == TS_PATH_HALO) return TS_ERR_UNSUPPORTED;
    if (!active && ((forced == TS_PATH_NONE && tuning().use_flat) || forced == TS_PATH_FLAT)) {
        const FlatPlan fp = plan_flat(g, es, x_is_dense(g), x, y, sms, forced == TS_PATH_FLAT);
        if (fp.ok) {
            t_last_path = TS_PATH_FLAT;
            return flat_gather(g, fp, dtype, x, y, 0ull, es, weights, 0, 0, s);
        }
    }
    if (forced == TS_PATH_FLAT) return TS_ERR_UNSUPPORTED;
    StagedPlan plan;
    plan.ok = false;
    if (forced != TS_PATH_GENERIC) plan = plan_staged(g, active ? 1 : 0, active, es, dtype, x_is_dense(g), x, y, nullptr, sms);
    if (forced == TS_PATH_STAGED && !plan.ok) return TS_ERR_UNSUPPORTED;
    if (plan.ok) {
        t_last_path = TS_PATH_STAGED;
        return active ? staged_active_forward(g, plan, dtype, x, weights, y, s)
                      : staged_gather(g, plan, dtype, x, y, 0ull, es, weights, 0, 0, s);
    }
    t_last_path = TS_PATH_GENERIC;
    return active ? generic_active_forward(g, dtype, x, weights, y, s)
                  : generic_gather(g, dtype, x, y, 0ull, es, weights, 0, 0, s);
}

int ts_shift2d_avgpool2_forward(const ts_geometry* gin, int dtype, int padding, int active, const void* x, const void* weights,
                                void* y_pooled, void* stream) {
    Geo g;
    int rc = make_geo(gin, padding, &g);
    if (rc != TS_OK) return rc;
    if (g.dim != 2 || dtype != TS_F32) return TS_ERR_UNSUPPORTED;
    if (g.N * g.C == 0 || g.out_plane == 0) return TS_OK;
    if (!x || !weights || !y_pooled) return TS_ERR_INVALID_ARGUMENT;
    int sms = 0;
    if ((rc = sm_count(&sms)) != TS_OK) return rc;
    const HaloPlan hp = plan_halo(g, active ? 1 : 0, active, dtype, x_is_dense(g), x, y_pooled, nullptr, sms, true);
    if (!hp.ok) return TS_ERR_UNSUPPORTED;
    t_last_path = TS_PATH_HALO;
    return halo_forward2d(g, hp, active, 1, x, weights, y_pooled, (cudaStream_t)stream);
}

size_t ts_shift_backward_workspace_bytes(const ts_geometry* gin, int dtype) {
    if (!gin) return 0;
    // memo of the last query of this thread: ts_shift_backward validates its workspace on every call
    static thread_local ts_geometry last_geo;
    static thread_local int last_dtype = -1;
    static thread_local unsigned last_epoch = 0;
    static thread_local size_t last_bytes = 0;
    const unsigned epoch = g_tuning_epoch.load();
    if (last_dtype == dtype && last_epoch == epoch && !memcmp(&last_geo, gin, sizeof(ts_geometry))) return last_bytes;
    Geo g;
    if (make_geo(gin, 0, &g) != TS_OK) return 0;
    size_t bytes = 16;
    if (g.N * g.C != 0) {
        int sms = 148;
        sm_count(&sms);
        const GenericBwdPlan gp = plan_generic_backward(g);
        size_t slots = (size_t)gp.units;
        // assume the staged / TMA paths may apply (pointer alignment is unknown here)
        for (int active = 0; active < 2; ++active) {
            const StagedPlan sp = plan_staged(g, 2, active, elem_size(dtype), dtype, true, nullptr, nullptr, nullptr, sms);
            if (sp.ok && (size_t)sp.slots > slots) slots = (size_t)sp.slots;
        }
        Geo gz = g;
        gz.pad = TS_PAD_ZEROS;
        for (int active = 0; active < 2; ++active) {
            const TmaPlan tp = plan_tma(gz, 2, active, elem_size(dtype), dtype, true, 0ull, nullptr, nullptr, nullptr, sms);
            if (tp.ok && (size_t)tp.slots > slots) slots = (size_t)tp.slots;
        }
        for (int pad = 0; pad < 2; ++pad)
            for (int active = 0; active < 2; ++active)
                for (int pool = 0; pool < 2; ++pool) {      // (pool: ts_shift2d_avgpool2_backward shares this query)
                    Geo gh = g;
                    gh.pad = pad;
                    const HaloPlan hp = plan_halo(gh, 2, active, dtype, true, nullptr, nullptr, nullptr, sms, true, pool != 0);
                    if (hp.ok && (size_t)hp.slots > slots) slots = (size_t)hp.slots;
                }
        bytes = slots * (size_t)(g.C * g.dim) * sizeof(double) + 16;
    }
    last_geo = *gin; last_dtype = dtype; last_epoch = epoch; last_bytes = bytes;
    return bytes;
}

static int backward_impl(const ts_geometry* gin, int dtype, int padding, int active, const void* grad, const void* x,
                         const void* weights, void* grad_input, void* grad_weight, void* workspace, size_t workspace_bytes,
                         const ts_peer_group* peers, void* stream) {
    Geo g;
    int rc = make_geo(gin, padding, &g);
    if (rc != TS_OK) return rc;
    const int es = elem_size(dtype);
    if (!es) return TS_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    if (g.C * g.dim == 0) return TS_OK;
    if (!grad_weight) return TS_ERR_INVALID_ARGUMENT;
    if (g.N == 0 || g.in_plane == 0) {
        if (peers) {      // an empty shard still takes part in the exchange: it contributes zeros
            const int outputs = (int)(g.C * g.dim);
            switch (dtype) {
            case TS_F32: return launch_reduce_partials<float>(nullptr, 0, outputs, grad_weight, peers, s);
            case TS_F16: return launch_reduce_partials<__half>(nullptr, 0, outputs, grad_weight, peers, s);
            default: return launch_reduce_partials<__nv_bfloat16>(nullptr, 0, outputs, grad_weight, peers, s);
            }
        }
        const cudaError_t e = cudaMemsetAsync(grad_weight, 0, (size_t)(g.C * g.dim) * es, s);
        return e == cudaSuccess ? TS_OK : cuda_fail(e);
    }
    if (!grad || !x || !weights || !grad_input) return TS_ERR_INVALID_ARGUMENT;
    if (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < ts_shift_backward_workspace_bytes(gin, dtype))
        return TS_ERR_WORKSPACE;
    int sms = 0;
    if ((rc = sm_count(&sms)) != TS_OK) return rc;
    const int forced = g_forced_path.load();
    // 3-D interpolating backward: the slab-streaming halo kernel (every slab staged once, windows of the previous slab in
    // registers) beats the TMA family's tile kernel under zeros padding as well (cfg4: 0.50 against 0.55 ms)
    if (forced == TS_PATH_NONE && tuning().use_halo && g.dim == 3 && active) {
        const HaloPlan hp = plan_halo(g, 2, active, dtype, x_is_dense(g), x, grad_input, grad, sms, false);
        if (hp.ok) {
            t_last_path = TS_PATH_HALO;
            return halo_backward(g, hp, active, grad, x, weights, grad_input, grad_weight, (double*)workspace, peers, s);
        }
    }
    if ((forced == TS_PATH_NONE && tuning().use_tma) || forced == TS_PATH_TMA) {
        const TmaPlan tp = plan_tma(g, 2, active, es, dtype, x_is_dense(g), 0ull, x, grad_input, grad, sms);
        if (tp.ok) {
            t_last_path = TS_PATH_TMA;
            return tma_backward(g, tp, active, grad, x, weights, grad_input, grad_weight, (double*)workspace, peers, s);
        }
    }
    if (forced == TS_PATH_TMA) return TS_ERR_UNSUPPORTED;
    if ((forced == TS_PATH_NONE && tuning().use_halo) || forced == TS_PATH_HALO) {
        const HaloPlan hp = plan_halo(g, 2, active, dtype, x_is_dense(g), x, grad_input, grad, sms, forced == TS_PATH_HALO);
        if (hp.ok) {
            t_last_path = TS_PATH_HALO;
            return halo_backward(g, hp, active, grad, x, weights, grad_input, grad_weight, (double*)workspace, peers, s);
        }
    }
    if (forced == TS_PATH_HALO || forced == TS_PATH_FLAT) return TS_ERR_UNSUPPORTED;
    StagedPlan plan;
    plan.ok = false;
    if (forced != TS_PATH_GENERIC) plan = plan_staged(g, 2, active, es, dtype, x_is_dense(g), x, grad_input, grad, sms);
    if (forced == TS_PATH_STAGED && !plan.ok) return TS_ERR_UNSUPPORTED;
    if (plan.ok) {
        t_last_path = TS_PATH_STAGED;
        return staged_backward(g, plan, dtype, active, grad, x, weights, grad_input, grad_weight, (double*)workspace, peers, s);
    }
    t_last_path = TS_PATH_GENERIC;
    return generic_backward(g, dtype, active, grad, x, weights, grad_input, grad_weight, (double*)workspace, peers, s);
}

int ts_shift2d_avgpool2_backward(const ts_geometry* gin, int dtype, int padding, int active, const void* grad_pooled, const void* x,
                                 const void* weights, void* grad_input, void* grad_weight, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    Geo g;
    int rc = make_geo(gin, padding, &g);
    if (rc != TS_OK) return rc;
    if (g.dim != 2 || dtype != TS_F32) return TS_ERR_UNSUPPORTED;
    if (g.N * g.C == 0 || g.in_plane == 0 || g.out_plane == 0) return TS_ERR_UNSUPPORTED;     // (empty tensors: the two-step path)
    if (!grad_pooled || !x || !weights || !grad_input || !grad_weight) return TS_ERR_INVALID_ARGUMENT;
    if (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < ts_shift_backward_workspace_bytes(gin, dtype))
        return TS_ERR_WORKSPACE;
    int sms = 0;
    if ((rc = sm_count(&sms)) != TS_OK) return rc;
    const HaloPlan hp = plan_halo(g, 2, active, dtype, x_is_dense(g), x, grad_input, grad_pooled, sms, true, true);
    if (!hp.ok) return TS_ERR_UNSUPPORTED;
    t_last_path = TS_PATH_HALO;
    return halo_backward(g, hp, active, grad_pooled, x, weights, grad_input, grad_weight, (double*)workspace, nullptr,
                         (cudaStream_t)stream, true);
}

int ts_shift_backward(const ts_geometry* gin, int dtype, int padding, int active, const void* grad, const void* x,
                      const void* weights, void* grad_input, void* grad_weight, void* workspace, size_t workspace_bytes,
                      void* stream) {
    return backward_impl(gin, dtype, padding, active, grad, x, weights, grad_input, grad_weight, workspace, workspace_bytes, nullptr,
                         stream);
}

int ts_shift_backward_allreduce(const ts_geometry* gin, int dtype, int padding, int active, const void* grad, const void* x,
                                const void* weights, void* grad_input, void* grad_weight, void* workspace, size_t workspace_bytes,
                                const ts_peer_group* peers, void* stream) {
    if (!peers || !gin) return TS_ERR_INVALID_ARGUMENT;
    if (peers->world < 1 || peers->world > 8 || peers->rank < 0 || peers->rank >= peers->world || !peers->state)
        return TS_ERR_INVALID_ARGUMENT;
    if (dtype != TS_F32 && dtype != TS_F16 && dtype != TS_BF16) return TS_ERR_UNSUPPORTED;   // contributions travel as fp32
    if (gin->C * gin->dim > peers->capacity || peers->capacity > 4096 || gin->C == 0) return TS_ERR_INVALID_ARGUMENT;
    for (int p = 0; p < peers->world; ++p)
        if (!peers->bufs[p] || ((uintptr_t)peers->bufs[p] & 7)) return TS_ERR_INVALID_ARGUMENT;
    return backward_impl(gin, dtype, padding, active, grad, x, weights, grad_input, grad_weight, workspace, workspace_bytes, peers,
                         stream);
}

int ts_qshift_forward(const ts_geometry* gin, int elem_bytes, int padding, int64_t zero_point, const void* xq,
                      const void* qweights, int qweight_kind, int64_t weight_zero_point, void* yq, void* stream) {
    Geo g;
    int rc = make_geo(gin, padding, &g);
    if (rc != TS_OK) return rc;
    if (elem_bytes != 1 && elem_bytes != 4) return TS_ERR_UNSUPPORTED;
    if (qweight_kind < TS_QW_U8 || qweight_kind > TS_QW_I32) return TS_ERR_INVALID_ARGUMENT;
    if (g.N * g.C == 0 || g.out_plane == 0) return TS_OK;
    if (!xq || !qweights || !yq) return TS_ERR_INVALID_ARGUMENT;
    int sms = 0;
    if ((rc = sm_count(&sms)) != TS_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned long long fill = elem_bytes == 1 ? (unsigned long long)((unsigned)zero_point & 0xffu)
                                                    : (unsigned long long)(uint32_t)(int32_t)zero_point;
    const int forced = g_forced_path.load();
    if ((forced == TS_PATH_NONE && tuning().use_tma) || forced == TS_PATH_TMA) {
        const TmaPlan tp = plan_tma(g, 0, 0, elem_bytes, -1, x_is_dense(g), fill, xq, yq, nullptr, sms);
        if (tp.ok) {
            t_last_path = TS_PATH_TMA;
            return tma_gather(g, tp, WK_QUANT, xq, yq, elem_bytes, qweights, qweight_kind, weight_zero_point, s);
        }
    }
    if (forced == TS_PATH_TMA || forced == TS_PATH_HALO) return TS_ERR_UNSUPPORTED;
    if ((forced == TS_PATH_NONE && tuning().use_flat) || forced == TS_PATH_FLAT) {
        const FlatPlan fp = plan_flat(g, elem_bytes, x_is_dense(g), xq, yq, sms, forced == TS_PATH_FLAT);
        if (fp.ok) {
            t_last_path = TS_PATH_FLAT;
            return flat_gather(g, fp, WK_QUANT, xq, yq, fill, elem_bytes, qweights, qweight_kind, weight_zero_point, s);
        }
    }
    if (forced == TS_PATH_FLAT) return TS_ERR_UNSUPPORTED;
    StagedPlan plan;
    plan.ok = false;
    if (forced != TS_PATH_GENERIC) plan = plan_staged(g, 0, 0, elem_bytes, -1, x_is_dense(g), xq, yq, nullptr, sms);
    if (forced == TS_PATH_STAGED && !plan.ok) return TS_ERR_UNSUPPORTED;
    if (plan.ok) {
        t_last_path = TS_PATH_STAGED;
        return staged_gather(g, plan, WK_QUANT, xq, yq, fill, elem_bytes, qweights, qweight_kind, weight_zero_point, s);
    }
    t_last_path = TS_PATH_GENERIC;
    return generic_gather(g, WK_QUANT, xq, yq, fill, elem_bytes, qweights, qweight_kind, weight_zero_point, s);
}

static int qshift_common(const ts_geometry* gin, int elem_bytes, int padding, int qweight_kind, Geo* g) {
    const int rc = make_geo(gin, padding, g);
    if (rc != TS_OK) return rc;
    if (elem_bytes != 1 && elem_bytes != 4) return TS_ERR_UNSUPPORTED;
    if (qweight_kind < TS_QW_U8 || qweight_kind > TS_QW_I32) return TS_ERR_INVALID_ARGUMENT;
    return TS_OK;
}

static unsigned long long qfill(int elem_bytes, int64_t zero_point) {
    return elem_bytes == 1 ? (unsigned long long)((unsigned)zero_point & 0xffu) : (unsigned long long)(uint32_t)(int32_t)zero_point;
}

int ts_qshift_forward_nhwc(const ts_geometry* gin, int elem_bytes, int padding, int64_t zero_point, const void* xq,
                           const void* qweights, int qweight_kind, int64_t weight_zero_point, void* yq, void* stream) {
    Geo g;
    int rc = qshift_common(gin, elem_bytes, padding, qweight_kind, &g);
    if (rc != TS_OK) return rc;
    if (g.N * g.C == 0 || g.out_plane == 0) return TS_OK;
    if (!xq || !qweights || !yq) return TS_ERR_INVALID_ARGUMENT;
    int sms = 0;
    if ((rc = sm_count(&sms)) != TS_OK) return rc;
    t_last_path = TS_PATH_NHWC;
    return nhwc_gather(g, xq, yq, qfill(elem_bytes, zero_point), elem_bytes, qweights, qweight_kind, weight_zero_point, sms, 0,
                       tuning().nhwc_variant, tuning().nhwc_ring_rows, false, (cudaStream_t)stream);
}

int ts_nhwc_to_nchw(const void* x, void* y, int64_t N, int64_t C, int64_t P, int elem_bytes, void* stream) {
    if (N < 0 || C < 0 || P < 0) return TS_ERR_INVALID_ARGUMENT;
    if (N == 0 || C == 0 || P == 0) return TS_OK;
    if (!x || !y) return TS_ERR_INVALID_ARGUMENT;
    int sms = 0;
    const int rc = sm_count(&sms);      // fails loudly without a device
    if (rc != TS_OK) return rc;
    return nhwc_to_planar(x, y, N, C, P, elem_bytes, (cudaStream_t)stream);
}

int ts_debug_nhwc_emulate(const ts_geometry* gin, int elem_bytes, int padding, int64_t zero_point, const void* xq_host,
                          const void* qweights_host, int qweight_kind, int64_t weight_zero_point, void* yq_host, int sm_count_,
                          int max_grid_x, int variant, int ring_rows) {
    Geo g;
    const int rc = qshift_common(gin, elem_bytes, padding, qweight_kind, &g);
    if (rc != TS_OK) return rc;
    if (g.N * g.C == 0 || g.out_plane == 0) return TS_OK;
    if (!xq_host || !qweights_host || !yq_host || sm_count_ < 1 || variant < 0 || variant > 2 || ring_rows < 0)
        return TS_ERR_INVALID_ARGUMENT;
    if ((double)g.N * (double)g.C * (double)g.in_plane > 4194304.0) return TS_ERR_TOO_LARGE;   // a test aid, not a CPU path
    return nhwc_gather(g, xq_host, yq_host, qfill(elem_bytes, zero_point), elem_bytes, qweights_host, qweight_kind, weight_zero_point,
                       sm_count_, max_grid_x, variant, ring_rows, true, nullptr);
}

}  // extern "C"
