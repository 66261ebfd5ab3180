// ts_nhwc.cu -- channels-last (NHWC / NDHWC) integer gather: input AND output keep the channel axis
// innermost, so a quantized channels-last pipeline pays one read and one write of the tensor instead
// of the three passes (to-NCHW copy, NCHW kernel, to-NHWC copy) the planar families would need.
//
// Semantics: reference body  ops/kernels/shifts_kernels.h:574-624 (shift_forward_kernel_nhwdc_q),
//            driver          ops/quantized/shifts_quantized.cpp:107-130 (output allocated in the
//                            input's memory format, :119-122).
//
// Layout of the work.  In NHWC the per-channel shift is a per-lane gather: the channel vector of one
// output pixel takes each of its channels from a different input pixel.  Consecutive lanes own
// consecutive 4-byte words of the channel vector (4 channels of a 1-byte type, 1 channel of a 4-byte
// type), so every warp store is one contiguous 128-byte line and every warp load touches at most
// (distinct shifts) x 4 sectors, all of them re-used by the neighbouring pixels through L1/L2.
// A CTA owns a few consecutive output rows of one image and its threads keep their channel group for
// the whole CTA lifetime: the shifts, the remapped outer-axis offsets and the validity flags are
// registers, and the inner loop over the pixels of a row is one remap + one byte load per channel.
#include "ts_kernels.h"

namespace ts {

namespace {

constexpr int NHWC_THREADS = 256;

struct NhwcPlan {
    int vec;            // elements per thread item (vec * esize == 4 when possible)
    int cg;             // channel groups = C / vec
    int tc, tp;         // threads along the channel groups / along the pixels of a row
    unsigned rows;      // N * prod(OS[0..dim-2])
    int rows_per_cta;
    unsigned groups;    // ceil(rows / rows_per_cta)
    int segs, seg_len;  // split of the last axis across blockIdx.y
    unsigned grid_x;    // CTAs along x (row groups are grid-strided when there are more)
};

template <typename E, int VEC> struct PackOut;
template <> struct PackOut<uint8_t, 4> {
    static TS_HD void store(uint8_t* p, const uint8_t* v) {
        *reinterpret_cast<uint32_t*>(p) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    }
};
template <typename E> struct PackOut<E, 1> {
    static TS_HD void store(E* p, const E* v) { *p = v[0]; }
};

// A size-1 axis needs no special case here: its reduced shift is 0 (reduce_shift), so the index is 0 and
// every padding mode maps 0 to 0.
template <int PAD> TS_HD int axis_index_c(int idx, int len) { return remap_bounded(idx, len, PAD); }

template <typename E> TS_HD E load_ro(const E* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// The whole per-thread program, host+device so the CPU-side tests can run the very same index logic
// over every (block, thread) of a small launch (ts_debug_nhwc_emulate) without a GPU.
template <typename E, int DIM, int VEC, int PAD>
TS_HD void nhwc_thread(const Geo& g, const NhwcPlan& pl, int tid, unsigned bx, unsigned by, unsigned grid_x,
                       const E* __restrict__ x, E* __restrict__ y, E fill, const void* __restrict__ w, int qkind, long long wzp) {
    constexpr int LAST = DIM - 1;
    const int cgi = tid % pl.tc;
    const int pw = tid / pl.tc;
    const int L = g.OS[LAST];
    const int p_begin = (int)by * pl.seg_len;
    const int p_end = L < p_begin + pl.seg_len ? L : p_begin + pl.seg_len;
    const unsigned xs_last_bytes = (unsigned)(g.xs[2 + LAST] * (long long)sizeof(E));   // < 2^31, checked by the launcher
    const int lb_last = g.lb[LAST], s_last = g.S[LAST];

    for (int cg = cgi; cg < pl.cg; cg += pl.tc) {
        const int c0 = cg * VEC;
        int sx[VEC][DIM];
#pragma unroll
        for (int v = 0; v < VEC; ++v) load_qshifts<DIM>(w, qkind, wzp, (long long)(c0 + v), g, sx[v]);

        for (unsigned grp = bx; grp < pl.groups; grp += grid_x) {
            const unsigned r_begin = grp * (unsigned)pl.rows_per_cta;
            const unsigned r_stop = r_begin + (unsigned)pl.rows_per_cta;
            const unsigned r_end = pl.rows < r_stop ? pl.rows : r_stop;
            for (unsigned row = r_begin; row < r_end; ++row) {
                // row -> (n, o0[, o1]) in output coordinates
                unsigned n = row;
                int o[2] = {0, 0};
                if (DIM == 2) {
                    n = row / (unsigned)g.OS[0];
                    o[0] = (int)(row - n * (unsigned)g.OS[0]);
                } else if (DIM == 3) {
                    const unsigned t = row / (unsigned)g.OS[1];
                    o[1] = (int)(row - t * (unsigned)g.OS[1]);
                    n = t / (unsigned)g.OS[0];
                    o[0] = (int)(t - n * (unsigned)g.OS[0]);
                }
                const char* xp[VEC];     // start of the source row of each channel (byte pointer)
                bool ok[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    long long base = (long long)n * g.xs[0] + (long long)(c0 + v) * g.xs[1];
                    ok[v] = true;
#pragma unroll
                    for (int a = 0; a < LAST; ++a) {
                        const int t = axis_index_c<PAD>(o[a] + g.lb[a] - sx[v][a], g.S[a]);
                        ok[v] = ok[v] && (t >= 0);
                        base += (long long)(t < 0 ? 0 : t) * g.xs[2 + a];
                    }
                    xp[v] = reinterpret_cast<const char*>(x + base);
                }
                E* yp = y + ((long long)row * L + (p_begin + pw)) * g.C + c0;
                const long long y_step = (long long)pl.tp * g.C;
#pragma unroll 2
                for (int p = p_begin + pw; p < p_end; p += pl.tp, yp += y_step) {
                    E val[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const int t = axis_index_c<PAD>(p + lb_last - sx[v][LAST], s_last);
                        const bool valid = ok[v] && (t >= 0);
                        // the load is unconditional (an invalid tap reads the row start, always inside x): no predicated
                        // address arithmetic in the loop, one select after the load
                        const unsigned long long off = (unsigned long long)(valid ? (unsigned)t : 0u) * xs_last_bytes;
                        const E got = load_ro(reinterpret_cast<const E*>(xp[v] + off));
                        val[v] = valid ? got : fill;
                    }
                    PackOut<E, VEC>::store(yp, val);
                }
            }
        }
    }
}

template <typename E, int DIM, int VEC, int PAD>
__global__ void __launch_bounds__(NHWC_THREADS) k_gather_nhwc(Geo g, NhwcPlan pl, const E* __restrict__ x, E* __restrict__ y,
                                                              E fill, const void* __restrict__ w, int qkind, long long wzp) {
    nhwc_thread<E, DIM, VEC, PAD>(g, pl, (int)threadIdx.x, blockIdx.x, blockIdx.y, gridDim.x, x, y, fill, w, qkind, wzp);
}

NhwcPlan make_plan(const Geo& g, int esize, const void* y, int sm_count, int max_grid_x) {
    NhwcPlan pl;
    pl.vec = (esize == 1 && g.C % 4 == 0 && ((uintptr_t)y & 3u) == 0) ? 4 : 1;
    pl.cg = (int)(g.C / pl.vec);
    if (pl.cg >= NHWC_THREADS) { pl.tc = NHWC_THREADS; pl.tp = 1; }
    else { pl.tc = pl.cg; pl.tp = NHWC_THREADS / pl.cg; }
    long long rows = g.N;
    for (int a = 0; a < g.dim - 1; ++a) rows *= g.OS[a];
    pl.rows = (unsigned)rows;
    const int L = g.OS[g.dim - 1];
    const long long want_ctas = (long long)sm_count * 16;
    long long rb = rows / want_ctas;
    if (rb < 1) rb = 1;
    if (rb > 8) rb = 8;
    const long long items_per_row = (long long)L * pl.cg;
    while (rb * items_per_row < 16 * NHWC_THREADS && rb < 64 && rb < rows) rb *= 2;
    pl.rows_per_cta = (int)rb;
    pl.groups = (unsigned)((rows + rb - 1) / rb);
    long long segs = 1;
    if ((long long)pl.groups < (long long)sm_count * 8) {
        segs = ((long long)sm_count * 8 + pl.groups - 1) / pl.groups;
        const long long max_segs = (L + pl.tp - 1) / pl.tp;
        if (segs > max_segs) segs = max_segs;
        if (segs > 65535) segs = 65535;
        if (segs < 1) segs = 1;
    }
    pl.seg_len = (int)((L + segs - 1) / segs);
    pl.segs = (L + pl.seg_len - 1) / pl.seg_len;
    const unsigned cap = max_grid_x > 0 ? (unsigned)max_grid_x : (1u << 20);
    pl.grid_x = pl.groups > cap ? cap : pl.groups;
    return pl;
}

template <typename E, int DIM, int VEC, int PAD>
void run_dim(const Geo& g, const NhwcPlan& pl, const void* x, void* y, E fill, const void* w, int qkind, long long wzp, cudaStream_t s,
             bool emulate) {
    const int threads = pl.tc * pl.tp;
    if (emulate) {      // host mirror: the same per-thread program, every (block, thread) in turn
        for (unsigned by = 0; by < (unsigned)pl.segs; ++by)
            for (unsigned bx = 0; bx < pl.grid_x; ++bx)
                for (int tid = 0; tid < threads; ++tid)
                    nhwc_thread<E, DIM, VEC, PAD>(g, pl, tid, bx, by, pl.grid_x, (const E*)x, (E*)y, fill, w, qkind, wzp);
        return;
    }
    const dim3 grid(pl.grid_x, (unsigned)pl.segs, 1);
    k_gather_nhwc<E, DIM, VEC, PAD><<<grid, threads, 0, s>>>(g, pl, (const E*)x, (E*)y, fill, w, qkind, wzp);
}

template <typename E, int VEC, int PAD>
void run_pad(const Geo& g, const NhwcPlan& pl, const void* x, void* y, E fill, const void* w, int qkind, long long wzp, cudaStream_t s,
             bool emulate) {
    switch (g.dim) {
    case 1: run_dim<E, 1, VEC, PAD>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    case 2: run_dim<E, 2, VEC, PAD>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    default: run_dim<E, 3, VEC, PAD>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    }
}

template <typename E, int VEC>
int run_type(const Geo& g, const NhwcPlan& pl, const void* x, void* y, E fill, const void* w, int qkind, long long wzp, cudaStream_t s,
             bool emulate) {
    switch (g.pad) {
    case TS_PAD_BORDER: run_pad<E, VEC, TS_PAD_BORDER>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    case TS_PAD_PERIODIC: run_pad<E, VEC, TS_PAD_PERIODIC>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    case TS_PAD_REFLECT: run_pad<E, VEC, TS_PAD_REFLECT>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    case TS_PAD_SYMMETRIC: run_pad<E, VEC, TS_PAD_SYMMETRIC>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    default: run_pad<E, VEC, TS_PAD_ZEROS>(g, pl, x, y, fill, w, qkind, wzp, s, emulate); break;
    }
    if (emulate) return TS_OK;
    note_launch();
    return check_launch();
}

}  // namespace

// x: any strides (g.xs), meant for channel stride 1; y: dense [N, OS0(,OS1(,OS2)), C].
// emulate: x / y / w are HOST pointers and the launch is walked on the host (tests only, no GPU work).
int nhwc_gather(const Geo& g, const void* x, void* y, unsigned long long fill, int esize, const void* w, int qkind,
                long long wzp, int sm_count, int max_grid_x, bool emulate, cudaStream_t s) {
    if (g.N * g.C == 0 || g.out_plane == 0) return TS_OK;
    long long rows = g.N;
    for (int a = 0; a < g.dim - 1; ++a) rows *= g.OS[a];
    if (rows >= 0x7fffffffLL || g.C >= 0x7fffffffLL) return TS_ERR_TOO_LARGE;
    const long long xsl = g.xs[2 + g.dim - 1];
    if (xsl < 0 || xsl * (long long)esize * g.S[g.dim - 1] >= 0x7fffffffLL) return TS_ERR_UNSUPPORTED;   // 32-bit row offsets
    const NhwcPlan pl = make_plan(g, esize, y, sm_count, max_grid_x);
    if (esize == 1) {
        if (pl.vec == 4) return run_type<uint8_t, 4>(g, pl, x, y, (uint8_t)fill, w, qkind, wzp, s, emulate);
        return run_type<uint8_t, 1>(g, pl, x, y, (uint8_t)fill, w, qkind, wzp, s, emulate);
    }
    if (esize == 4) return run_type<uint32_t, 1>(g, pl, x, y, (uint32_t)fill, w, qkind, wzp, s, emulate);
    return TS_ERR_UNSUPPORTED;
}

}  // namespace ts
