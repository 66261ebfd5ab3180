// ts_tma.cu -- the zero-padding fast path, built on TMA *tensor* copies (cp.async.bulk.tensor.5d,
// SASS UTMALDG).
//
// Idea.  A tiled TMA load whose box starts at the (possibly negative) coordinates
//     (col0 - s_col rounded down to 16 bytes, row0 - s_row, slab0 - s_slab, c, n0)
// delivers a tile of the plane(s) already SHIFTED along every axis, with every element that lies
// outside the tensor filled with zero by the copy engine.  That is the reference's zero-padded
// gather (ops/kernels/shifts_kernels.h:10-54 with BIPadding::Zeros) done by hardware, so the
// consumers have no index remapping, no validity masks and no edge cases at all: every item reads
// aligned 16-byte groups from shared memory and only resolves the residual sub-16-byte column
// misalignment (0..3 fp32 elements; the copy engine requires the innermost coordinate to be a
// multiple of 16 bytes -- measured: an unaligned inner coordinate raises "illegal instruction").
// The box carries one extra 16-byte group per row (and one extra row / slab for the arithmetic
// kernels) for the residual window and the +1 interpolation neighbours.
//
//   mode 0  sparse / quantized forward (pad value 0): y item = funnel-shifted pair of groups
//   mode 1  active forward:  x box (TA+1, TB+1, TG+1 groups) -> exact unfused lerp nest
//   mode 2  backward (no border crop): x box as mode 1, grad box unshifted (gv), and grad box
//           shifted by +s (sparse: grad_input is a gather) or by -s with +1 neighbours (active)
//
// CTA = `nw` consumer warps + 1 producer warp (one elected lane issues the TMA loads), persistent,
// one CTA per SM, ring of `stages` stages with full/empty mbarriers; grad_weight partials are
// per-(unit, warp) in fp64, fixed shuffle tree, second pass in ts_generic.cu -> deterministic.
//
// Applicability (plan_tma): zeros padding, dense NCHW x, row bytes a multiple of 16, pad value 0,
// fp32 for the arithmetic modes, no border crop for the backward.  Everything else runs on
// ts_staged.cu / ts_generic.cu.
#include <cuda.h>

#include <cstdio>

#include "ts_kernels.h"

namespace ts {

namespace {

constexpr int SMEM_LIMIT = 232448;
constexpr int MAXT_TMA_GATHER = 1024, MAXT_TMA_ARITH = 512;
constexpr int TABLE_MAX_C = 512;      // channels whose shift parameters are tabulated in shared memory (24 bytes each)

// ---- exact division by a launch-invariant (n < 2^31) ------------------------------------------
struct FastDiv { unsigned m, l, d; };
FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d ? d : 1;
    unsigned l = 0;
    while ((1ull << l) < f.d) ++l;
    f.l = l;
    f.m = (unsigned)(((((unsigned long long)1 << l) - f.d) << 32) / f.d + 1);
    return f;
}
TS_D unsigned fdiv(unsigned n, const FastDiv& f) { return (__umulhi(n, f.m) + n) >> f.l; }

// ---- driver entry point (no libcuda link dependency) ------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        (void)cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// rank-5 map over a dense [N][C][A][B][L] tensor of `es`-byte elements, box {bl, bb, ba, 1, bn}.
// Encoded maps are memoised per thread (a training loop presents the same few (pointer, geometry) pairs every
// step; the descriptor is a pure function of these arguments).
struct MapKey {
    const void* base;
    long long N, C;
    int es, A, B, L, bl, bb, ba, bn;
};
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
constexpr int MAP_CACHE = 32;

bool make_map(CUtensorMap* map, const void* base, int es, long long N, long long C, int A, int B, int L, int bl, int bb, int ba,
              int bn) {
    static thread_local MapSlot cache[MAP_CACHE];
    static thread_local int next_victim = 0;
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.base = base; key.N = N; key.C = C; key.es = es; key.A = A; key.B = B; key.L = L; key.bl = bl; key.bb = bb; key.ba = ba; key.bn = bn;
    for (int i = 0; i < MAP_CACHE; ++i)
        if (cache[i].used && !memcmp(&cache[i].key, &key, sizeof(key))) { *map = cache[i].map; return true; }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context current on the calling
    // thread.  A thread that has only used cached allocations so far (PyTorch's autograd worker on
    // its first backward) has none yet (CUDA_ERROR_INVALID_CONTEXT); cudaFree(0) binds it.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { (void)cudaFree(nullptr); ctx_bound = true; }
    const CUtensorMapDataType dt = es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                 : es == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    cuuint64_t dims[5] = {(cuuint64_t)L, (cuuint64_t)B, (cuuint64_t)A, (cuuint64_t)C, (cuuint64_t)N};
    const cuuint64_t row = (cuuint64_t)L * es;
    cuuint64_t strides[4] = {row, row * B, row * B * A, row * B * A * (cuuint64_t)C};
    cuuint32_t box[5] = {(cuuint32_t)bl, (cuuint32_t)bb, (cuuint32_t)ba, 1u, (cuuint32_t)bn};
    cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    const CUresult r = enc(map, dt, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[240];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled -> %d: es %d dims {%d,%d,%d,%lld,%lld} box {%d,%d,%d,1,%d} base %p", (int)r, es, L, B,
                 A, C, N, bl, bb, ba, bn, base);
        note_error(msg);
        return false;
    }
    MapSlot& slot = cache[next_victim];
    next_victim = (next_victim + 1) % MAP_CACHE;
    slot.key = key; slot.map = *map; slot.used = true;
    return true;
}

// ---- PTX wrappers -----------------------------------------------------------------------------
TS_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
TS_D void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
TS_D void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TS_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TS_D void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
TS_D void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// tiled 5-D TMA load global -> shared; coordinates innermost first; OOB elements arrive as zero.
// c0 must be a multiple of 16 bytes worth of elements.
TS_D void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}

// ---- arguments -------------------------------------------------------------------------------
struct alignas(64) TArgs {
    CUtensorMap map_x;       // x,    box {(TG+1)*vec, xb, xa, 1, np}
    CUtensorMap map_gs;      // grad, box {(TG+1)*vec, TB, TA, 1, np}     (backward)
    CUtensorMap map_gb;      // grad, box {(TG+1)*vec, xb, xa, 1, np}     (active backward)
    Geo g;
    unsigned char* out;
    const void* w;
    double* partials;
    long long wzp;
    int qkind, wk, es, vec;  // vec = elements per 16-byte group
    int mode, active;
    int OA, OB, OGR;         // output slabs, rows, 16-byte groups per row
    int lbA, lbB, lbL;
    int TA, TB, TG;          // tile extents (slabs, rows, groups)
    int xa, xb;              // x box slabs / rows (TA or TA+1, TB or TB+1)
    int tiles_b, tiles_g, tiles;
    int np;                  // images per stage (1 unless a whole plane is one tile)
    int x_img_chunks;        // 16-byte groups of one image's x box
    int g_img_chunks;        // 16-byte groups of one image's small grad box
    int off_gv, off_g2;      // byte offsets of the grad boxes inside a stage
    int tx_bytes;            // bytes landing per stage
    int stage_stride, stages, nw, n_per_unit, units, chunks, unit_order;
    int GP, img_items;       // padded groups per row of the item index space; items per image = TA*TB*GP
    int img_stride16;        // output distance between consecutive images of one channel, in 16-byte units
    int R, nchunk;           // strip-mined arithmetic kernels: rows per strip, strips per tile column
    FastDiv d_img, d_GP, d_TB, d_tg, d_tb, d_TG, d_nchunk, d_TA, d_C, d_chunks;
    int table;               // per-channel shift table in shared memory (C <= TABLE_MAX_C): built once per CTA
};

TS_D int level_axis(int level, int dim) { return level - (3 - dim); }
TS_D int floor_to(int v, int q) { return v & ~(q - 1); }      // q is a power of two (elements per 16 bytes)

struct UnitShift {
    int sh[3];     // integer shift per level (0 slab, 1 row, 2 column); absent levels 0
    float d[3];    // fractional part per TENSOR AXIS (reference order), 0 for the sparse forward
};

TS_D UnitShift compute_unit_shift(const TArgs& a, long long c) {
    UnitShift u;
    const int dim = a.g.dim;
    u.d[0] = u.d[1] = u.d[2] = 0.f;
#pragma unroll
    for (int lev = 0; lev < 3; ++lev) {
        const int ax = level_axis(lev, dim);
        long long iw = 0;
        if (ax >= 0) {
            const long long idx = c * dim + ax;
            if (a.mode == 0) {
                switch (a.wk) {
                case WK_F32: { float d; split_forward<float>(((const float*)a.w)[idx], false, iw, d); break; }
                case WK_F64: { double d; split_forward<double>(((const double*)a.w)[idx], false, iw, d); break; }
                case WK_F16: { float d; split_forward<float>(__half2float(((const __half*)a.w)[idx]), false, iw, d); break; }
                case WK_BF16: { float d; split_forward<float>(__bfloat162float(((const __nv_bfloat16*)a.w)[idx]), false, iw, d); break; }
                default:
                    if (a.qkind == TS_QW_U8) iw = (long long)((const uint8_t*)a.w)[idx] - a.wzp;
                    else if (a.qkind == TS_QW_I8) iw = (long long)((const int8_t*)a.w)[idx] - a.wzp;
                    else iw = (long long)((const int32_t*)a.w)[idx] - a.wzp;
                }
            } else {
                float d;
                if (a.mode == 1) split_forward<float>(((const float*)a.w)[idx], true, iw, d);
                else split_backward<float>(((const float*)a.w)[idx], a.active != 0, iw, d);
                u.d[ax] = d;
            }
            u.sh[lev] = reduce_shift(iw, a.g.S[ax], TS_PAD_ZEROS);
        } else {
            u.sh[lev] = 0;
        }
    }
    return u;
}

// The per-channel parameters cost a few hundred dependent instructions (64-bit conversions, saturation, the reduced
// shift): computed per work unit by the single producer thread they bounded small batches (one stage per unit).
// Every CTA tabulates them once in shared memory instead; a unit then costs one 24-byte read.
TS_D UnitShift unit_shift(const TArgs& a, const UnitShift* tbl, long long c) { return a.table ? tbl[c] : compute_unit_shift(a, c); }

// u -> (channel, chunk) without an integer division
TS_D void decode_unit(const TArgs& a, int u, int& c, int& chunk) {
    if (a.unit_order) { chunk = (int)fdiv((unsigned)u, a.d_C); c = u - chunk * (int)a.g.C; }
    else { c = (int)fdiv((unsigned)u, a.d_chunks); chunk = u - c * a.chunks; }
}

struct Tile { int a0, b0, g0; };
TS_D Tile tile_of(const TArgs& a, int t) {
    Tile tl;
    const int tg = (int)fdiv((unsigned)t, a.d_tg);          // t / tiles_g
    tl.g0 = (t - tg * a.tiles_g) * a.TG;
    const int ta = (int)fdiv((unsigned)tg, a.d_tb);         // (t / tiles_g) / tiles_b
    tl.b0 = (tg - ta * a.tiles_b) * a.TB;
    tl.a0 = ta * a.TA;
    return tl;
}

// ---- producer ---------------------------------------------------------------------------------
TS_D void producer(const TArgs& a, unsigned char* smem, uint64_t* full, uint64_t* empty, const UnitShift* tbl) {
    int s = 0, k = 0;
    const long long C = a.g.C, N = a.g.N;
    const int dim = a.g.dim;
    const UnitRange ur = unit_range(a.units, a.unit_order);
    for (int u = ur.u; u < ur.end; u += ur.step) {
        int ci, chunki;
        decode_unit(a, u, ci, chunki);
        const long long c = ci, chunk = chunki;
        const long long n0 = chunk * a.n_per_unit;
        const long long n1 = n0 + a.n_per_unit < N ? n0 + a.n_per_unit : N;
        const UnitShift us = unit_shift(a, tbl, c);
        for (long long nb = n0; nb < n1; nb += a.np) {
            for (int t = 0; t < a.tiles; ++t) {
                if (k > 0) mbar_wait(&empty[s], (unsigned)((k - 1) & 1));
                const Tile tl = tile_of(a, t);
                unsigned char* st = smem + (size_t)s * a.stage_stride;
                mbar_expect_tx(&full[s], (unsigned)a.tx_bytes);
                const int l0 = tl.g0 * a.vec;
                // x box: source coordinates of the tile's first output element, column rounded down to a group
                const int xc = floor_to(l0 + a.lbL - us.sh[2], a.vec);
                const int xr = dim >= 2 ? tl.b0 + a.lbB - us.sh[1] : 0;
                const int xs = dim == 3 ? tl.a0 + a.lbA - us.sh[0] : 0;
                tma_load_5d(st, &a.map_x, xc, xr, xs, (int)c, (int)nb, &full[s]);
                if (a.mode == 2) {
                    tma_load_5d(st + a.off_gv, &a.map_gs, l0, dim >= 2 ? tl.b0 : 0, dim == 3 ? tl.a0 : 0, (int)c, (int)nb, &full[s]);
                    if (a.active) {      // grad at (o - s) with +1 neighbours: same geometry as the x box
                        tma_load_5d(st + a.off_g2, &a.map_gb, xc, xr, xs, (int)c, (int)nb, &full[s]);
                    } else {             // grad_input of the sparse shift gathers grad at (o + s)
                        tma_load_5d(st + a.off_g2, &a.map_gs, floor_to(l0 + us.sh[2], a.vec), dim >= 2 ? tl.b0 + us.sh[1] : 0,
                                    dim == 3 ? tl.a0 + us.sh[0] : 0, (int)c, (int)nb, &full[s]);
                    }
                }
                if (++s == a.stages) { s = 0; ++k; }
            }
        }
    }
}

// ---- consumer skeleton -----------------------------------------------------------------------
// Everything a body needs about the current stage; all values are CTA-uniform.
struct Stage {
    const unsigned char* st;   // stage base in shared memory
    unsigned char* dst;        // output address of (first image of the stage, channel c, tile origin)
    int npl;                   // images in this stage
    int total;                 // items of this stage in the padded index space: images * img_items
    int an, bn, gn;            // valid extents of this tile (slabs, rows, groups)
};

template <class Body>
TS_D void consumer_loop(const TArgs& a, unsigned char* smem, uint64_t* full, uint64_t* empty, int lane, Body& body) {
    int s = 0;
    unsigned phase = 0;
    const int C = (int)a.g.C, N = (int)a.g.N, np = a.np, tiles = a.tiles, stages = a.stages;
    const long long plane_bytes = (a.mode == 2 ? a.g.in_plane : a.g.out_plane) * a.es;
    const UnitRange ur = unit_range(a.units, a.unit_order);
    for (int u = ur.u; u < ur.end; u += ur.step) {
        int chunk, c;
        decode_unit(a, u, c, chunk);
        const int n0 = chunk * a.n_per_unit;
        const int n1 = n0 + a.n_per_unit < N ? n0 + a.n_per_unit : N;
        body.begin_unit(c);
        for (int nb = n0; nb < n1; nb += np) {
            Stage sg;
            sg.npl = n1 - nb < np ? n1 - nb : np;
            sg.total = sg.npl * a.img_items;
            unsigned char* img = a.out + ((long long)nb * C + c) * plane_bytes;
            for (int t = 0; t < tiles; ++t) {
                sg.an = a.TA; sg.bn = a.TB; sg.gn = a.TG;
                sg.dst = img;
                if (tiles > 1) {
                    const Tile tl = tile_of(a, t);
                    sg.an = a.OA - tl.a0 < a.TA ? a.OA - tl.a0 : a.TA;
                    sg.bn = a.OB - tl.b0 < a.TB ? a.OB - tl.b0 : a.TB;
                    sg.gn = a.OGR - tl.g0 < a.TG ? a.OGR - tl.g0 : a.TG;
                    sg.dst = img + ((long long)(tl.a0 * a.OB + tl.b0) * a.OGR + tl.g0) * 16;
                }
                sg.st = smem + (size_t)s * a.stage_stride;
                mbar_wait(&full[s], phase);
                body.step(sg);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == stages) { s = 0; phase ^= 1u; }
            }
        }
        body.end_unit(c, chunk);
    }
}

// item (padded index space: [image][slab][row][GP groups]) -> position inside the tile.  The
// padded row length GP is a multiple of 8 when that wastes <= 15 % of the lanes: a quarter-warp
// then never straddles two box rows, which keeps the 128-bit shared-memory loads conflict-free.
struct Item { int pl, a, b, cg; };
template <bool SLABS>
TS_D bool decode_item(const TArgs& a, const Stage& sg, int item, Item& p) {
    p.pl = 0;
    int rem = item;
    if (a.np > 1) { p.pl = (int)fdiv((unsigned)item, a.d_img); rem = item - p.pl * a.img_items; }
    const int row = (int)fdiv((unsigned)rem, a.d_GP);
    p.cg = rem - row * a.GP;
    p.a = 0;
    p.b = row;
    if (SLABS) { p.a = (int)fdiv((unsigned)row, a.d_TB); p.b = row - p.a * a.TB; }
    return p.cg < sg.gn && p.b < sg.bn && p.a < sg.an;
}
// 16-byte offset of the item's output inside the stage's destination
TS_D unsigned char* item_dst(const TArgs& a, const Stage& sg, const Item& p) {
    const int off16 = p.pl * a.img_stride16 + (p.a * a.OB + p.b) * a.OGR + p.cg;
    return sg.dst + (size_t)(unsigned)off16 * 16;
}

// ================================================================================================
// mode 0: sparse / quantized forward.  WS = word misalignment of the source window (0..3); SUB =
// elements narrower than 4 bytes, whose residual byte shift is a run-time funnel-shift amount.
template <bool SLABS>
struct GatherBody {
    const TArgs& a;
    const int tid, nt;
    int ws, bs8;

    const UnitShift* tbl;
    TS_D GatherBody(const TArgs& a_, int tid_, int nt_, const UnitShift* tbl_) : a(a_), tid(tid_), nt(nt_), ws(0), bs8(0), tbl(tbl_) {}
    TS_D void begin_unit(int c) {
        const UnitShift us = unit_shift(a, tbl, c);
        const int mb = pmod((a.lbL - us.sh[2]) * a.es, 16);    // tiles start at multiples of 16 bytes
        ws = mb >> 2;
        bs8 = (mb & 3) * 8;
    }
    TS_D void end_unit(int, int) {}

    template <int WS, bool SUB>
    TS_D void run(const Stage& sg) const {
        const unsigned box = shared_addr(sg.st);
        const int pitch = a.TG + 1, ximg = a.x_img_chunks, xb = a.xb;
        for (int item = tid; item < sg.total; item += nt) {
            Item p;
            if (!decode_item<SLABS>(a, sg, item, p)) continue;
            const int ch = p.pl * ximg + (p.a * xb + p.b) * pitch + p.cg;
            const uint4 A = lds128(box + ch * 16);
            unsigned W[8] = {A.x, A.y, A.z, A.w, 0u, 0u, 0u, 0u};
            if (WS > 0 || SUB) { const uint4 B = lds128(box + ch * 16 + 16); W[4] = B.x; W[5] = B.y; W[6] = B.z; W[7] = B.w; }
            unsigned o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = SUB ? __funnelshift_r(W[k + WS], W[(k + WS + 1) & 7], bs8) : W[k + WS];
            __stcs((uint4*)item_dst(a, sg, p), make_uint4(o[0], o[1], o[2], o[3]));
        }
    }
    TS_D void step(const Stage& sg) const {
        if (bs8 == 0) {
            switch (ws) {
            case 0: run<0, false>(sg); break;
            case 1: run<1, false>(sg); break;
            case 2: run<2, false>(sg); break;
            default: run<3, false>(sg); break;
            }
        } else {
            switch (ws) {
            case 0: run<0, true>(sg); break;
            case 1: run<1, true>(sg); break;
            case 2: run<2, true>(sg); break;
            default: run<3, true>(sg); break;
            }
        }
    }
};

// ================================================================================================
// fp32 arithmetic bodies.  A window of NV consecutive elements starting M elements into the
// aligned group `ch` of a staged box.
template <int M, int NV>
TS_D void load_win(unsigned box, int ch, float* out) {
    const uint4 A = lds128(box + ch * 16);
    float W[8] = {__uint_as_float(A.x), __uint_as_float(A.y), __uint_as_float(A.z), __uint_as_float(A.w), 0.f, 0.f, 0.f, 0.f};
    if (M + NV > 4) {
        const uint4 B = lds128(box + ch * 16 + 16);
        W[4] = __uint_as_float(B.x); W[5] = __uint_as_float(B.y); W[6] = __uint_as_float(B.z); W[7] = __uint_as_float(B.w);
    }
#pragma unroll
    for (int t = 0; t < NV; ++t) out[t] = W[t + M];
}

// group offset of the +1 neighbour rows inside a box with `rows` rows per slab, `pitch` groups per row
template <int DIM>
TS_D int neighbour_row_offset(int rv, int rows, int pitch) {
    if (DIM == 1) return 0;
    if (DIM == 2) return (rv & 1) * pitch;
    return ((rv & 1) * rows + ((rv >> 1) & 1)) * pitch;
}

template <int DIM>
TS_D void neighbours_from_rows(const float (*X)[5], int t, float* v) {
    constexpr int NR = 1 << (DIM - 1);
#pragma unroll
    for (int q = 0; q < (1 << DIM); ++q) v[q] = X[q & (NR - 1)][t + (q >> (DIM - 1))];
}

// The reference's per-element weight-gradient factors (ts_common.cuh weight_partials) written with
// ordinary, contractable arithmetic: grad_weight is tolerance-checked (rtol 1e-5 vs an fp64 CPU
// evaluation of the reference formulas), only forward / grad_input have to be bit-exact.
template <int DIM>
TS_D void weight_partials_fast(const float* v, const float* d, float* g) {
    if (DIM == 1) { g[0] = v[1] - v[0]; return; }
    if (DIM == 2) {
        const float p = v[2] - v[0], q = (v[3] - v[1]) - p;   // both factors are p + frac * q
        g[0] = fmaf(d[1], q, p);
        g[1] = fmaf(d[0], q, p);
        return;
    }
    const float p0 = v[2] - v[0], q0 = (v[3] - v[1]) - p0, p1 = v[6] - v[4], q1 = (v[7] - v[5]) - p1;
    const float x0 = fmaf(d[1], q0, p0), x1 = fmaf(d[1], q1, p1);
    const float y0 = fmaf(d[0], q0, p0), y1 = fmaf(d[0], q1, p1);
    g[0] = fmaf(d[2], x1 - x0, x0);
    g[1] = fmaf(d[2], y1 - y0, y0);
    const float i0 = fmaf(d[0], v[1] - v[0], v[0]), i1 = fmaf(d[0], v[3] - v[2], v[2]);
    const float i2 = fmaf(d[0], v[5] - v[4], v[4]), i3 = fmaf(d[0], v[7] - v[6], v[6]);
    g[2] = fmaf(d[1], i3 - i2, i2) - fmaf(d[1], i1 - i0, i0);
}

template <int DIM>
struct ActiveFwdBody {
    const TArgs& a;
    const int tid, nt;
    UnitShift us;
    int m;

    const UnitShift* tbl;
    TS_D ActiveFwdBody(const TArgs& a_, int tid_, int nt_, const UnitShift* tbl_) : a(a_), tid(tid_), nt(nt_), m(0), tbl(tbl_) {}
    TS_D void begin_unit(int c) {
        us = unit_shift(a, tbl, c);
        m = pmod(a.lbL - us.sh[2], 4);
    }
    TS_D void end_unit(int, int) {}

    template <int M>
    TS_D void run(const Stage& sg) const {
        constexpr int NR = 1 << (DIM - 1);
        const unsigned box = shared_addr(sg.st);
        const int pitch = a.TG + 1, ximg = a.x_img_chunks, xb = a.xb;
        const float d[3] = {us.d[0], us.d[1], us.d[2]};
        for (int item = tid; item < sg.total; item += nt) {
            Item p;
            if (!decode_item<DIM == 3>(a, sg, item, p)) continue;
            const int base = p.pl * ximg + (p.a * xb + p.b) * pitch + p.cg;
            float X[NR][5];
#pragma unroll
            for (int rv = 0; rv < NR; ++rv) load_win<M, 5>(box, base + neighbour_row_offset<DIM>(rv, xb, pitch), X[rv]);
            float o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float v[8];
                neighbours_from_rows<DIM>(X, t, v);
                o[t] = interpolate<float, DIM>(v, d);
            }
            __stcs((float4*)item_dst(a, sg, p), make_float4(o[0], o[1], o[2], o[3]));
        }
    }
    // strip-mined variant, used for 3-D (see BackwardBody::run_strip)
    template <int M>
    TS_D void run_strip(const Stage& sg) const {
        constexpr int NR = 1 << (DIM - 1);
        constexpr int S = DIM == 3 ? 2 : 1;
        const unsigned box = shared_addr(sg.st);
        const int pitch = a.TG + 1, ximg = a.x_img_chunks, xb = a.xb;
        const int xslab = xb * pitch;
        const float d[3] = {us.d[0], us.d[1], us.d[2]};
        const int R = a.R, nchunk = a.nchunk, TG = a.TG, TA = a.TA;
        const int strips = sg.npl * TA * nchunk * TG;
        const int orow = a.OGR * 16;
        for (int sidx = tid; sidx < strips; sidx += nt) {
            const int r = (int)fdiv((unsigned)sidx, a.d_TG), cg = sidx - r * TG;
            const int r2 = (int)fdiv((unsigned)r, a.d_nchunk), kc = r - r2 * nchunk;
            int pl = r2, ia = 0;
            if (TA > 1) { pl = (int)fdiv((unsigned)r2, a.d_TA); ia = r2 - pl * TA; }
            const int b0 = kc * R;
            if (cg >= sg.gn || ia >= sg.an || b0 >= sg.bn) continue;
            const int bend = b0 + R < sg.bn ? b0 + R : sg.bn;
            Item p;
            p.pl = pl; p.a = ia; p.b = b0; p.cg = cg;
            int xch = pl * ximg + (ia * xb + b0) * pitch + cg;
            unsigned char* dst = item_dst(a, sg, p);
            float Xlo[S][5];
#pragma unroll
            for (int q = 0; q < S; ++q) load_win<M, 5>(box, xch + q * xslab, Xlo[q]);
            for (int b = b0; b < bend; ++b) {
                float X[NR][5];
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    load_win<M, 5>(box, xch + pitch + q * xslab, X[S + q]);
#pragma unroll
                    for (int t = 0; t < 5; ++t) X[q][t] = Xlo[q][t];
                }
                float o[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v[8];
                    neighbours_from_rows<DIM>(X, t, v);
                    o[t] = interpolate<float, DIM>(v, d);
                }
                __stcs((float4*)dst, make_float4(o[0], o[1], o[2], o[3]));
#pragma unroll
                for (int q = 0; q < S; ++q)
#pragma unroll
                    for (int t = 0; t < 5; ++t) Xlo[q][t] = X[S + q][t];
                xch += pitch; dst += orow;
            }
        }
    }
    template <int M>
    TS_D void run_any(const Stage& sg) const {
        if constexpr (DIM == 3) run_strip<M>(sg); else run<M>(sg);   // 2-D is HBM-bound either way and the flat,
                                                                     // padded index space is measurably faster there
    }
    TS_D void step(const Stage& sg) const {
        switch (m) {
        case 0: run_any<0>(sg); break;
        case 1: run_any<1>(sg); break;
        case 2: run_any<2>(sg); break;
        default: run_any<3>(sg); break;
        }
    }
};

template <int DIM, bool ACTIVE>
struct BackwardBody {
    const TArgs& a;
    const int tid, nt, wid, lane;
    UnitShift us;
    int m;
    double acc[DIM];

    const UnitShift* tbl;
    TS_D BackwardBody(const TArgs& a_, int tid_, int nt_, int wid_, int lane_, const UnitShift* tbl_)
        : a(a_), tid(tid_), nt(nt_), wid(wid_), lane(lane_), m(0), tbl(tbl_) {}
    TS_D void begin_unit(int c) {
        us = unit_shift(a, tbl, c);
        m = pmod(-us.sh[2], 4);
#pragma unroll
        for (int d = 0; d < DIM; ++d) acc[d] = 0.0;
    }
    // one partial per (unit, consumer warp): fixed shuffle tree, no atomics
    TS_D void end_unit(int c, int chunk) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double v = acc[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) a.partials[((long long)chunk * a.nw + wid) * (a.g.C * DIM) + (long long)c * DIM + d] = v;
        }
    }

    template <int M>
    TS_D void run(const Stage& sg) {
        constexpr int NR = 1 << (DIM - 1);
        constexpr int MG = ACTIVE ? M : ((4 - M) & 3);      // misalignment of the grad window used for grad_input
        const unsigned xbox = shared_addr(sg.st);
        const unsigned gv_box = xbox + a.off_gv;
        const unsigned g2_box = xbox + a.off_g2;
        const int pitch = a.TG + 1, ximg = a.x_img_chunks, gimg = a.g_img_chunks, xb = a.xb, tb = a.TB;
        const float d[3] = {us.d[0], us.d[1], us.d[2]};
        float ts[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) ts[k] = 0.f;
        for (int item = tid; item < sg.total; item += nt) {
            Item p;
            if (!decode_item<DIM == 3>(a, sg, item, p)) continue;
            const int gch = p.pl * gimg + (p.a * tb + p.b) * pitch + p.cg;
            const int xch = p.pl * ximg + (p.a * xb + p.b) * pitch + p.cg;
            float gv[4];
            load_win<0, 4>(gv_box, gch, gv);
            // ---- grad_weight terms ----
            float X[NR][5];
#pragma unroll
            for (int rv = 0; rv < NR; ++rv) load_win<M, 5>(xbox, xch + neighbour_row_offset<DIM>(rv, xb, pitch), X[rv]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float v[8], wg[3];
                neighbours_from_rows<DIM>(X, t, v);
                weight_partials_fast<DIM>(v, d, wg);
#pragma unroll
                for (int k = 0; k < DIM; ++k) ts[k] = fmaf(gv[t], wg[k], ts[k]);
            }
            // ---- grad_input ----
            float o[4];
            if (ACTIVE) {
                float G[NR][5];
#pragma unroll
                for (int rv = 0; rv < NR; ++rv) load_win<MG, 5>(g2_box, xch + neighbour_row_offset<DIM>(rv, xb, pitch), G[rv]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v[8];
                    neighbours_from_rows<DIM>(G, t, v);
                    o[t] = interpolate<float, DIM>(v, d);
                }
            } else {
                load_win<MG, 4>(g2_box, gch, o);
            }
            __stcs((float4*)item_dst(a, sg, p), make_float4(o[0], o[1], o[2], o[3]));
        }
#pragma unroll
        for (int k = 0; k < DIM; ++k) acc[k] += (double)ts[k];   // fp32 inside a stage, fp64 across stages
    }
    // Strip-mined variant (used for 3-D, where the kernels are instruction-bound): a thread owns (image, slab, group, chunk of R rows); the "+1 row"
    // windows of one item are the "+0 row" windows of the next and stay in registers.
    template <int M>
    TS_D void run_strip(const Stage& sg) {
        constexpr int NR = 1 << (DIM - 1);
        constexpr int S = DIM == 3 ? 2 : 1;
        constexpr int MG = ACTIVE ? M : ((4 - M) & 3);
        const unsigned xbox = shared_addr(sg.st);
        const unsigned gv_box = xbox + a.off_gv;
        const unsigned g2_box = xbox + a.off_g2;
        const int pitch = a.TG + 1, ximg = a.x_img_chunks, gimg = a.g_img_chunks, xb = a.xb, tb = a.TB;
        const int xslab = xb * pitch;
        const float d[3] = {us.d[0], us.d[1], us.d[2]};
        float ts[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) ts[k] = 0.f;
        const int R = a.R, nchunk = a.nchunk, TG = a.TG, TA = a.TA;
        const int strips = sg.npl * TA * nchunk * TG;
        const int orow = a.OGR * 16;
        for (int sidx = tid; sidx < strips; sidx += nt) {
            const int r = (int)fdiv((unsigned)sidx, a.d_TG), cg = sidx - r * TG;
            const int r2 = (int)fdiv((unsigned)r, a.d_nchunk), kc = r - r2 * nchunk;
            int pl = r2, ia = 0;
            if (TA > 1) { pl = (int)fdiv((unsigned)r2, a.d_TA); ia = r2 - pl * TA; }
            const int b0 = kc * R;
            if (cg >= sg.gn || ia >= sg.an || b0 >= sg.bn) continue;
            const int bend = b0 + R < sg.bn ? b0 + R : sg.bn;
            Item p;
            p.pl = pl; p.a = ia; p.b = b0; p.cg = cg;
            int gch = pl * gimg + (ia * tb + b0) * pitch + cg;
            int xch = pl * ximg + (ia * xb + b0) * pitch + cg;
            unsigned char* dst = item_dst(a, sg, p);
            float Xlo[S][5], Glo[S][5];
#pragma unroll
            for (int q = 0; q < S; ++q) load_win<M, 5>(xbox, xch + q * xslab, Xlo[q]);
            if (ACTIVE) {
#pragma unroll
                for (int q = 0; q < S; ++q) load_win<MG, 5>(g2_box, xch + q * xslab, Glo[q]);
            }
            for (int b = b0; b < bend; ++b) {
                float gv[4];
                load_win<0, 4>(gv_box, gch, gv);
                float X[NR][5];
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    load_win<M, 5>(xbox, xch + pitch + q * xslab, X[S + q]);
#pragma unroll
                    for (int t = 0; t < 5; ++t) X[q][t] = Xlo[q][t];
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v[8], wg[3];
                    neighbours_from_rows<DIM>(X, t, v);
                    weight_partials_fast<DIM>(v, d, wg);
#pragma unroll
                    for (int k = 0; k < DIM; ++k) ts[k] = fmaf(gv[t], wg[k], ts[k]);
                }
                float o[4];
                if (ACTIVE) {
                    float G[NR][5];
#pragma unroll
                    for (int q = 0; q < S; ++q) {
                        load_win<MG, 5>(g2_box, xch + pitch + q * xslab, G[S + q]);
#pragma unroll
                        for (int t = 0; t < 5; ++t) G[q][t] = Glo[q][t];
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        float v[8];
                        neighbours_from_rows<DIM>(G, t, v);
                        o[t] = interpolate<float, DIM>(v, d);
                    }
#pragma unroll
                    for (int q = 0; q < S; ++q)
#pragma unroll
                        for (int t = 0; t < 5; ++t) Glo[q][t] = G[S + q][t];
                } else {
                    load_win<MG, 4>(g2_box, gch, o);
                }
                __stcs((float4*)dst, make_float4(o[0], o[1], o[2], o[3]));
#pragma unroll
                for (int q = 0; q < S; ++q)
#pragma unroll
                    for (int t = 0; t < 5; ++t) Xlo[q][t] = X[S + q][t];
                xch += pitch; gch += pitch; dst += orow;
            }
        }
#pragma unroll
        for (int k = 0; k < DIM; ++k) acc[k] += (double)ts[k];
    }
    template <int M>
    TS_D void run_any(const Stage& sg) {
        if constexpr (DIM == 3) run_strip<M>(sg); else run<M>(sg);   // 2-D is HBM-bound either way and the flat,
                                                                     // padded index space is measurably faster there
    }
    TS_D void step(const Stage& sg) {
        switch (m) {
        case 0: run_any<0>(sg); break;
        case 1: run_any<1>(sg); break;
        case 2: run_any<2>(sg); break;
        default: run_any<3>(sg); break;
        }
    }
};

// ---- kernels -----------------------------------------------------------------------------------
TS_D const UnitShift* setup_barriers(const TArgs& a, unsigned char* smem, uint64_t*& full, uint64_t*& empty) {
    full = (uint64_t*)(smem + (size_t)a.stages * a.stage_stride);
    empty = full + a.stages;
    UnitShift* tbl = (UnitShift*)(empty + a.stages);
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (unsigned)a.nw); }
        fence_barrier_init();
    }
    pdl_trigger();          // the next kernel of the stream may be scheduled as soon as SMs free up
    pdl_wait();             // nothing above touches global memory; everything below may depend on the previous kernel
    if (a.table)
        for (int c = threadIdx.x; c < (int)a.g.C; c += blockDim.x) tbl[c] = compute_unit_shift(a, c);
    __syncthreads();
    return tbl;
}

template <bool SLABS>
__global__ void __launch_bounds__(MAXT_TMA_GATHER, 1) k_tma_gather(const __grid_constant__ TArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full, *empty;
    const UnitShift* tbl = setup_barriers(a, smem, full, empty);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == a.nw) { if (lane == 0) producer(a, smem, full, empty, tbl); return; }
    GatherBody<SLABS> body(a, threadIdx.x, a.nw * 32, tbl);
    consumer_loop(a, smem, full, empty, lane, body);
}

template <int DIM>
__global__ void __launch_bounds__(MAXT_TMA_ARITH, 1) k_tma_active_forward(const __grid_constant__ TArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full, *empty;
    const UnitShift* tbl = setup_barriers(a, smem, full, empty);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == a.nw) { if (lane == 0) producer(a, smem, full, empty, tbl); return; }
    ActiveFwdBody<DIM> body(a, threadIdx.x, a.nw * 32, tbl);
    consumer_loop(a, smem, full, empty, lane, body);
}

template <int DIM, bool ACTIVE>
__global__ void __launch_bounds__(MAXT_TMA_ARITH, 1) k_tma_backward(const __grid_constant__ TArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full, *empty;
    const UnitShift* tbl = setup_barriers(a, smem, full, empty);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == a.nw) { if (lane == 0) producer(a, smem, full, empty, tbl); return; }
    BackwardBody<DIM, ACTIVE> body(a, threadIdx.x, a.nw * 32, wid, lane, tbl);
    consumer_loop(a, smem, full, empty, lane, body);
}

template <class K>
int launch(K kernel, const TArgs& a, const TmaPlan& p, cudaStream_t s) {
    if (!ensure_dynamic_smem((const void*)kernel, p.smem_bytes)) return check_launch();
    if (!tuning().no_pdl) {
        if (launch_pdl(kernel, dim3(p.grid), dim3((p.warps + 1) * 32), p.smem_bytes, s, a) != cudaSuccess) {
            (void)cudaGetLastError();
            kernel<<<p.grid, (p.warps + 1) * 32, p.smem_bytes, s>>>(a);
        }
    } else {
        kernel<<<p.grid, (p.warps + 1) * 32, p.smem_bytes, s>>>(a);
    }
    note_launch();
    return check_launch();
}

long long round_up(long long v, long long q) { return (v + q - 1) / q * q; }

// fills everything of TArgs that does not depend on pointers
bool make_args(const Geo& g, const TmaPlan& p, int mode, int active, int es, TArgs* out) {
    TArgs& a = *out;
    memset(&a, 0, sizeof(a));
    const int d = g.dim;
    a.g = g;
    a.es = es;
    a.vec = 16 / es;
    a.mode = mode;
    a.active = active;
    a.OA = d == 3 ? g.OS[0] : 1;      a.lbA = d == 3 ? g.lb[0] : 0;
    a.OB = d >= 2 ? g.OS[d - 2] : 1;  a.lbB = d >= 2 ? g.lb[d - 2] : 0;
    a.OGR = g.OS[d - 1] / a.vec;      a.lbL = g.lb[d - 1];
    a.TA = p.ta; a.TB = p.tb; a.TG = p.tg;
    a.xa = p.xa; a.xb = p.xb;
    a.tiles_b = (a.OB + a.TB - 1) / a.TB;
    a.tiles_g = (a.OGR + a.TG - 1) / a.TG;
    a.tiles = p.tiles_per_plane;
    a.np = p.np;
    a.x_img_chunks = a.xa * a.xb * (a.TG + 1);
    a.g_img_chunks = a.TA * a.TB * (a.TG + 1);
    a.off_gv = p.off_gv;
    a.off_g2 = p.off_g2;
    a.tx_bytes = p.tx_bytes;
    a.stage_stride = p.stage_stride;
    a.stages = p.stages;
    a.nw = p.warps;
    a.n_per_unit = p.n_per_unit;
    a.units = p.units;
    a.chunks = (int)(p.units / (g.C > 0 ? g.C : 1));
    a.unit_order = tuning().unit_order;
    a.d_C = make_fastdiv((unsigned)g.C);
    a.d_chunks = make_fastdiv((unsigned)a.chunks);
    a.table = (g.C <= TABLE_MAX_C && !tuning().no_table) ? 1 : 0;
    a.GP = p.gp;
    a.img_items = a.TA * a.TB * a.GP;
    a.img_stride16 = (int)(g.C * (mode == 2 ? g.in_plane : g.out_plane) * es / 16);
    {   // strips: about three per thread and stage, at least two rows each (the row reuse is the point)
        const long long items = (long long)a.np * a.TA * a.TB * a.TG, nt = 32ll * p.warps;
        long long R = (items + 3 * nt - 1) / (3 * nt);
        if (R < 2) R = 2;
        if (R > a.TB) R = a.TB;
        a.R = (int)R;
        a.nchunk = (a.TB + a.R - 1) / a.R;
    }
    a.d_TG = make_fastdiv((unsigned)a.TG);
    a.d_nchunk = make_fastdiv((unsigned)a.nchunk);
    a.d_TA = make_fastdiv((unsigned)a.TA);
    a.d_img = make_fastdiv((unsigned)a.img_items);
    a.d_GP = make_fastdiv((unsigned)a.GP);
    a.d_TB = make_fastdiv((unsigned)a.TB);
    a.d_tg = make_fastdiv((unsigned)a.tiles_g);
    a.d_tb = make_fastdiv((unsigned)a.tiles_b);
    return true;
}

}  // namespace

bool tma_available() { return encode_tiled() != nullptr; }
bool make_tensor_map5(void* map, const void* base, int es, long long N, long long C, int A, int B, int L, int bl, int bb, int ba, int bn) {
    return make_map((CUtensorMap*)map, base, es, N, C, A, B, L, bl, bb, ba, bn);
}

// ---- planning -----------------------------------------------------------------------------------
TmaPlan plan_tma(const Geo& g, int mode, int active, int esize, int dtype, bool dense_x, unsigned long long fill, const void* x,
                 const void* out, const void* grad, int sm_count) {
    TmaPlan p;
    memset(&p, 0, sizeof(p));
    p.ok = false;
    if (!encode_tiled()) return p;
    if (g.pad != TS_PAD_ZEROS || !dense_x || fill != 0ull) return p;
    if (g.N * g.C == 0 || g.in_plane == 0 || g.out_plane == 0) return p;
    if (mode != 0 && dtype != TS_F32) return p;          // arithmetic kernels: fp32
    if (mode != 0) esize = 4;
    if (esize != 1 && esize != 2 && esize != 4 && esize != 8) return p;
    const int d = g.dim;
    if (mode == 2)                                       // the backward tiles assume output space == input space
        for (int ax = 0; ax < d; ++ax)
            if (g.lb[ax] != 0 || g.OS[ax] != g.S[ax]) return p;
    if (mode != 0)                                       // a size-1 axis ignores its shift AND its +1 neighbour is the element
        for (int ax = 0; ax < d; ++ax)                   // itself (shifts_kernels.h:40-50), not the zero the copy engine would fill
            if (g.S[ax] == 1) return p;
    const int vec = 16 / esize;
    const int L = g.S[d - 1], OL = g.OS[d - 1];
    const int OB = d >= 2 ? g.OS[d - 2] : 1, OA = d == 3 ? g.OS[0] : 1;
    if (((long long)L * esize) % 16 || ((long long)OL * esize) % 16) return p;   // global strides / output rows: 16-byte multiples
    if (((uintptr_t)x & 15) || ((uintptr_t)out & 15) || ((uintptr_t)grad & 15)) return p;
    if (g.N >= (1ll << 31) || g.C >= (1ll << 31)) return p;
    if (g.in_plane * esize >= (1ll << 40) / (g.C > 0 ? g.C : 1)) return p;       // tensor-map strides < 2^40

    const Tuning& t = tuning();
    const int OGR = OL / vec;
    const int max_groups = 256 / vec - 1;               // box inner extent (TG+1)*vec <= 256 elements
    const int tiles_g = (OGR + max_groups - 1) / max_groups;
    const int TG = (OGR + tiles_g - 1) / tiles_g;
    const int ex = mode != 0 ? 1 : 0;                   // +1 neighbour row / slab in the arithmetic boxes
    const int exb = d >= 2 ? ex : 0, exa = d == 3 ? ex : 0;
    long long TB = OB < 256 - exb ? OB : 256 - exb;
    long long TA = OA < 256 - exa ? OA : 256 - exa;
    auto image_bytes = [&](long long ta, long long tb, int* off_gv, int* off_g2) {
        const long long xbytes = round_up((ta + exa) * (tb + exb) * (TG + 1) * 16, 128);
        const long long gsbytes = round_up(ta * tb * (TG + 1) * 16, 128);
        if (off_gv) *off_gv = (int)xbytes;
        if (off_g2) *off_g2 = (int)(xbytes + gsbytes);
        if (mode != 2) return xbytes;
        return xbytes + gsbytes + (active ? xbytes : gsbytes);
    };
    const long long table_bytes = g.C <= TABLE_MAX_C ? g.C * 24 : 0;
    const long long budget = SMEM_LIMIT - 1024 - table_bytes;
    // defaults from the cfg3 sweep on B200 (tools/tune.py --tma): forward 6 x 28 KB, backward 5 x 42 KB
    // 3-D volumes: few large stages (deep slab tiles re-read fewer +1 neighbour slabs)
    // Ring depth, 1-D / 2-D: re-tuned in round 2 (tools/tma_sweep.py, cfg3, CUDA-graph replays).  Once the producer stopped
    // computing the shift parameters per unit it ran a full ring ahead, and MORE stages in flight made the kernels slower
    // (HBM page locality: forward 293 us with 6 stages, 256 us with 3; sparse backward 409 us with 5, 372 us with 4).
    const int want_stages = t.tma_stages > 0 ? t.tma_stages : d == 3 ? (mode == 2 ? 2 : 3) : (mode == 0 ? 3 : mode == 1 ? 4 : (active ? 5 : 4));
    const long long auto_target = d == 3 ? (mode == 2 ? 108 * 1024 : 72 * 1024) : (mode == 2 ? 42 * 1024 : 28 * 1024);
    const long long target = t.tma_stage_kb > 0 ? (long long)t.tma_stage_kb * 1024
                                                : (auto_target < budget / want_stages - 64 ? auto_target : budget / want_stages - 64);
    while (image_bytes(TA, TB, nullptr, nullptr) > target) {
        if (TA > 1) TA = (TA + 1) / 2;
        else if (TB > 1) TB = (TB + 1) / 2;
        else return p;
    }
    const int tiles_a = (int)((OA + TA - 1) / TA), tiles_b = (int)((OB + TB - 1) / TB);
    const long long tiles = (long long)tiles_a * tiles_b * tiles_g;
    if (tiles > 0x7fffffffLL) return p;
    long long np = 1;
    if (tiles == 1) {
        // several images per stage only through a rank-5 box; the sub-box offsets then scale with np
        np = target / image_bytes(TA, TB, nullptr, nullptr);
        if (np < 1) np = 1;
        if (np > 256) np = 256;
        if (np > g.N) np = g.N;
    }
    int off_gv = 0, off_g2 = 0;
    // with np images the three regions are [np x-boxes][np gv-boxes][np g2-boxes]
    const long long xbytes1 = (TA + exa) * (TB + exb) * (TG + 1) * 16, gsbytes1 = TA * TB * (TG + 1) * 16;
    const long long xreg = round_up(np * xbytes1, 128), gsreg = round_up(np * gsbytes1, 128);
    off_gv = (int)xreg;
    off_g2 = (int)(xreg + gsreg);
    long long stage_bytes = xreg, tx = np * xbytes1;
    if (mode == 2) {
        stage_bytes += gsreg + (active ? xreg : gsreg);
        tx += np * gsbytes1 + np * (active ? xbytes1 : gsbytes1);
    }
    const long long stride = round_up(stage_bytes, 1024);
    long long stages = budget / (stride + 16);
    if (stages > want_stages) stages = want_stages;
    if (stages < 2) return p;
    if (tx >= (1 << 20)) return p;                       // mbarrier tx-count range

    const long long planes = g.N * g.C;
    const long long grid_max = (long long)sm_count;
    long long npu = t.chunk_planes > 0 ? t.chunk_planes : pick_unit_images(g.N, g.C, np, grid_max, mode == 2 ? 0.4 : 0.1);
    npu = (npu / np) * np;
    if (npu < np) npu = np;
    if (npu > g.N) npu = g.N;
    const long long chunks = (g.N + npu - 1) / npu;
    const long long units = chunks * g.C;
    if (units > 0x7fffffffLL) return p;
    (void)planes;

    // padded row length of the item index space: a multiple of 8 groups keeps every quarter-warp inside
    // one box row (conflict-free 128-bit shared loads) -- used when it idles <= 15 % of the lanes
    int GP = TG;
    if (TG % 8 != 0 && (double)TG / (double)((TG + 7) / 8 * 8) >= 0.85) GP = (TG + 7) / 8 * 8;
    const long long plane16 = g.C * (mode == 2 ? g.in_plane : g.out_plane) * esize / 16;
    if (np * plane16 >= 0x7fffffffLL || np * TA * TB * (long long)GP >= 0x7fffffffLL) return p;
    // consumer warps: fill the last pass over a stage's items as well as possible
    const long long items = np * TA * TB * GP;
    const int max_warps = mode == 0 ? MAXT_TMA_GATHER / 32 - 1 : MAXT_TMA_ARITH / 32 - 1;
    int warps = t.tma_warps > 0 ? (t.tma_warps < max_warps ? t.tma_warps : max_warps) : 0;
    if (!warps) {
        auto eff = [&](int w) {
            const long long nt = 32ll * w, passes = (items + nt - 1) / nt;
            return (double)items / (double)(passes * nt);
        };
        // best fill of the last pass over a stage's items, preferring 8..16 warps: on cfg3 more
        // consumer warps never helped (the kernels are HBM-bound) and 25+ were measurably slower
        const int lo = 8, hi = 15;
        int best = lo;
        for (int w = lo; w <= hi && w <= max_warps; ++w)
            if (eff(w) >= eff(best) - 1e-9) best = w;
        int best_all = 8;
        for (int w = 8; w <= max_warps; ++w)
            if (eff(w) >= eff(best_all) - 1e-9) best_all = w;
        warps = eff(best) >= eff(best_all) - 0.08 ? best : best_all;
    }
    if (chunks * warps > 0x7fffffffLL) return p;

    p.ok = true;
    p.ta = (int)TA; p.tb = (int)TB; p.tg = TG; p.gp = GP;
    p.xa = (int)(TA + exa); p.xb = (int)(TB + exb);
    p.tiles_per_plane = (int)tiles;
    p.np = (int)np;
    p.off_gv = off_gv; p.off_g2 = off_g2;
    p.tx_bytes = (int)tx;
    p.stages = (int)stages;
    p.stage_stride = (int)stride;
    p.n_per_unit = (int)npu;
    p.units = (int)units;
    p.grid = (int)(units < grid_max ? units : grid_max);
    p.warps = warps;
    p.slots = (int)(chunks * warps);
    p.smem_bytes = (size_t)(stages * stride + 16 * stages + 64 + table_bytes);
    return p;
}

int tma_gather(const Geo& g, const TmaPlan& p, int wk, const void* x, void* y, int esize, const void* w, int qkind, long long wzp,
               cudaStream_t s) {
    TArgs a;
    make_args(g, p, 0, 0, esize, &a);
    const int d = g.dim;
    a.out = (unsigned char*)y;
    a.w = w;
    a.wk = wk;
    a.qkind = qkind;
    a.wzp = wzp;
    if (!make_map(&a.map_x, x, esize, g.N, g.C, d == 3 ? g.S[0] : 1, d >= 2 ? g.S[d - 2] : 1, g.S[d - 1], (a.TG + 1) * a.vec, a.xb,
                  a.xa, a.np))
        return TS_ERR_UNSUPPORTED;
    return a.TA > 1 ? launch(k_tma_gather<true>, a, p, s) : launch(k_tma_gather<false>, a, p, s);
}

int tma_active_forward(const Geo& g, const TmaPlan& p, const void* x, const void* w, void* y, cudaStream_t s) {
    TArgs a;
    make_args(g, p, 1, 1, 4, &a);
    const int d = g.dim;
    a.out = (unsigned char*)y;
    a.w = w;
    if (!make_map(&a.map_x, x, 4, g.N, g.C, d == 3 ? g.S[0] : 1, d >= 2 ? g.S[d - 2] : 1, g.S[d - 1], (a.TG + 1) * 4, a.xb, a.xa, a.np))
        return TS_ERR_UNSUPPORTED;
    switch (d) {
    case 1: return launch(k_tma_active_forward<1>, a, p, s);
    case 2: return launch(k_tma_active_forward<2>, a, p, s);
    default: return launch(k_tma_active_forward<3>, a, p, s);
    }
}

int tma_backward(const Geo& g, const TmaPlan& p, int active, const void* grad, const void* x, const void* w, void* gi, void* gw,
                 double* partials, const ts_peer_group* peers, cudaStream_t s) {
    TArgs a;
    make_args(g, p, 2, active ? 1 : 0, 4, &a);
    const int d = g.dim;
    a.out = (unsigned char*)gi;
    a.w = w;
    a.partials = partials;
    const int A = d == 3 ? g.S[0] : 1, B = d >= 2 ? g.S[d - 2] : 1, L = g.S[d - 1];
    if (!make_map(&a.map_x, x, 4, g.N, g.C, A, B, L, (a.TG + 1) * 4, a.xb, a.xa, a.np)) return TS_ERR_UNSUPPORTED;
    if (!make_map(&a.map_gs, grad, 4, g.N, g.C, A, B, L, (a.TG + 1) * 4, a.TB, a.TA, a.np)) return TS_ERR_UNSUPPORTED;
    if (active && !make_map(&a.map_gb, grad, 4, g.N, g.C, A, B, L, (a.TG + 1) * 4, a.xb, a.xa, a.np)) return TS_ERR_UNSUPPORTED;
    int rc;
    switch (d * 2 + (active ? 1 : 0)) {
    case 2: rc = launch(k_tma_backward<1, false>, a, p, s); break;
    case 3: rc = launch(k_tma_backward<1, true>, a, p, s); break;
    case 4: rc = launch(k_tma_backward<2, false>, a, p, s); break;
    case 5: rc = launch(k_tma_backward<2, true>, a, p, s); break;
    case 6: rc = launch(k_tma_backward<3, false>, a, p, s); break;
    default: rc = launch(k_tma_backward<3, true>, a, p, s); break;
    }
    if (rc != TS_OK) return rc;
    return launch_reduce_partials<float>(partials, p.slots, (int)(g.C * g.dim), gw, peers, s);
}

}  // namespace ts
