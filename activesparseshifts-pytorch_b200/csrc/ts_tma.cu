// ts_tma.cu -- zero-padding fast path built on TMA *tensor* copies (cp.async.bulk.tensor, SASS
// UTMALDG).  Idea: a tiled TMA load whose box starts at (possibly negative) coordinates
// (col - s_col, row - s_row, slab - s_slab, plane) delivers the plane already SHIFTED, with every
// out-of-bounds element filled with zero by the copy engine.  That is exactly the reference's
// zeros-padded integer gather (ops/kernels/shifts_kernels.h:10-54 with BIPadding::Zeros), done
// by hardware:
//
//   * sparse forward (and grad_input of the sparse backward without borders): TMA load of the
//     shifted tile -> 1-D bulk store of the tile to the dense output.  ONE thread per CTA drives a
//     ring of stages; no SM instruction touches the data.
//   * backward / active forward: tiles arrive pre-shifted and 16-byte aligned, with one extra
//     column group / row / slab for the +1 neighbours, so consumers use aligned LDS.128 only: no
//     masks, no funnel shifts, no index remapping.
//
// Applicability (plan_tma): zeros padding, dense NCHW x, row bytes multiple of 16, every box
// extent <= 256, pad value 0 (so not qint8 with a non-zero zero point).  Everything else runs on
// ts_staged.cu / ts_generic.cu.
#include <cuda.h>

#include "ts_kernels.h"

namespace ts {

namespace {

constexpr int SMEM_LIMIT = 232448;

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

// rank-4 map over a dense [planes][A][B][L] tensor of `es`-byte elements with box {bl, bb, ba, 1}
bool make_map(CUtensorMap* map, const void* base, int es, long long planes, int A, int B, int L, int bl, int bb, int ba) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const CUtensorMapDataType dt = es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                 : es == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    cuuint64_t dims[4] = {(cuuint64_t)L, (cuuint64_t)B, (cuuint64_t)A, (cuuint64_t)planes};
    cuuint64_t strides[3] = {(cuuint64_t)L * es, (cuuint64_t)L * B * es, (cuuint64_t)L * B * A * es};
    cuuint32_t box[4] = {(cuuint32_t)bl, (cuuint32_t)bb, (cuuint32_t)ba, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
// tiled 4-D TMA load global -> shared; coordinates innermost first; OOB elements arrive as zero
TS_D void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk store shared -> global (dense destination), tracked by bulk async-groups
TS_D void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
TS_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> TS_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> TS_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- arguments -------------------------------------------------------------------------------
struct alignas(64) TArgs {
    CUtensorMap map_x;       // source of the shifted tiles
    CUtensorMap map_g;       // grad (backward)
    Geo g;
    unsigned char* out;
    const void* w;
    double* partials;
    long long wzp;
    int qkind, wk, es;
    int A, B, L, OA, OB, OL, lbA, lbB, lbL;
    int ta;                  // output slabs per tile (== OA unless 3-D volumes are tiled)
    int tiles_per_plane;
    int stages, stage_stride, np;
    int n_per_unit, units, nw;
    int tile_bytes;          // bytes of one shifted output tile (box bytes)
};

TS_D int level_axis(int level, int dim) { return level - (3 - dim); }

TS_D long long raw_int_shift(const TArgs& a, long long idx) {
    long long iw;
    switch (a.wk) {
    case WK_F32: { float d; split_forward<float>(((const float*)a.w)[idx], false, iw, d); return iw; }
    case WK_F64: { double d; split_forward<double>(((const double*)a.w)[idx], false, iw, d); return iw; }
    case WK_F16: { float d; split_forward<float>(__half2float(((const __half*)a.w)[idx]), false, iw, d); return iw; }
    case WK_BF16: { float d; split_forward<float>(__bfloat162float(((const __nv_bfloat16*)a.w)[idx]), false, iw, d); return iw; }
    default:
        if (a.qkind == TS_QW_U8) return (long long)((const uint8_t*)a.w)[idx] - a.wzp;
        if (a.qkind == TS_QW_I8) return (long long)((const int8_t*)a.w)[idx] - a.wzp;
        return (long long)((const int32_t*)a.w)[idx] - a.wzp;
    }
}

// ================================================================================================
// Sparse forward, zeros padding: the copy engines do everything.  One thread per CTA.
//   step q: TMA-load tile(q) into stage q % S;  after it lands, bulk-store it to y.
// The loads run D = S-2 steps ahead of the stores; a stage is reloaded only after the store that
// read it has finished reading shared memory (cp.async.bulk.wait_group.read 1).
__global__ void __launch_bounds__(32, 1) k_tma_shiftcopy(const __grid_constant__ TArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = (uint64_t*)(smem + (size_t)a.stages * a.stage_stride);
    if (threadIdx.x != 0) return;
    for (int s = 0; s < a.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();

    const int S = a.stages, D = S - 2;
    const long long C = a.g.C, N = a.g.N;
    const int dim = a.g.dim;
    // two cursors over the same step sequence: `ld` (loads) runs D steps ahead of `st` (stores)
    struct Cursor { int u; long long n, n1, c; int tile; int sh[3]; bool valid; };
    auto open_unit = [&](Cursor& k) {
        if (k.u >= a.units) { k.valid = false; return; }
        k.c = k.u % C;
        const long long chunk = k.u / C;
        k.n = chunk * a.n_per_unit;
        k.n1 = k.n + a.n_per_unit < N ? k.n + a.n_per_unit : N;
        k.tile = 0;
        for (int lev = 0; lev < 3; ++lev) {
            const int ax = level_axis(lev, dim);
            k.sh[lev] = ax >= 0 ? reduce_shift(raw_int_shift(a, k.c * dim + ax), a.g.S[ax], TS_PAD_ZEROS) : 0;
        }
        k.valid = true;
    };
    auto advance = [&](Cursor& k) {
        if (++k.tile < a.tiles_per_plane) return;
        k.tile = 0;
        if (++k.n < k.n1) return;
        k.u += gridDim.x;
        open_unit(k);
    };
    Cursor ld, st;
    ld.u = st.u = blockIdx.x;
    open_unit(ld);
    open_unit(st);
    int q_ld = 0, q_st = 0;
    const long long tile_out_elems = (long long)a.ta * a.OB * a.OL;
    while (st.valid) {
        // issue loads until D steps ahead
        while (ld.valid && q_ld < q_st + D + 1) {
            const int s = q_ld % S;
            if (q_ld >= S) bulk_wait_read<1>();          // the store that last read stage s is done reading
            mbar_expect_tx(&full[s], (unsigned)a.tile_bytes);
            const int a0 = ld.tile * a.ta;
            tma_load_4d(smem + (size_t)s * a.stage_stride, &a.map_x, a.lbL - ld.sh[2], a.lbB - ld.sh[1], a0 + a.lbA - ld.sh[0],
                        (int)(ld.n * C + ld.c), &full[s]);
            ++q_ld;
            advance(ld);
        }
        // store the oldest landed tile
        {
            const int s = q_st % S;
            mbar_wait(&full[s], (unsigned)((q_st / S) & 1));
            const int a0 = st.tile * a.ta;
            const int slabs = a.OA - a0 < a.ta ? a.OA - a0 : a.ta;
            unsigned char* dst = a.out + ((st.n * C + st.c) * a.g.out_plane + (long long)a0 * a.OB * a.OL) * a.es;
            bulk_s2g(dst, smem + (size_t)s * a.stage_stride, (unsigned)((long long)slabs * a.OB * a.OL * a.es));
            bulk_commit();
            ++q_st;
            advance(st);
            (void)tile_out_elems;
        }
    }
    bulk_wait_all<0>();
}

}  // namespace

// ---- planning -----------------------------------------------------------------------------------
TmaPlan plan_tma(const Geo& g, int mode, int esize, int dtype, bool dense_x, unsigned long long fill, const void* x,
                 const void* out, const void* grad, int sm_count) {
    TmaPlan p;
    memset(&p, 0, sizeof(p));
    p.ok = false;
    if (!encode_tiled()) return p;
    if (g.pad != TS_PAD_ZEROS || !dense_x || fill != 0ull) return p;
    if (g.N * g.C == 0 || g.in_plane == 0 || g.out_plane == 0) return p;
    if (mode != 0) return p;                     // (backward / active forward tiles: added below as they land)
    const int d = g.dim;
    const int L = g.S[d - 1], OL = g.OS[d - 1];
    const int B = d >= 2 ? g.S[d - 2] : 1, OB = d >= 2 ? g.OS[d - 2] : 1;
    const int A = d == 3 ? g.S[0] : 1, OA = d == 3 ? g.OS[0] : 1;
    if (((long long)L * esize) % 16 || ((long long)OL * esize) % 16) return p;      // global strides / box rows: 16-byte multiples
    if (OL > 256 || OB > 256) return p;
    if (((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return p;
    if (g.N * g.C >= (1ll << 31)) return p;
    // slabs per tile: whole volume if it fits a stage of <= 48 KB, else as many slabs as fit
    const long long slab_bytes = (long long)OB * OL * esize;
    long long ta = OA;
    const long long target = 48 * 1024;
    if (ta * slab_bytes > target) ta = target / slab_bytes;
    if (ta < 1) ta = 1;
    if (ta > 256) ta = 256;
    if (ta * slab_bytes > 100 * 1024) return p;
    const Tuning& t = tuning();
    const long long stride = ((ta * slab_bytes + 127) / 128) * 128;
    int ctas = t.tma_ctas_per_sm > 0 ? t.tma_ctas_per_sm : 2;
    long long stages = t.tma_stages > 0 ? t.tma_stages : 8;
    const long long budget = SMEM_LIMIT / ctas - 1024;
    if (stages * stride + 8 * stages > budget) stages = budget / (stride + 8);
    if (stages < 3) { ctas = 1; stages = (SMEM_LIMIT - 1024) / (stride + 8); }
    if (stages < 3) return p;
    if (stages > 16) stages = 16;
    const long long planes = g.N * g.C;
    const long long grid_max = (long long)sm_count * ctas;
    long long npu = t.chunk_planes > 0 ? t.chunk_planes : planes / (grid_max * 32);
    if (npu < 1) npu = 1;
    if (npu > g.N) npu = g.N;
    const long long chunks = (g.N + npu - 1) / npu;
    const long long units = chunks * g.C;
    if (units > 0x7fffffffLL) return p;
    p.ok = true;
    p.ta = (int)ta;
    p.tiles_per_plane = (int)((OA + ta - 1) / ta);
    p.stages = (int)stages;
    p.stage_stride = (int)stride;
    p.n_per_unit = (int)npu;
    p.units = (int)units;
    p.grid = (int)(units < grid_max ? units : grid_max);
    p.smem_bytes = (size_t)(stages * stride + 8 * stages + 64);
    (void)A; (void)B; (void)grad; (void)dtype;
    return p;
}

int tma_gather(const Geo& g, const TmaPlan& p, int wk, const void* x, void* y, int esize, const void* w, int qkind, long long wzp,
               cudaStream_t s) {
    TArgs a;
    memset(&a, 0, sizeof(a));
    const int d = g.dim;
    a.g = g;
    a.es = esize;
    a.A = d == 3 ? g.S[0] : 1;        a.OA = d == 3 ? g.OS[0] : 1;      a.lbA = d == 3 ? g.lb[0] : 0;
    a.B = d >= 2 ? g.S[d - 2] : 1;    a.OB = d >= 2 ? g.OS[d - 2] : 1;  a.lbB = d >= 2 ? g.lb[d - 2] : 0;
    a.L = g.S[d - 1];                 a.OL = g.OS[d - 1];               a.lbL = g.lb[d - 1];
    a.ta = p.ta;
    a.tiles_per_plane = p.tiles_per_plane;
    a.stages = p.stages;
    a.stage_stride = p.stage_stride;
    a.n_per_unit = p.n_per_unit;
    a.units = p.units;
    a.tile_bytes = p.ta * a.OB * a.OL * esize;
    a.out = (unsigned char*)y;
    a.w = w;
    a.wk = wk;
    a.qkind = qkind;
    a.wzp = wzp;
    if (!make_map(&a.map_x, x, esize, g.N * g.C, a.A, a.B, a.L, a.OL, a.OB, p.ta)) return TS_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(k_tma_shiftcopy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return check_launch();
    k_tma_shiftcopy<<<p.grid, 32, p.smem_bytes, s>>>(a);
    note_launch();
    return check_launch();
}

}  // namespace ts
