// ts_kernels.h -- internal interface between the C ABI (ts_capi.cu) and the kernel families.
#pragma once
#include "ts_common.cuh"

namespace ts {

// How the integer shifts are obtained in the sparse/quantized gather kernels.
enum { WK_F32 = 0, WK_F64 = 1, WK_F16 = 2, WK_BF16 = 3, WK_QUANT = 4 };

template <int WK> struct WType;
template <> struct WType<WK_F32> { using type = float; };
template <> struct WType<WK_F64> { using type = double; };
template <> struct WType<WK_F16> { using type = __half; };
template <> struct WType<WK_BF16> { using type = __nv_bfloat16; };

template <int WK, int DIM>
TS_D void load_int_shifts(const void* __restrict__ w, int qkind, long long wzp, long long c, const Geo& g, int* sx) {
    if constexpr (WK == WK_QUANT) {
        load_qshifts<DIM>(w, qkind, wzp, c, g, sx);
    } else {
        using ST = typename WType<WK>::type;
        const ShiftParams<typename Elem<ST>::CT, DIM> p = load_params<ST, DIM>((const ST*)w, c, g, false, false);
#pragma unroll
        for (int a = 0; a < DIM; ++a) sx[a] = p.sx[a];
    }
}

// Images per work unit (a multiple of `np`, the images of one stage).  A unit has a fixed cost (parameter fetch, the
// pipeline restart of its first stage, for the backward the shuffle tree and the partial store: measured ~0.4 plane
// times in the 2-D backward, ~0.1 in the forward), the last round of units leaves CTAs idle: pick the size that
// minimises  rounds * (images + unit_cost)  -- large units for small per-GPU batches, where one-image units made the
// 32-image shard of the 8-GPU strong-scaling run 60 % slower than its HBM floor.
inline long long pick_unit_images(long long N, long long C, long long np, long long grid, double unit_cost) {
    long long best = np < N ? np : N;
    double best_cost = 1e300;
    const long long kmax = (N + np - 1) / np < 4096 ? (N + np - 1) / np : 4096;
    for (long long k = 1; k <= kmax; ++k) {
        const long long npu = k * np < N ? k * np : N;
        const long long chunks = (N + npu - 1) / npu, units = chunks * C, rounds = (units + grid - 1) / grid;
        const double cost = (double)rounds * ((double)npu + unit_cost);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = npu; }
    }
    return best;
}

// ---- generic family (ts_generic.cu) ----------------------------------------------------------
struct GenericBwdPlan { int threads, tiles, n_per_chunk, units; };
GenericBwdPlan plan_generic_backward(const Geo& g);

int generic_gather(const Geo& g, int wk, const void* x, void* y, unsigned long long fill, int esize, const void* w,
                   int qkind, long long wzp, cudaStream_t s);
int generic_active_forward(const Geo& g, int dtype, const void* x, const void* w, void* y, cudaStream_t s);
int generic_backward(const Geo& g, int dtype, int active, const void* grad, const void* x, const void* w, void* gi,
                     void* gw, double* partials, const ts_peer_group* peers, cudaStream_t s);

// peers != nullptr: the pass-2 launch is the variant fused with the all-reduce over peer memory
// (ts_shift_backward_allreduce); slots may be 0 (a rank with an empty shard contributes zeros).
template <typename ST>
int launch_reduce_partials(const double* partials, int slots, int outputs, void* gw, const ts_peer_group* peers, cudaStream_t stream);

// ---- channels-last gather (ts_nhwc.cu): input and output keep the channel axis innermost -------
int nhwc_gather(const Geo& g, const void* x, void* y, unsigned long long fill, int esize, const void* w, int qkind,
                long long wzp, int sm_count, int max_grid_x, int variant, int ring_rows, bool emulate, cudaStream_t s);

int nhwc_to_planar(const void* x, void* y, long long N, long long C, long long P, int esize, cudaStream_t s);

// ---- staged family (ts_staged.cu): bulk-async shared-memory staging ---------------------------
struct Tuning {
    int stages;        // ring depth
    int stage_kb;      // target bytes per stage (KiB)
    int warps;         // consumer warps per CTA
    int ctas_per_sm;   // persistent CTAs per SM
    int chunk_planes;  // planes of one channel per work unit (0 = auto)
    int tma_stages;    // ring depth of the TMA-tensor kernels (0 = auto)
    int tma_ctas_per_sm;
    int tma_warps;     // consumer warps of the TMA-tensor arithmetic kernels
    int use_tma;       // 0: the automatic path choice never picks the TMA-tensor family
    int tma_stage_kb;  // target bytes per stage of the TMA-tensor kernels (0 = auto)
    int nhwc_variant;  // channels-last gather: 0 auto, 1 direct (L1) kernel only, 2 ring (shared-memory) kernel only
    int nhwc_rows_warps; // consumer warps of the row-pipelined channels-last kernel (0 = auto)
    int nhwc_ring_rows;  // cap on the ring slots of the channels-last ring kernel (0 = as many as fit)
    int use_halo;      // 0: the automatic path choice never picks the halo family
    int halo;          // halo rows / columns of the halo family (0 = 4)
    int halo_stages;   // ring depth of the halo family (0 = auto)
    int halo_warps;    // consumer warps of the 2-D halo kernels (0 = auto)
    int halo_probe;    // measurement only (results are WRONG when set): 1 = 3-D consumers skip loads + arithmetic (pipeline rate), 2 = consumers do not wait for the data (compute rate)
    int halo_compact;  // 3-D: 1 = pairs packed without padding the groups of a row to a multiple of 8 when that saves a consumer warp
    int halo_split;    // 3-D interpolating backward: 0 (default) = one pair per thread, 1 = x-window warps + grad-window warps (two pairs per thread)
    int unit_order;    // 1 (default): units dealt round-robin over (chunk, channel); 0: contiguous channel-major ranges per CTA
                       // (measured on B200: 4 % faster for 32-image shards, 5 % slower for cfg3 at N=256)
    int no_pdl;        // 1: plain launches instead of programmatic dependent launches (TMA family, pass 2)
    int no_table;      // 1: do not tabulate the per-channel shift parameters in shared memory (A/B measurements)
    int use_flat;      // 0: the automatic path choice never picks the flat zero-padding gather
    int flat_variant;  // 0 auto (a warp owns whole planes when C <= 512), 1 item-per-thread with shuffled shifts
    int flat_ctas, flat_stage_kb, flat_stages, flat_warps;   // flat gather: CTAs per SM (0 = 2), stage KiB (0 = 24), ring depth (0 = 4), consumer warps (0 = auto)
};
Tuning& tuning();

struct StagedPlan {
    bool ok;           // staged path applicable
    int vec_bytes;     // bytes per thread item (16 / 8 / 4)
    int planes_per_step, stages, stage_stride, warps, grid, units, n_per_unit, slots;  // slots = partial slots (backward)
    bool table;        // per-channel shift parameters tabulated in shared memory
    int ta, tiles;     // slabs per tile (3-D volumes are tiled over their first axis), tiles per image
    int xs, gvs, gis;  // slab slots per image: x, grad at the output position, grad for grad_input (0: shares gvs)
    int gp;            // padded items per row of the item index space
    int off_gv, off_gi;
    size_t smem_bytes;
};
// mode: 0 sparse/quantized forward (esize = element bytes), 1 active forward, 2 backward
StagedPlan plan_staged(const Geo& g, int mode, int active, int esize, int dtype, bool dense_x, const void* x, const void* y_or_gi,
                       const void* grad, int sm_count);

int staged_gather(const Geo& g, const StagedPlan& p, int wk, const void* x, void* y, unsigned long long fill, int esize,
                  const void* w, int qkind, long long wzp, cudaStream_t s);
int staged_active_forward(const Geo& g, const StagedPlan& p, int dtype, const void* x, const void* w, void* y, cudaStream_t s);
int staged_backward(const Geo& g, const StagedPlan& p, int dtype, int active, const void* grad, const void* x, const void* w,
                    void* gi, void* gw, double* partials, const ts_peer_group* peers, cudaStream_t s);

// ---- TMA-tensor family (ts_tma.cu): zeros padding done by the copy engine ----------------------
struct TmaPlan {
    bool ok;
    int ta, tb, tg, gp;        // tile extents: slabs, rows, 16-byte column groups; gp = padded groups per row (index space)
    int xa, xb;                // x box slabs / rows (tile + the +1 neighbours of the arithmetic kernels)
    int tiles_per_plane, np;   // np = images per stage
    int off_gv, off_g2, tx_bytes;
    int stages, stage_stride, n_per_unit, units, grid, warps, slots;
    size_t smem_bytes;
};
// mode: 0 sparse/quantized forward, 1 active forward, 2 backward (active = interpolating backward)
TmaPlan plan_tma(const Geo& g, int mode, int active, int esize, int dtype, bool dense_x, unsigned long long fill, const void* x,
                 const void* out, const void* grad, int sm_count);
int tma_gather(const Geo& g, const TmaPlan& p, int wk, const void* x, void* y, int esize, const void* w, int qkind, long long wzp,
               cudaStream_t s);
int tma_active_forward(const Geo& g, const TmaPlan& p, const void* x, const void* w, void* y, cudaStream_t s);
int tma_backward(const Geo& g, const TmaPlan& p, int active, const void* grad, const void* x, const void* w, void* gi, void* gw,
                 double* partials, const ts_peer_group* peers, cudaStream_t s);

// rank-5 tensor map over a dense [N][C][A][B][L] tensor, box {bl, bb, ba, 1, bn} (memoised per thread); `map` points to a
// CUtensorMap (128 bytes, 64-byte aligned)
bool tma_available();
bool make_tensor_map5(void* map, const void* base, int es, long long N, long long C, int A, int B, int L, int bl, int bb, int ba, int bn);

// ---- flat gather (ts_flat.cu): zeros padding, no crop -> a linear shifted copy of each dense slab + byte masks ----
struct FlatPlan {
    bool ok;
    int np, stages, stage_stride, warps, n_per_unit, units, grid;
    size_t smem_bytes;
};
FlatPlan plan_flat(const Geo& g, int esize, bool dense_x, const void* x, const void* y, int sm_count, bool forced);
int flat_gather(const Geo& g, const FlatPlan& p, int wk, const void* x, void* y, unsigned long long fill, int esize, const void* w,
                int qkind, long long wzp, cudaStream_t s);

// ---- halo family (ts_halo.cu): whole slabs staged WITH a padded halo, every item is interior ----
struct HaloPlan {
    bool ok;
    int hr, hc;                 // halo rows / columns on each side of a staged slab
    int px, bpx, pg, bpg;       // row pitch (elements) and rows of the x / grad tiles in shared memory
    int tile_x, tile_g, tile_v; // bytes per tile (128-byte multiples); tile_v: dense grad slab (3-D backward)
    int tile_p;                 // pooled-gradient tile of the fused avg-pool backward (0 otherwise)
    int np, GP, positions, ncol;
    int stages, stage_stride, warps, n_per_unit, units, grid, slots;
    size_t smem_bytes;
};
// mode: 0 sparse forward (2-D only), 1 active forward, 2 backward (active = interpolating backward).  fp32, dims 2 and 3,
// every padding, border crops.
// pool_bwd (mode 2, 2-D): `grad` is the gradient of the 2x2 / stride-2 / ceil_mode average pooling of the shift's output;
// the kernel expands it while staging (ts_shift2d_avgpool2_backward).
HaloPlan plan_halo(const Geo& g, int mode, int active, int dtype, bool dense_x, const void* x, const void* out, const void* grad,
                   int sm_count, bool forced, bool pool_bwd = false);
int halo_active_forward(const Geo& g, const HaloPlan& p, const void* x, const void* w, void* y, cudaStream_t s);
int halo_forward2d(const Geo& g, const HaloPlan& p, int active, int pool, const void* x, const void* w, void* y, cudaStream_t s);
int halo_backward(const Geo& g, const HaloPlan& p, int active, const void* grad, const void* x, const void* w, void* gi, void* gw,
                  double* partials, const ts_peer_group* peers, cudaStream_t s, bool pool_bwd = false);

}  // namespace ts
