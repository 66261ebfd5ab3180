// ts_staged.cu -- the bandwidth path: persistent, warp-specialised kernels that stage whole
// (n,c) planes in shared memory with 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) signalled
// through mbarriers, resolve the per-channel shift while READING shared memory, and write global
// memory with 128-bit streaming stores.
//
//   CTA = `nw` consumer warps + 1 producer warp (one elected lane), one CTA per SM, persistent.
//   Work unit = (channel c, chunk of the batch); units are dealt round-robin to CTAs.  Inside a
//   unit the shift parameters are registers and (backward) the grad_weight terms accumulate in
//   per-thread registers; every consumer warp writes ONE partial per unit -> deterministic.
//   Step = up to `np` planes of the unit = one ring stage.  full[s]: producer's expect_tx +
//   the copies' complete_tx.  empty[s]: one arrive per consumer warp.  No __syncthreads in the loop.
//
// Planes are contiguous in NCHW, so a plane (or a plane of grad) is ONE bulk copy, 16-byte
// aligned and a multiple of 16 bytes (checked by plan_staged; otherwise the generic path runs).
// A shift along the fastest axis breaks 16-byte alignment of the source: the consumer loads the
// two aligned 16-byte groups that cover its item from shared memory (conflict-free LDS.128) and
// funnel-shifts by the (unit-uniform) misalignment.  Row shifts are just a different row index.
//
// Semantics: identical to ts_generic.cu / the reference (ops/kernels/shifts_kernels.h:156-327,
// :532-571); the arithmetic helpers are the same functions (ts_common.cuh), so active forward and
// grad_input are bit-identical to the generic path and to the CPU reference.
#include "ts_kernels.h"

namespace ts {

Tuning& tuning() {
    static Tuning t = {4, 48, 16, 1, 0, 0, 0, 0, 1, 0};
    return t;
}

namespace {

constexpr int SMEM_LIMIT = 232448;   // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr int GUARD = 16;            // readable slack before/after the planes of a stage

// ---- exact division by a launch-invariant (Granlund-Montgomery, n < 2^31) -------------------
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
// 1-D bulk async copy global -> shared, completion reported to an mbarrier (SASS: UBLKCP)
TS_D void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
TS_D void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- kernel arguments ---------------------------------------------------------------------------
struct SArgs {
    Geo g;
    const unsigned char* x;
    const unsigned char* grad;
    unsigned char* out;
    const void* w;
    double* partials;
    long long wzp;
    unsigned long long fill;
    int qkind, wk, es;
    int xpb, gpb;                  // bytes of one staged x plane / grad plane (0 unless backward)
    int np, stages, stage_stride, nw;
    int n_per_unit, units;
    int A, B, L, OA, OB, OL, lbA, lbB, lbL;   // slab / row / column structure (absent levels: 1, 0)
    int gpr, ipp;                  // items per row, items per plane
    FastDiv d_ipp, d_gpr, d_rows;
};

// level (0 slab, 1 row, 2 col) -> tensor axis, or -1 when the level is absent for this dim
TS_D int level_axis(int level, int dim) { return level - (3 - dim); }

// ---- producer ---------------------------------------------------------------------------------
TS_D void producer(const SArgs& a, unsigned char* smem, uint64_t* full, uint64_t* empty) {
    int s = 0, k = 0;
    const long long C = a.g.C, N = a.g.N;
    for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
        const long long c = u % C, chunk = u / C;
        const long long n0 = chunk * a.n_per_unit;
        const long long n1 = n0 + a.n_per_unit < N ? n0 + a.n_per_unit : N;
        for (long long nb = n0; nb < n1; nb += a.np) {
            if (k > 0) mbar_wait(&empty[s], (unsigned)((k - 1) & 1));
            const int npl = (int)(n1 - nb < a.np ? n1 - nb : a.np);
            unsigned char* st = smem + (size_t)s * a.stage_stride + GUARD;
            mbar_expect_tx(&full[s], (unsigned)(npl * (a.xpb + a.gpb)));
            for (int pl = 0; pl < npl; ++pl) {
                const long long plane = (nb + pl) * C + c;
                bulk_g2s(st + (size_t)pl * a.xpb, a.x + plane * a.xpb, (unsigned)a.xpb, &full[s]);
                if (a.gpb)
                    bulk_g2s(st + (size_t)a.np * a.xpb + (size_t)pl * a.gpb, a.grad + plane * a.gpb, (unsigned)a.gpb, &full[s]);
            }
            if (++s == a.stages) { s = 0; ++k; }
        }
    }
}

// ---- consumer skeleton -----------------------------------------------------------------------
template <class Body>
TS_D void consumer_loop(const SArgs& a, unsigned char* smem, uint64_t* full, uint64_t* empty, int lane, Body& body) {
    int s = 0, k = 0;
    const long long C = a.g.C, N = a.g.N;
    for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
        const long long c = u % C, chunk = u / C;
        const long long n0 = chunk * a.n_per_unit;
        const long long n1 = n0 + a.n_per_unit < N ? n0 + a.n_per_unit : N;
        body.begin_unit(c);
        for (long long nb = n0; nb < n1; nb += a.np) {
            mbar_wait(&full[s], (unsigned)(k & 1));
            const int npl = (int)(n1 - nb < a.np ? n1 - nb : a.np);
            body.step(smem + (size_t)s * a.stage_stride + GUARD, npl, nb, c);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == a.stages) { s = 0; ++k; }
        }
        body.end_unit(c, chunk);
    }
}

// ---- word-vector helpers ----------------------------------------------------------------------
template <int G> TS_D void lds_words(const unsigned char* p, unsigned* w);
template <> TS_D void lds_words<4>(const unsigned char* p, unsigned* w) { const uint4 v = *(const uint4*)p; w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
template <> TS_D void lds_words<2>(const unsigned char* p, unsigned* w) { const uint2 v = *(const uint2*)p; w[0] = v.x; w[1] = v.y; }
template <> TS_D void lds_words<1>(const unsigned char* p, unsigned* w) { w[0] = *(const unsigned*)p; }
template <int G> TS_D void stg_words(unsigned char* p, const unsigned* w);
template <> TS_D void stg_words<4>(unsigned char* p, const unsigned* w) { __stcs((uint4*)p, make_uint4(w[0], w[1], w[2], w[3])); }
template <> TS_D void stg_words<2>(unsigned char* p, const unsigned* w) { __stcs((uint2*)p, make_uint2(w[0], w[1])); }
template <> TS_D void stg_words<1>(unsigned char* p, const unsigned* w) { __stcs((unsigned*)p, w[0]); }

// ================================================================================================
// Sparse / quantized forward: a pure byte mover.  G = 32-bit words per item (4: 128-bit stores),
// ES = element bytes.  dim / padding / weight kind are run-time (they only touch per-unit setup,
// the row remap and the rare edge path).
template <int G, int ES>
struct GatherBody {
    static constexpr int VB = 4 * G, VEC = VB / ES;
    const SArgs& a;
    const int tid, nt;
    int sh[3];        // reduced shifts per level
    int mb;           // byte misalignment of the source inside its 4G-byte group (unit-uniform)
    unsigned fillw[G];

    TS_D GatherBody(const SArgs& a_, int tid_, int nt_) : a(a_), tid(tid_), nt(nt_) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            if (ES == 1) fillw[k] = 0x01010101u * (unsigned)(a.fill & 0xffu);
            else if (ES == 2) fillw[k] = 0x00010001u * (unsigned)(a.fill & 0xffffu);
            else if (ES == 4) fillw[k] = (unsigned)a.fill;
            else fillw[k] = (k & 1) ? (unsigned)(a.fill >> 32) : (unsigned)a.fill;
        }
    }

    TS_D long long raw_shift(long long idx) const {
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

    TS_D void begin_unit(long long c) {
        const int dim = a.g.dim;
#pragma unroll
        for (int lev = 0; lev < 3; ++lev) {
            const int ax = level_axis(lev, dim);
            sh[lev] = ax >= 0 ? reduce_shift(raw_shift(c * dim + ax), a.g.S[ax], a.g.pad) : 0;
        }
        mb = pmod((a.lbL - sh[2]) * ES, VB);
    }
    TS_D void end_unit(long long, long long) {}

    TS_D void step(const unsigned char* st, int npl, long long nb, long long c) {
        const int pad = a.g.pad;
        const int total = npl * a.ipp;
        const int ws = mb >> 2, bs8 = (mb & 3) * 8;
        for (int item = tid; item < total; item += nt) {
            const int pl = (int)fdiv((unsigned)item, a.d_ipp);
            const int rem = item - pl * a.ipp;
            const int orow = (int)fdiv((unsigned)rem, a.d_gpr);
            const int cg = rem - orow * a.gpr;
            int oa = 0, ob = orow;
            if (a.OA > 1 || a.A > 1) { oa = (int)fdiv((unsigned)orow, a.d_rows); ob = orow - oa * a.OB; }
            const int ra = axis_index(oa + a.lbA - sh[0], a.A, pad);
            const int rb = axis_index(ob + a.lbB - sh[1], a.B, pad);
            const int cs = cg * VEC + a.lbL - sh[2];
            unsigned char* dst = a.out + (((nb + pl) * a.g.C + c) * a.g.out_plane + (long long)orow * a.OL) * ES + (long long)cg * VB;
            unsigned o[G];
            const bool interior = cs >= 0 && cs + VEC <= a.L;
            if (ra < 0 || rb < 0 || cs + VEC <= 0 || cs >= a.L) {
                // nothing of this item lies inside the source (only possible with zeros padding,
                // or a column window fully outside which non-zero paddings resolve element-wise)
                if (pad == TS_PAD_ZEROS || ra < 0 || rb < 0) {
#pragma unroll
                    for (int k = 0; k < G; ++k) o[k] = fillw[k];
                    stg_words<G>(dst, o);
                    continue;
                }
            }
            const unsigned char* pb = st + (size_t)pl * a.xpb;
            const long long rowoff = ((long long)ra * a.B + rb) * a.L;
            if (interior || pad == TS_PAD_ZEROS) {
                const long long bo = (rowoff + cs) * ES;
                const unsigned char* ag = pb + (bo - mb);          // aligned to VB by construction
                unsigned W[2 * G];
                lds_words<G>(ag, W);
                lds_words<G>(ag + VB, W + G);
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    unsigned lo = W[k], hi = W[k + 1];
#pragma unroll
                    for (int j = 1; j < G; ++j)
                        if (ws == j) { lo = W[k + j]; hi = (k + j + 1 < 2 * G) ? W[k + j + 1] : 0u; }
                    o[k] = __funnelshift_r(lo, hi, bs8);
                }
                if (!interior) {   // zeros padding, partially outside: blend the pad value in
                    const int lo_b = (cs < 0 ? -cs : 0) * ES;
                    const int hi_b = (a.L - cs < VEC ? a.L - cs : VEC) * ES;
#pragma unroll
                    for (int k = 0; k < G; ++k) {
                        int l = lo_b - 4 * k, h = hi_b - 4 * k;
                        l = l < 0 ? 0 : (l > 4 ? 4 : l);
                        h = h < 0 ? 0 : (h > 4 ? 4 : h);
                        const unsigned m = h > l ? ((0xffffffffu >> (8 * (4 - h))) & (0xffffffffu << (8 * l))) : 0u;
                        o[k] = (o[k] & m) | (fillw[k] & ~m);
                    }
                }
            } else {               // wrap / reflect / clamp at a row edge: element by element
#pragma unroll
                for (int k = 0; k < G; ++k) o[k] = 0u;
#pragma unroll
                for (int t = 0; t < VEC; ++t) {
                    const int col = axis_index(cs + t, a.L, pad);
                    const unsigned char* ep = pb + (rowoff + col) * ES;
                    if (ES == 1) o[t / 4] |= (unsigned)(*ep) << (8 * (t % 4));
                    else if (ES == 2) o[t / 2] |= (unsigned)(*(const unsigned short*)ep) << (16 * (t % 2));
                    else if (ES == 4) o[t] = *(const unsigned*)ep;
                    else { const uint2 v = *(const uint2*)ep; o[(2 * t) % G] = v.x; o[(2 * t + 1) % G] = v.y; }
                }
            }
            stg_words<G>(dst, o);
        }
    }
};

// ================================================================================================
// fp32 arithmetic kernels (active forward, backward).  Items are 4 consecutive elements of a row.
//
// load_seg: NV consecutive source columns cs..cs+NV-1 of row `r` of a staged plane with the
// padding rule applied.  Fast path: the two aligned float4 that cover the window + a
// (warp-uniform) select; zeros padding blends zeros in at the row ends; the other paddings fall
// back to element-wise remapped loads only for windows that cross a row end.
template <int NV>
TS_D void load_seg(const float* __restrict__ plane, int r, int L, int cs, int pad, float* out) {
    const bool interior = cs >= 0 && cs + NV <= L;
    if (r < 0 || (pad == TS_PAD_ZEROS && (cs + NV <= 0 || cs >= L))) {
#pragma unroll
        for (int t = 0; t < NV; ++t) out[t] = 0.f;
        return;
    }
    if (interior || pad == TS_PAD_ZEROS) {
        const int f = r * L + cs;
        const int a0 = f & ~3;
        const float4 A = *(const float4*)(plane + a0);
        const float4 B = *(const float4*)(plane + a0 + 4);
        const float W[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
        const int m = f & 3;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            float v = W[t];
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (m == j) v = W[t + j];
            out[t] = v;
        }
        if (!interior) {
#pragma unroll
            for (int t = 0; t < NV; ++t)
                if (cs + t < 0 || cs + t >= L) out[t] = 0.f;
        }
        return;
    }
#pragma unroll
    for (int t = 0; t < NV; ++t) out[t] = plane[r * L + axis_index(cs + t, L, pad)];
}

// Row index inside a staged plane for (slab a, row b) shifted by the reduced shifts, with the +1
// neighbour selected by `rv` (bit0 = +1 on tensor axis 0, bit1 = +1 on tensor axis 1).
template <int DIM>
TS_D int source_row(int a, int b, const int* s, int rv, int A, int B, int pad) {
    if (DIM == 1) return 0;
    if (DIM == 2) return axis_index(b - s[0] + (rv & 1), B, pad);
    const int ra = axis_index(a - s[0] + (rv & 1), A, pad);
    const int rb = axis_index(b - s[1] + ((rv >> 1) & 1), B, pad);
    return (ra < 0 || rb < 0) ? -1 : ra * B + rb;
}

// Gather the 2^DIM neighbours of the 4 elements of an item from 2^(DIM-1) row segments of 5.
template <int DIM>
TS_D void neighbours_from_rows(const float (*X)[5], int t, float* v) {
    constexpr int NR = 1 << (DIM - 1);
#pragma unroll
    for (int q = 0; q < (1 << DIM); ++q) v[q] = X[q & (NR - 1)][t + (q >> (DIM - 1))];
}

template <int DIM>
struct ActiveFwdBody {
    const SArgs& a;
    const int tid, nt;
    ShiftParams<float, DIM> sp;

    TS_D ActiveFwdBody(const SArgs& a_, int tid_, int nt_) : a(a_), tid(tid_), nt(nt_) {}
    TS_D void begin_unit(long long c) { sp = load_params<float, DIM>((const float*)a.w, c, a.g, true, false); }
    TS_D void end_unit(long long, long long) {}

    TS_D void step(const unsigned char* st, int npl, long long nb, long long c) {
        constexpr int NR = 1 << (DIM - 1);
        const int pad = a.g.pad;
        const int total = npl * a.ipp;
        for (int item = tid; item < total; item += nt) {
            const int pl = (int)fdiv((unsigned)item, a.d_ipp);
            const int rem = item - pl * a.ipp;
            const int orow = (int)fdiv((unsigned)rem, a.d_gpr);
            const int cg = rem - orow * a.gpr;
            int oa = 0, ob = orow;
            if (DIM == 3) { oa = (int)fdiv((unsigned)orow, a.d_rows); ob = orow - oa * a.OB; }
            const float* plane = (const float*)(st + (size_t)pl * a.xpb);
            const int cs = cg * 4 + a.lbL - sp.sx[DIM - 1];
            float X[NR][5];
#pragma unroll
            for (int rv = 0; rv < NR; ++rv)
                load_seg<5>(plane, source_row<DIM>(oa + a.lbA, ob + a.lbB, sp.sx, rv, a.A, a.B, pad), a.L, cs, pad, X[rv]);
            float o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float v[8];
                neighbours_from_rows<DIM>(X, t, v);
                o[t] = interpolate<float, DIM>(v, sp.d);
            }
            float* dst = (float*)a.out + ((nb + pl) * a.g.C + c) * a.g.out_plane + (long long)orow * a.OL + cg * 4;
            __stcs((float4*)dst, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
};

template <int DIM, bool ACTIVE>
struct BackwardBody {
    const SArgs& a;
    const int tid, nt, wid, lane;
    ShiftParams<float, DIM> sp;
    double acc[DIM];

    TS_D BackwardBody(const SArgs& a_, int tid_, int nt_, int wid_, int lane_) : a(a_), tid(tid_), nt(nt_), wid(wid_), lane(lane_) {}
    TS_D void begin_unit(long long c) {
        sp = load_params<float, DIM>((const float*)a.w, c, a.g, ACTIVE, true);
#pragma unroll
        for (int d = 0; d < DIM; ++d) acc[d] = 0.0;
    }
    // one partial per (unit, consumer warp): fixed shuffle tree, no atomics
    TS_D void end_unit(long long c, long long chunk) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double v = acc[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) a.partials[((long long)chunk * a.nw + wid) * (a.g.C * DIM) + c * DIM + d] = v;
        }
    }

    TS_D void step(const unsigned char* st, int npl, long long nb, long long c) {
        constexpr int NR = 1 << (DIM - 1);
        const int pad = a.g.pad;
        const int total = npl * a.ipp;
        for (int item = tid; item < total; item += nt) {
            const int pl = (int)fdiv((unsigned)item, a.d_ipp);
            const int rem = item - pl * a.ipp;
            const int row = (int)fdiv((unsigned)rem, a.d_gpr);      // input-space row (slab*B + b)
            const int cg = rem - row * a.gpr;
            int ia = 0, ib = row;
            if (DIM == 3) { ia = (int)fdiv((unsigned)row, a.d_rows); ib = row - ia * a.B; }
            const int oa = ia - a.lbA, ob = ib - a.lbB, oj0 = cg * 4 - a.lbL;
            float* dst = (float*)a.out + ((nb + pl) * a.g.C + c) * a.g.in_plane + (long long)row * a.L + cg * 4;
            const bool row_ok = oa >= 0 && oa < a.OA && ob >= 0 && ob < a.OB;
            if (!row_ok || oj0 + 4 <= 0 || oj0 >= a.OL) {
                __stcs((float4*)dst, make_float4(0.f, 0.f, 0.f, 0.f));
                continue;
            }
            const float* xpl = (const float*)(st + (size_t)pl * a.xpb);
            const float* gpl = (const float*)(st + (size_t)a.np * a.xpb + (size_t)pl * a.gpb);
            const int grow = oa * a.OB + ob;
            float gv[4];
            load_seg<4>(gpl, grow, a.OL, oj0, TS_PAD_ZEROS, gv);     // zero outside the output window
            // ---- grad_weight terms ----
            float X[NR][5];
            const int cs = cg * 4 - sp.sx[DIM - 1];
#pragma unroll
            for (int rv = 0; rv < NR; ++rv)
                load_seg<5>(xpl, source_row<DIM>(ia, ib, sp.sx, rv, a.A, a.B, pad), a.L, cs, pad, X[rv]);
            float ts[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) ts[d] = 0.f;
            bool okc[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                okc[t] = oj0 + t >= 0 && oj0 + t < a.OL;
                float v[8], wg[3];
                neighbours_from_rows<DIM>(X, t, v);
                weight_partials<float, DIM>(v, sp.d, wg);
#pragma unroll
                for (int d = 0; d < DIM; ++d) ts[d] += okc[t] ? Arith<float>::mul(gv[t], wg[d]) : 0.f;
            }
#pragma unroll
            for (int d = 0; d < DIM; ++d) acc[d] += (double)ts[d];
            // ---- grad_input ----
            float o[4];
            if (ACTIVE) {
                float Gs[NR][5];
                const int gcs = oj0 - sp.sg[DIM - 1];
#pragma unroll
                for (int rv = 0; rv < NR; ++rv)
                    load_seg<5>(gpl, source_row<DIM>(oa, ob, sp.sg, rv, a.OA, a.OB, pad), a.OL, gcs, pad, Gs[rv]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v[8];
                    neighbours_from_rows<DIM>(Gs, t, v);
                    o[t] = okc[t] ? interpolate<float, DIM>(v, sp.d) : 0.f;
                }
            } else {
                int ns[3] = {0, 0, 0};
#pragma unroll
                for (int d = 0; d < DIM; ++d) ns[d] = -sp.sg[d];       // gather at o + shift
                float Gs[4];
                load_seg<4>(gpl, source_row<DIM>(oa, ob, ns, 0, a.OA, a.OB, pad), a.OL, oj0 + sp.sg[DIM - 1], pad, Gs);
#pragma unroll
                for (int t = 0; t < 4; ++t) o[t] = okc[t] ? Gs[t] : 0.f;
            }
            __stcs((float4*)dst, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
};

// ---- kernels -----------------------------------------------------------------------------------
constexpr int MAXT_GATHER = 1024, MAXT_ARITH = 544;

template <int MAXT>
TS_D void setup_barriers(const SArgs& a, unsigned char* smem, uint64_t*& full, uint64_t*& empty) {
    full = (uint64_t*)(smem + (size_t)a.stages * a.stage_stride);
    empty = full + a.stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (unsigned)a.nw); }
        fence_barrier_init();
    }
    __syncthreads();
}

template <int G, int ES>
__global__ void __launch_bounds__(MAXT_GATHER, 1) k_staged_gather(const __grid_constant__ SArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full, *empty;
    setup_barriers<MAXT_GATHER>(a, smem, full, empty);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == a.nw) { if (lane == 0) producer(a, smem, full, empty); return; }
    GatherBody<G, ES> body(a, threadIdx.x, a.nw * 32);
    consumer_loop(a, smem, full, empty, lane, body);
}

template <int DIM>
__global__ void __launch_bounds__(MAXT_ARITH, 1) k_staged_active_forward(const __grid_constant__ SArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full, *empty;
    setup_barriers<MAXT_ARITH>(a, smem, full, empty);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == a.nw) { if (lane == 0) producer(a, smem, full, empty); return; }
    ActiveFwdBody<DIM> body(a, threadIdx.x, a.nw * 32);
    consumer_loop(a, smem, full, empty, lane, body);
}

template <int DIM, bool ACTIVE>
__global__ void __launch_bounds__(MAXT_ARITH, 1) k_staged_backward(const __grid_constant__ SArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full, *empty;
    setup_barriers<MAXT_ARITH>(a, smem, full, empty);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == a.nw) { if (lane == 0) producer(a, smem, full, empty); return; }
    BackwardBody<DIM, ACTIVE> body(a, threadIdx.x, a.nw * 32, wid, lane);
    consumer_loop(a, smem, full, empty, lane, body);
}

template <class K>
int launch(K kernel, const SArgs& a, const StagedPlan& p, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return check_launch();
    kernel<<<p.grid, (p.warps + 1) * 32, p.smem_bytes, s>>>(a);
    note_launch();
    return check_launch();
}

SArgs make_args(const Geo& g, const StagedPlan& p, int mode, int es) {
    SArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.es = es;
    a.xpb = (int)(g.in_plane * es);
    a.gpb = mode == 2 ? (int)(g.out_plane * es) : 0;
    a.np = p.planes_per_step;
    a.stages = p.stages;
    a.nw = p.warps;
    a.stage_stride = (int)((p.smem_bytes - 2 * 8 * p.stages) / p.stages);
    a.n_per_unit = p.n_per_unit;
    a.units = p.units;
    const int d = g.dim;
    a.A = d == 3 ? g.S[0] : 1;        a.OA = d == 3 ? g.OS[0] : 1;      a.lbA = d == 3 ? g.lb[0] : 0;
    a.B = d >= 2 ? g.S[d - 2] : 1;    a.OB = d >= 2 ? g.OS[d - 2] : 1;  a.lbB = d >= 2 ? g.lb[d - 2] : 0;
    a.L = g.S[d - 1];                 a.OL = g.OS[d - 1];               a.lbL = g.lb[d - 1];
    const int vec = p.vec_bytes / es;
    if (mode == 2) {                  // items tile the INPUT plane
        a.gpr = a.L / vec;
        a.ipp = a.A * a.B * a.gpr;
        a.d_rows = make_fastdiv((unsigned)a.B);
    } else {                          // items tile the OUTPUT plane
        a.gpr = a.OL / vec;
        a.ipp = a.OA * a.OB * a.gpr;
        a.d_rows = make_fastdiv((unsigned)a.OB);
    }
    a.d_ipp = make_fastdiv((unsigned)a.ipp);
    a.d_gpr = make_fastdiv((unsigned)a.gpr);
    return a;
}

}  // namespace

// ---- planning -----------------------------------------------------------------------------------
StagedPlan plan_staged(const Geo& g, int mode, int esize, int dtype, bool dense_x, const void* x, const void* y_or_gi,
                       const void* grad, int sm_count) {
    StagedPlan p;
    memset(&p, 0, sizeof(p));
    p.ok = false;
    if (!dense_x || g.N * g.C == 0 || g.in_plane == 0 || g.out_plane == 0) return p;
    if (mode != 0 && dtype != TS_F32) return p;                    // arithmetic kernels: fp32 only (so far)
    if (mode != 0) esize = 4;
    const int d = g.dim;
    const long long Lb = (long long)g.S[d - 1] * esize, OLb = (long long)g.OS[d - 1] * esize;
    int vb = 0;
    if (mode == 0) {
        for (int cand : {16, 8, 4})
            if (cand >= esize && Lb % cand == 0 && OLb % cand == 0) { vb = cand; break; }
    } else if (Lb % 16 == 0 && OLb % 16 == 0) vb = 16;
    if (!vb) return p;
    const long long xpb = g.in_plane * esize, gpb = mode == 2 ? g.out_plane * esize : 0;
    if (xpb % 16 || gpb % 16) return p;
    if (((uintptr_t)x & 15) || ((uintptr_t)grad & 15) || ((uintptr_t)y_or_gi & (uintptr_t)(vb - 1))) return p;
    if ((g.out_plane * esize) % vb) return p;

    const Tuning& t = tuning();
    int stages = t.stages, ctas = t.ctas_per_sm, warps = t.warps;
    const int max_warps = mode == 0 ? 31 : 16;
    if (warps > max_warps) warps = max_warps;
    const long long per_plane = xpb + gpb;
    const long long budget_cta = SMEM_LIMIT / ctas - 1024;
    if (per_plane + 2 * GUARD > budget_cta / 2) {                  // cannot even double-buffer one plane
        if (ctas > 1) { ctas = 1; }
        if (per_plane + 2 * GUARD > (long long)(SMEM_LIMIT - 1024) / 2) return p;
    }
    const long long budget = SMEM_LIMIT / ctas - 1024;
    long long np = ((long long)t.stage_kb * 1024) / per_plane;
    if (np < 1) np = 1;
    if (np > g.N) np = g.N;
    auto stride_of = [&](long long n) { return ((n * per_plane + 2 * GUARD + 127) / 128) * 128; };
    for (;;) {                                                     // shrink until the ring fits
        if (stages * stride_of(np) + 16 * stages <= budget) break;
        if (np > 1) --np;
        else if (stages > 2) --stages;
        else return p;
    }

    const long long planes = g.N * g.C;
    const long long grid_max = (long long)sm_count * ctas;
    long long npu = t.chunk_planes > 0 ? t.chunk_planes : planes / (grid_max * 32);
    npu = (npu / np) * np;
    if (npu < np) npu = np;
    if (npu > g.N) npu = g.N;
    const long long chunks = (g.N + npu - 1) / npu;
    const long long units = chunks * g.C;
    if (units > 0x7fffffffLL || chunks * warps > 0x7fffffffLL) return p;

    p.ok = true;
    p.vec_bytes = vb;
    p.planes_per_step = (int)np;
    p.stages = stages;
    p.warps = warps;
    p.n_per_unit = (int)npu;
    p.units = (int)units;
    p.grid = (int)(units < grid_max ? units : grid_max);
    p.slots = (int)(chunks * warps);
    p.smem_bytes = (size_t)(stages * stride_of(np) + 16 * stages);
    return p;
}

// ---- launchers ----------------------------------------------------------------------------------
int staged_gather(const Geo& g, const StagedPlan& p, int wk, const void* x, void* y, unsigned long long fill, int esize,
                  const void* w, int qkind, long long wzp, cudaStream_t s) {
    SArgs a = make_args(g, p, 0, esize);
    a.x = (const unsigned char*)x;
    a.out = (unsigned char*)y;
    a.w = w;
    a.wk = wk;
    a.qkind = qkind;
    a.wzp = wzp;
    a.fill = fill;
    const int G = p.vec_bytes / 4;
#define TS_GATHER(GG, EE) if (G == GG && esize == EE) return launch(k_staged_gather<GG, EE>, a, p, s);
    TS_GATHER(4, 1) TS_GATHER(2, 1) TS_GATHER(1, 1)
    TS_GATHER(4, 2) TS_GATHER(2, 2) TS_GATHER(1, 2)
    TS_GATHER(4, 4) TS_GATHER(2, 4) TS_GATHER(1, 4)
    TS_GATHER(4, 8) TS_GATHER(2, 8)
#undef TS_GATHER
    return TS_ERR_UNSUPPORTED;
}

int staged_active_forward(const Geo& g, const StagedPlan& p, const void* x, const void* w, void* y, cudaStream_t s) {
    SArgs a = make_args(g, p, 1, 4);
    a.x = (const unsigned char*)x;
    a.out = (unsigned char*)y;
    a.w = w;
    switch (g.dim) {
    case 1: return launch(k_staged_active_forward<1>, a, p, s);
    case 2: return launch(k_staged_active_forward<2>, a, p, s);
    default: return launch(k_staged_active_forward<3>, a, p, s);
    }
}

int staged_backward(const Geo& g, const StagedPlan& p, int active, const void* grad, const void* x, const void* w,
                    void* gi, void* gw, double* partials, cudaStream_t s) {
    SArgs a = make_args(g, p, 2, 4);
    a.x = (const unsigned char*)x;
    a.grad = (const unsigned char*)grad;
    a.out = (unsigned char*)gi;
    a.w = w;
    a.partials = partials;
    int rc;
    switch (g.dim * 2 + (active ? 1 : 0)) {
    case 2: rc = launch(k_staged_backward<1, false>, a, p, s); break;
    case 3: rc = launch(k_staged_backward<1, true>, a, p, s); break;
    case 4: rc = launch(k_staged_backward<2, false>, a, p, s); break;
    case 5: rc = launch(k_staged_backward<2, true>, a, p, s); break;
    case 6: rc = launch(k_staged_backward<3, false>, a, p, s); break;
    default: rc = launch(k_staged_backward<3, true>, a, p, s); break;
    }
    if (rc != TS_OK) return rc;
    return launch_reduce_partials<float>(partials, p.slots, (int)(g.C * g.dim), gw, s);
}

}  // namespace ts
