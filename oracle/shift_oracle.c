/*
 * shift_oracle.c -- CPU ORACLE for the Sparse/Active Shift operator.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference algorithm
 * (DeadAt0m/ActiveSparseShifts-PyTorch, torchshifts/csrc).  It exists so that the CUDA product
 * path can be checked against the reference's semantics on any box.  It must never be imported,
 * linked or called from the product path (activesparseshifts-pytorch_b200/): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the known-answer vectors of SURVEY.md 8(c),
 *   (2) golden fixtures under tests/golden/ produced by the reference's own torch extension
 *       (oracle/build_ref_full.sh + tests/golden/make_golden.py), and
 *   (3) oracle/_ref/libref_shifts.so, which compiles the reference's own per-element headers.
 *
 * Reference lines followed (paths relative to torchshifts/csrc/ops/):
 *   mod                      kernels/shifts_kernels.h:7-8
 *   remap (infer_index)      kernels/shifts_kernels.h:10-29
 *   fetch                    kernels/shifts_kernels.h:32-54    (get_shifted_value)
 *   fetch_nb                 kernels/shifts_kernels.h:58-103   (get_shifted_values)
 *   lerp / interpolate       kernels/interpolation.h:3-38, shifts_kernels.h:110-130
 *   weight_partials          kernels/interpolation.h:9-61, shifts_kernels.h:132-154
 *   forward body             kernels/shifts_kernels.h:156-220, cpu/shifts_cpu.cpp:78-98
 *   backward body            kernels/shifts_kernels.h:222-327, cpu/shifts_cpu.cpp:184-208
 *   weight split             cpu/shifts_cpu.cpp:223-224 (fwd), :242-244 (bwd)
 *   quantized forward        kernels/shifts_kernels.h:532-571, quantized/shifts_quantized.cpp:107-130
 *
 * Compile:  gcc -O2 -ffp-contract=off -fPIC -shared -o liboracle_shifts.so shift_oracle.c -lm
 * (-ffp-contract=off: the reference CPU build has no FMA contraction)
 *
 * Conventions: tensors are [N, C, S0, S1, S2] with unused spatial sizes = 1.  `xs` are the five
 * input strides in elements; outputs are dense.  lb/rb are the (already validated) borders in
 * input coordinates, output size OS[a] = rb[a]-lb[a].  Padding modes: 0 zeros, 1 border,
 * 2 periodic, 3 reflect, 4 symmetric.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

static i64 pmod(i64 a, i64 b) { return (b + (a % b)) % b; }

/* index remap of one axis; a negative result means "outside, use the pad value" */
static i64 remap(i64 idx, i64 len, int pad)
{
    i64 neg, odd;
    switch (pad) {
    case 1:  return idx < 0 ? 0 : (idx > len - 1 ? len - 1 : idx);
    case 2:  return pmod(idx, len);
    case 3:
        neg = idx < 0;
        odd = (neg + (llabs(idx) - neg) / (len - 1)) & 1;
        return odd ? (len - 1 - pmod(idx, len - 1)) : pmod(idx, len - 1);
    case 4:
        neg = idx < 0;
        odd = (neg + (llabs(idx) - neg) / len) & 1;
        return odd ? (len - 1 - pmod(idx, len)) : pmod(idx, len);
    default: return idx > len - 1 ? -1 : idx;
    }
}

/* element offset of the remapped position inside one (n,c) plane, or -1 for "pad value".
 * A size-1 axis ignores its shift (always index 0). */
static i64 plane_offset(int dim, const i64 idx[3], const i64 sizes[3], const i64 str[3], int pad)
{
    i64 off = 0;
    for (int a = 0; a < dim; ++a) {
        i64 t = sizes[a] == 1 ? 0 : remap(idx[a], sizes[a], pad);
        if (t < 0) return -1;
        off += t * str[a];
    }
    return off;
}

void oracle_remap_axis(int pad, i64 len, i64 n, const i64 *idx, i64 *out)
{
    for (i64 t = 0; t < n; ++t) out[t] = len == 1 ? 0 : remap(idx[t], len, pad);
}

/* ------------------------------------------------------------------------------------------ */
#define DEFINE_FLOAT_ORACLE(T, SFX, FLOORF, CEILF, RINTF)                                        \
                                                                                                  \
static T lerp_##SFX(T a, T b, T x) { return a * ((T)1 - x) + b * x; }                           \
                                                                                                  \
static T fetch_##SFX(const T *plane, int dim, const i64 idx[3], const i64 sizes[3],              \
                     const i64 str[3], int pad, int ok)                                          \
{                                                                                                 \
    i64 off = plane_offset(dim, idx, sizes, str, pad);                                            \
    return (ok && off >= 0) ? plane[off] : (T)0;                                                  \
}                                                                                                 \
                                                                                                  \
/* neighbour q: +1 along axis 0 if q&1, axis 1 if q&2, axis 2 if q&4 */                           \
static void fetch_nb_##SFX(const T *plane, int dim, const i64 idx[3], const i64 sizes[3],        \
                           const i64 str[3], int pad, int ok, T v[8])                            \
{                                                                                                 \
    int nq = 1 << dim;                                                                            \
    for (int q = 0; q < nq; ++q) {                                                                \
        i64 t[3] = { idx[0] + (q & 1), idx[1] + ((q >> 1) & 1), idx[2] + ((q >> 2) & 1) };        \
        v[q] = fetch_##SFX(plane, dim, t, sizes, str, pad, ok);                                   \
    }                                                                                             \
}                                                                                                 \
                                                                                                  \
static T interp2_##SFX(const T *v, T x, T y)                                                      \
{ return lerp_##SFX(lerp_##SFX(v[0], v[1], x), lerp_##SFX(v[2], v[3], x), y); }                   \
                                                                                                  \
static T interpolate_##SFX(const T v[8], int dim, const T d[3])                                   \
{                                                                                                 \
    if (dim == 1) return lerp_##SFX(v[0], v[1], d[0]);                                            \
    if (dim == 2) return interp2_##SFX(v, d[0], d[1]);                                            \
    return lerp_##SFX(interp2_##SFX(v, d[0], d[1]), interp2_##SFX(v + 4, d[0], d[1]), d[2]);      \
}                                                                                                 \
                                                                                                  \
/* the reference's "dx/dy/dz" partials (note: in 2D/3D g[0] is a difference along axis 1) */      \
static T dx2_##SFX(const T *v, T y) { return lerp_##SFX(v[2] - v[0], v[3] - v[1], y); }           \
static T dy2_##SFX(const T *v, T x)                                                               \
{ return lerp_##SFX(v[2], v[3], x) - lerp_##SFX(v[0], v[1], x); }                                 \
                                                                                                  \
static void weight_partials_##SFX(const T v[8], int dim, const T d[3], T g[3])                    \
{                                                                                                 \
    if (dim == 1) { g[0] = v[1] - v[0]; return; }                                                 \
    if (dim == 2) { g[0] = dx2_##SFX(v, d[1]); g[1] = dy2_##SFX(v, d[0]); return; }               \
    g[0] = lerp_##SFX(dx2_##SFX(v, d[1]), dx2_##SFX(v + 4, d[1]), d[2]);                          \
    g[1] = lerp_##SFX(dy2_##SFX(v, d[0]), dy2_##SFX(v + 4, d[0]), d[2]);                          \
    g[2] = interp2_##SFX(v + 4, d[0], d[1]) - interp2_##SFX(v, d[0], d[1]);                       \
}                                                                                                 \
                                                                                                  \
/* integer / fractional parts of the shifts, forward flavour */                                    \
void oracle_split_forward_##SFX(int active, i64 count, const T *w, i64 *iw, T *dw)                \
{                                                                                                 \
    for (i64 t = 0; t < count; ++t) {                                                             \
        iw[t] = (i64)(active ? FLOORF(w[t]) : RINTF(w[t]));                                       \
        dw[t] = active ? (w[t] - (T)iw[t]) : (T)0;                                                \
    }                                                                                             \
}                                                                                                 \
                                                                                                  \
/* backward flavour (note the SSL fraction and the truncating cast of w - dw) */                  \
void oracle_split_backward_##SFX(int active, i64 count, const T *w, i64 *iw, T *dw)               \
{                                                                                                 \
    for (i64 t = 0; t < count; ++t) {                                                             \
        if (active) { dw[t] = w[t] - FLOORF(w[t]); iw[t] = (i64)(w[t] - dw[t]); }                 \
        else { dw[t] = w[t] > 0 ? (w[t] - FLOORF(w[t])) : (CEILF(w[t]) - w[t]);                   \
               iw[t] = (i64)RINTF(w[t]); }                                                        \
    }                                                                                             \
}                                                                                                 \
                                                                                                  \
void oracle_shift_forward_##SFX(int dim, int pad, int active, const T *x, const i64 xs[5],       \
                                T *y, const T *w, i64 N, i64 C, const i64 S[3],                  \
                                const i64 lb[3], const i64 rb[3])                                \
{                                                                                                 \
    i64 OS[3] = { rb[0] - lb[0], rb[1] - lb[1], rb[2] - lb[2] };                                  \
    i64 *iw = (i64 *)malloc(sizeof(i64) * (size_t)(C * dim));                                     \
    T *dw = (T *)malloc(sizeof(T) * (size_t)(C * dim));                                           \
    oracle_split_forward_##SFX(active, C * dim, w, iw, dw);                                       \
    for (i64 n = 0; n < N; ++n)                                                                   \
    for (i64 c = 0; c < C; ++c) {                                                                 \
        const T *plane = x + n * xs[0] + c * xs[1];                                               \
        T *out = y + (n * C + c) * OS[0] * OS[1] * OS[2];                                         \
        i64 sh[3] = { 0, 0, 0 }; T d[3] = { 0, 0, 0 };                                            \
        for (int a = 0; a < dim; ++a) { sh[a] = iw[c * dim + a]; d[a] = dw[c * dim + a]; }        \
        for (i64 i = lb[0]; i < rb[0]; ++i)                                                       \
        for (i64 j = lb[1]; j < rb[1]; ++j)                                                       \
        for (i64 k = lb[2]; k < rb[2]; ++k) {                                                     \
            i64 s[3] = { i - sh[0], j - sh[1], k - sh[2] };                                       \
            i64 o = ((i - lb[0]) * OS[1] + (j - lb[1])) * OS[2] + (k - lb[2]);                    \
            if (active) {                                                                         \
                T v[8]; fetch_nb_##SFX(plane, dim, s, S, xs + 2, pad, 1, v);                      \
                out[o] = interpolate_##SFX(v, dim, d);                                            \
            } else out[o] = fetch_##SFX(plane, dim, s, S, xs + 2, pad, 1);                        \
        }                                                                                         \
    }                                                                                             \
    free(iw); free(dw);                                                                           \
}                                                                                                 \
                                                                                                  \
/* grad is dense [N,C,OS]; gi dense [N,C,S]; gw [C,dim] is accumulated serially in T, in the    */\
/* (n, c, i, j, k) order of the single-threaded reference loop.                                 */\
void oracle_shift_backward_##SFX(int dim, int pad, int active, const T *grad, const T *x,        \
                                 const i64 xs[5], T *gi, T *gw, const T *w, i64 N, i64 C,        \
                                 const i64 S[3], const i64 lb[3], const i64 rb[3])               \
{                                                                                                 \
    i64 OS[3] = { rb[0] - lb[0], rb[1] - lb[1], rb[2] - lb[2] };                                  \
    i64 gs[3] = { OS[1] * OS[2], OS[2], 1 };                                                      \
    i64 *iw = (i64 *)malloc(sizeof(i64) * (size_t)(C * dim));                                     \
    T *dw = (T *)malloc(sizeof(T) * (size_t)(C * dim));                                           \
    oracle_split_backward_##SFX(active, C * dim, w, iw, dw);                                      \
    for (i64 t = 0; t < C * dim; ++t) gw[t] = (T)0;                                               \
    for (i64 n = 0; n < N; ++n)                                                                   \
    for (i64 c = 0; c < C; ++c) {                                                                 \
        const T *xp = x + n * xs[0] + c * xs[1];                                                  \
        const T *gp = grad + (n * C + c) * OS[0] * OS[1] * OS[2];                                 \
        T *gip = gi + (n * C + c) * S[0] * S[1] * S[2];                                           \
        i64 sh[3] = { 0, 0, 0 }; T d[3] = { 0, 0, 0 };                                            \
        for (int a = 0; a < dim; ++a) { sh[a] = iw[c * dim + a]; d[a] = dw[c * dim + a]; }        \
        for (i64 i = 0; i < S[0]; ++i)                                                            \
        for (i64 j = 0; j < S[1]; ++j)                                                            \
        for (i64 k = 0; k < S[2]; ++k) {                                                          \
            int ok = i >= lb[0] && i < rb[0] && j >= lb[1] && j < rb[1] && k >= lb[2] && k < rb[2];\
            i64 o[3] = { i - lb[0], j - lb[1], k - lb[2] };                                       \
            T g = ok ? gp[o[0] * gs[0] + o[1] * gs[1] + o[2]] : (T)0;                             \
            i64 s[3] = { i - sh[0], j - sh[1], k - sh[2] };                                       \
            T v[8], wg[3] = { 0, 0, 0 };                                                          \
            fetch_nb_##SFX(xp, dim, s, S, xs + 2, pad, ok, v);                                    \
            if (ok) weight_partials_##SFX(v, dim, d, wg);                                         \
            for (int a = 0; a < dim; ++a) gw[c * dim + a] += g * wg[a];                           \
            T r;                                                                                  \
            if (active) {                                                                         \
                i64 q[3] = { o[0] - sh[0], o[1] - sh[1], o[2] - sh[2] };                          \
                fetch_nb_##SFX(gp, dim, q, OS, gs, pad, ok, v);                                   \
                r = ok ? interpolate_##SFX(v, dim, d) : (T)0;                                     \
            } else {                                                                              \
                i64 q[3] = { o[0] + sh[0], o[1] + sh[1], o[2] + sh[2] };                          \
                r = fetch_##SFX(gp, dim, q, OS, gs, pad, ok);                                     \
            }                                                                                     \
            gip[(i * S[1] + j) * S[2] + k] = r;                                                   \
        }                                                                                         \
    }                                                                                             \
    free(iw); free(dw);                                                                           \
}

DEFINE_FLOAT_ORACLE(float, f32, floorf, ceilf, rintf)
DEFINE_FLOAT_ORACLE(double, f64, floor, ceil, rint)

/* ------------------------------------------------------------------------------------------ */
/* Quantized forward: integer copy of the raw representation.  `wq` are the raw integer weights  */
/* (int_repr), `wzp` their zero point, `zp` the input zero point (pad value), esize 1 or 4.      */
void oracle_qshift_forward(int dim, int pad, int esize, const void *x, const i64 xs[5], void *y,
                           const i64 *wq, i64 wzp, i64 zp, i64 N, i64 C, const i64 S[3],
                           const i64 lb[3], const i64 rb[3])
{
    i64 OS[3] = { rb[0] - lb[0], rb[1] - lb[1], rb[2] - lb[2] };
    const unsigned char *xb = (const unsigned char *)x;
    unsigned char *yb = (unsigned char *)y;
    unsigned char fill[4];
    if (esize == 1) { fill[0] = (unsigned char)(zp & 0xff); }
    else { int32_t z = (int32_t)zp; memcpy(fill, &z, 4); }
    for (i64 n = 0; n < N; ++n)
    for (i64 c = 0; c < C; ++c) {
        i64 base = n * xs[0] + c * xs[1];
        i64 obase = (n * C + c) * OS[0] * OS[1] * OS[2];
        i64 sh[3] = { 0, 0, 0 };
        for (int a = 0; a < dim; ++a) sh[a] = wq[c * dim + a] - wzp;
        for (i64 i = lb[0]; i < rb[0]; ++i)
        for (i64 j = lb[1]; j < rb[1]; ++j)
        for (i64 k = lb[2]; k < rb[2]; ++k) {
            i64 s[3] = { i - sh[0], j - sh[1], k - sh[2] };
            i64 o = obase + ((i - lb[0]) * OS[1] + (j - lb[1])) * OS[2] + (k - lb[2]);
            i64 off = plane_offset(dim, s, S, xs + 2, pad);
            if (off >= 0) memcpy(yb + o * esize, xb + (base + off) * esize, (size_t)esize);
            else memcpy(yb + o * esize, fill, (size_t)esize);
        }
    }
}

/* check_borders restatement (csrc/ops/shifts.cpp:93-135).  `user` is [dim,2] (left_cut,
 * right_cut) or NULL.  Writes lb/rb (unused axes: 0/1).  Returns 0, or -1 when the reference
 * would attempt to allocate a negative dimension. */
int oracle_check_borders(int dim, const i64 S[3], const i64 *user, i64 lb[3], i64 rb[3])
{
    for (int a = 0; a < 3; ++a) { lb[a] = 0; rb[a] = a < dim ? S[a] : 1; }
    if (user) {
        for (int a = 0; a < dim; ++a) {
            int l, r, size = (int)S[a];
            r = (int)rb[a] - (int)user[2 * a + 1];
            l = (int)user[2 * a];
            if (r - l < 1) r = l + 1;
            if (l == size) { l = size - 1; r = l + 1; }
            if (r == 0) { l = 0; r = 1; }
            if (l < 0) l = 0;
            if (r > size) r = size;
            lb[a] = l; rb[a] = r;
            if (r - l < 0) return -1;
        }
    }
    return 0;
}
