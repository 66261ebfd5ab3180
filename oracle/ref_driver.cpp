// ref_driver.cpp -- C-ABI loop nest around the REFERENCE's own per-element kernels.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into the product path.
//
// The arithmetic here is not ours: this translation unit #includes the reference header
//   /root/reference/torchshifts/csrc/ops/kernels/shifts_kernels.h   (and, through it,
//   kernels/interpolation.h and global_scope.h)
// from where it lies (oracle/Makefile passes -I$(REFERENCE)/torchshifts/csrc/ops) and calls
//   shift_forward_kernel_nchwd    (shifts_kernels.h:156-220)
//   shift_backward_kernel_nchwd   (shifts_kernels.h:222-327)
//   shift_forward_kernel_nchwd_q  (shifts_kernels.h:532-571)
// once per element, exactly as the reference's torch-dependent drivers do
// (cpu/shifts_cpu.cpp:78-98, :184-208; quantized/shifts_quantized.cpp:83-100).  Only the loop
// nest (at::parallel_for there, OpenMP here) and the C ABI are written here, so the library is
// torch-free, loads next to the product in one process, and builds in seconds.
//
// One deliberate difference: the reference accumulates grad_weight with a plain `+=` from all
// at::parallel_for threads (a data race).  Here every OpenMP thread accumulates
// into a private copy and the copies are summed in thread order, so multi-threaded timing runs
// are race-free; with threads == 1 the accumulation order is the reference's serial order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SHIFTS_CPU
#include "kernels/shifts_kernels.h"

namespace {

struct Geo {
    int64_t N, C, S[3], xs[5], lb[3], rb[3], OS[3], os[5], gs[5];
};

Geo make_geo(int64_t N, int64_t C, const int64_t* S, const int64_t* xs, const int64_t* lb, const int64_t* rb) {
    Geo g;
    g.N = N; g.C = C;
    for (int a = 0; a < 3; ++a) { g.S[a] = S[a]; g.lb[a] = lb[a]; g.rb[a] = rb[a]; g.OS[a] = rb[a] - lb[a]; }
    for (int a = 0; a < 5; ++a) g.xs[a] = xs[a];
    g.os[4] = 1; g.os[3] = g.OS[2]; g.os[2] = g.OS[1] * g.OS[2]; g.os[1] = g.OS[0] * g.os[2]; g.os[0] = C * g.os[1];
    g.gs[4] = 1; g.gs[3] = g.S[2]; g.gs[2] = g.S[1] * g.S[2]; g.gs[1] = g.S[0] * g.gs[2]; g.gs[0] = C * g.gs[1];
    return g;
}

template <typename T, int D, BIPadding P, bool A>
void fwd(const Geo& g, const T* x, T* y, const int64_t* iw, const T* dw, int threads) {
    const int64_t planes = g.N * g.C;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t index = 0; index < planes; ++index) {
        const int64_t c = index % g.C, n = index / g.C;
        for (int64_t i = 0; i < g.S[0]; ++i)
            for (int64_t j = 0; j < g.S[1]; ++j)
                for (int64_t k = 0; k < g.S[2]; ++k)
                    shift_forward_kernel_nchwd<T, int64_t, D, P, A>(
                        x, y, iw, dw, n, c, i, j, k, g.S[0], g.S[1], g.S[2],
                        g.xs[0], g.xs[1], g.xs[2], D < 2 ? 0 : g.xs[3], D < 3 ? 0 : g.xs[4],
                        g.os[0], g.os[1], g.os[2], D < 2 ? 0 : g.os[3], D < 3 ? 0 : g.os[4],
                        (int64_t)D, (int64_t)1, (int64_t)D, (int64_t)1,
                        g.lb[0], g.lb[1], g.lb[2], g.rb[0], g.rb[1], g.rb[2]);
    }
}

template <typename T, int D, BIPadding P, bool A>
void bwd(const Geo& g, const T* grad, const T* x, T* gi, T* gw, const int64_t* iw, const T* dw, int threads) {
    const int64_t planes = g.N * g.C;
    const int64_t wn = g.C * D;
    if (threads < 1) threads = 1;
    std::vector<T> priv((size_t)threads * (size_t)wn, (T)0);
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        T* mygw = priv.data() + (size_t)tid * (size_t)wn;
#pragma omp for schedule(static)
        for (int64_t index = 0; index < planes; ++index) {
            const int64_t c = index % g.C, n = index / g.C;
            for (int64_t i = 0; i < g.S[0]; ++i)
                for (int64_t j = 0; j < g.S[1]; ++j)
                    for (int64_t k = 0; k < g.S[2]; ++k)
                        shift_backward_kernel_nchwd<T, int64_t, D, P, A>(
                            grad, x, gi, iw, dw, mygw, n, c, i, j, k, g.C, g.S[0], g.S[1], g.S[2],
                            g.os[0], g.os[1], g.os[2], D < 2 ? 0 : g.os[3], D < 3 ? 0 : g.os[4],
                            g.xs[0], g.xs[1], g.xs[2], D < 2 ? 0 : g.xs[3], D < 3 ? 0 : g.xs[4],
                            g.gs[0], g.gs[1], g.gs[2], D < 2 ? 0 : g.gs[3], D < 3 ? 0 : g.gs[4],
                            (int64_t)D, (int64_t)1, (int64_t)D, (int64_t)1, (int64_t)D, (int64_t)1,
                            g.lb[0], g.lb[1], g.lb[2], g.rb[0], g.rb[1], g.rb[2]);
        }
    }
    for (int64_t t = 0; t < wn; ++t) {
        T acc = priv[(size_t)t];
        for (int th = 1; th < threads; ++th) acc += priv[(size_t)th * (size_t)wn + (size_t)t];
        gw[t] = acc;
    }
}

template <typename T, int D, BIPadding P>
void qfwd(const Geo& g, const T* x, T* y, const int64_t* wq, int64_t wzp, T zp, int threads) {
    const int64_t planes = g.N * g.C;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t index = 0; index < planes; ++index) {
        const int64_t c = index % g.C, n = index / g.C;
        for (int64_t i = 0; i < g.S[0]; ++i)
            for (int64_t j = 0; j < g.S[1]; ++j)
                for (int64_t k = 0; k < g.S[2]; ++k)
                    shift_forward_kernel_nchwd_q<T, int64_t, D, P>(
                        x, y, wq, n, c, i, j, k, g.S[0], g.S[1], g.S[2],
                        g.xs[0], g.xs[1], g.xs[2], D < 2 ? 0 : g.xs[3], D < 3 ? 0 : g.xs[4],
                        g.os[0], g.os[1], g.os[2], D < 2 ? 0 : g.os[3], D < 3 ? 0 : g.os[4],
                        (int64_t)D, (int64_t)1,
                        g.lb[0], g.lb[1], g.lb[2], g.rb[0], g.rb[1], g.rb[2], zp, wzp);
    }
}

#define PAD_SWITCH(CALL)                                     \
    switch (pad) {                                           \
    case 0: CALL(BIPadding::Zeros); break;                   \
    case 1: CALL(BIPadding::Border); break;                  \
    case 2: CALL(BIPadding::Periodic); break;                \
    case 3: CALL(BIPadding::Reflect); break;                 \
    case 4: CALL(BIPadding::Symmetric); break;               \
    default: return -1;                                      \
    }

template <typename T>
int fwd_dispatch(int dim, int pad, int active, const Geo& g, const T* x, T* y, const int64_t* iw, const T* dw, int th) {
#define F(P) do { if (dim == 1) { if (active) fwd<T,1,P,true>(g,x,y,iw,dw,th); else fwd<T,1,P,false>(g,x,y,iw,dw,th); } \
             else if (dim == 2) { if (active) fwd<T,2,P,true>(g,x,y,iw,dw,th); else fwd<T,2,P,false>(g,x,y,iw,dw,th); } \
             else if (dim == 3) { if (active) fwd<T,3,P,true>(g,x,y,iw,dw,th); else fwd<T,3,P,false>(g,x,y,iw,dw,th); } \
             else return -1; } while (0)
    PAD_SWITCH(F)
#undef F
    return 0;
}

template <typename T>
int bwd_dispatch(int dim, int pad, int active, const Geo& g, const T* grad, const T* x, T* gi, T* gw,
                 const int64_t* iw, const T* dw, int th) {
#define B(P) do { if (dim == 1) { if (active) bwd<T,1,P,true>(g,grad,x,gi,gw,iw,dw,th); else bwd<T,1,P,false>(g,grad,x,gi,gw,iw,dw,th); } \
             else if (dim == 2) { if (active) bwd<T,2,P,true>(g,grad,x,gi,gw,iw,dw,th); else bwd<T,2,P,false>(g,grad,x,gi,gw,iw,dw,th); } \
             else if (dim == 3) { if (active) bwd<T,3,P,true>(g,grad,x,gi,gw,iw,dw,th); else bwd<T,3,P,false>(g,grad,x,gi,gw,iw,dw,th); } \
             else return -1; } while (0)
    PAD_SWITCH(B)
#undef B
    return 0;
}

template <typename T>
int q_dispatch(int dim, int pad, const Geo& g, const T* x, T* y, const int64_t* wq, int64_t wzp, T zp, int th) {
#define Q(P) do { if (dim == 1) qfwd<T,1,P>(g,x,y,wq,wzp,zp,th); else if (dim == 2) qfwd<T,2,P>(g,x,y,wq,wzp,zp,th); \
             else if (dim == 3) qfwd<T,3,P>(g,x,y,wq,wzp,zp,th); else return -1; } while (0)
    PAD_SWITCH(Q)
#undef Q
    return 0;
}

}  // namespace

extern "C" {

int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// iw / dw are the already split integer and fractional shifts, dense [C, dim] (weights_sC = dim,
// weights_sS = 1), produced by the caller exactly as cpu/shifts_cpu.cpp:223-224 / :242-244 do.
int ref_shift_forward_f32(int dim, int pad, int active, const float* x, const int64_t* xs, float* y,
                          const int64_t* iw, const float* dw, int64_t N, int64_t C, const int64_t* S,
                          const int64_t* lb, const int64_t* rb, int threads) {
    return fwd_dispatch<float>(dim, pad, active, make_geo(N, C, S, xs, lb, rb), x, y, iw, dw, threads);
}
int ref_shift_forward_f64(int dim, int pad, int active, const double* x, const int64_t* xs, double* y,
                          const int64_t* iw, const double* dw, int64_t N, int64_t C, const int64_t* S,
                          const int64_t* lb, const int64_t* rb, int threads) {
    return fwd_dispatch<double>(dim, pad, active, make_geo(N, C, S, xs, lb, rb), x, y, iw, dw, threads);
}
int ref_shift_backward_f32(int dim, int pad, int active, const float* grad, const float* x, const int64_t* xs,
                           float* gi, float* gw, const int64_t* iw, const float* dw, int64_t N, int64_t C,
                           const int64_t* S, const int64_t* lb, const int64_t* rb, int threads) {
    return bwd_dispatch<float>(dim, pad, active, make_geo(N, C, S, xs, lb, rb), grad, x, gi, gw, iw, dw, threads);
}
int ref_shift_backward_f64(int dim, int pad, int active, const double* grad, const double* x, const int64_t* xs,
                           double* gi, double* gw, const int64_t* iw, const double* dw, int64_t N, int64_t C,
                           const int64_t* S, const int64_t* lb, const int64_t* rb, int threads) {
    return bwd_dispatch<double>(dim, pad, active, make_geo(N, C, S, xs, lb, rb), grad, x, gi, gw, iw, dw, threads);
}
// raw integer representation, esize 1 (qint8 / quint8 storage) or 4 (qint32 storage)
int ref_qshift_forward(int dim, int pad, int esize, const void* x, const int64_t* xs, void* y,
                       const int64_t* wq, int64_t wzp, int64_t zp, int64_t N, int64_t C, const int64_t* S,
                       const int64_t* lb, const int64_t* rb, int threads) {
    Geo g = make_geo(N, C, S, xs, lb, rb);
    if (esize == 1)
        return q_dispatch<uint8_t>(dim, pad, g, (const uint8_t*)x, (uint8_t*)y, wq, wzp, (uint8_t)(zp & 0xff), threads);
    if (esize == 4)
        return q_dispatch<int32_t>(dim, pad, g, (const int32_t*)x, (int32_t*)y, wq, wzp, (int32_t)zp, threads);
    return -1;
}

}  // extern "C"
