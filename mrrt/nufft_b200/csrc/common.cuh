// Shared device helpers for libb200nufft (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2n {

constexpr int kMaxJ = 16;   // largest supported kernel width per axis
constexpr int kMaxDim = 3;

template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };
template <typename T> using cplx_t = typename Cplx<T>::type;

template <typename T> __host__ __device__ inline cplx_t<T> make_c(T re, T im);
template <> __host__ __device__ inline float2 make_c<float>(float re, float im) { return make_float2(re, im); }
template <> __host__ __device__ inline double2 make_c<double>(double re, double im) { return make_double2(re, im); }

// IEEE-correct division regardless of compile flags (tm must be bit-exact)
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ bool is_finite(float a) { return isfinite(a); }
__device__ __forceinline__ bool is_finite(double a) { return isfinite(a); }

// window origin: koff = 1 + floor(t - J/2.) evaluated in double
// (c/nufft_table.template.c:865-867; the promotion of a float t is exact)
template <typename T> __device__ __forceinline__ int window_origin(T t, int J) {
    return 1 + (int)floor((double)t - (double)J * 0.5);
}

__device__ __forceinline__ int wrap_index(int k, int K) {
    int r = k % K;
    return r < 0 ? r + K : r;
}

// grid index held by THIS plan of global index k: periodic wrap on the global grid, then the
// offset of the plan's slab (Kg == K and korg == 0 for an ordinary plan: plain wrap)
__device__ __forceinline__ int local_index(int k, int Kg, int korg) {
    int r = wrap_index(k, Kg) - korg;
    return r < 0 ? r + Kg : r;
}

// Device-side description of the transform geometry (passed by value to kernels)
struct Geom {
    int ndim;
    int N[3];
    int K[3];         // oversampled grid held by this plan (a slab plan: the LOCAL extent)
    int Kg[3];        // global oversampled grid size: period of the coordinates (= K unless slab)
    int korg[3];      // slab plans: global index of local row 0 along each axis (else 0)
    int J[3];
    int L;
    int order;        // table lookup: 1 linear interpolation between entries (the reference's only
                      // CPU mode), 0 the entry at floor(p) ("order 0" of cuda/jinja/table_*.jinja)
    int ncenter[3];   // floor(J*L/2): centre of each table
    int tlen[3];      // J*L+1
    int tile[3];      // bin shape in grid cells
    int nbin[3];      // bins per axis
    int tile_b[3];    // bin shape of the adjoint sort order (long along the last axis)
    int nbin_b[3];
    int colmode;      // adjoint sort order = COLUMN order (spread_column.cuh): bins of tile_b[0] x tile_b[1]
                      // grid columns over the whole last axis, samples ordered by the origin along
                      // the last axis inside a bin (else: cells last-axis-fastest inside the bin)
    int64_t PK;       // prod(K)
    int64_t PN;       // prod(N)
    int64_t M;        // samples
    int cells_per_tile;
};

// One tap coefficient by linear interpolation of the centred table
// (template.c:870-873): p=(t-k)*L in T, n=floor(p), alf=p-n,
// coef=(1-alf)*h[n]+alf*h[n+1].  The n+1 read is clamped to the last entry: it can
// only exceed the table when alf==0 (the reference reads one past the end there).
// The n read can fall one entry BEFORE the table when `t - J/2.` rounds to an integer in
// the window-origin formula while t itself is a hair below it (e.g. K=32, J=6,
// omega=-13*2pi/32 in float64: t=-13.000000000000002, koff=-15, last tap at
// p=-3072.000000000002): the reference reads h[-1] there (undefined behaviour, weight
// 1-alf ~ 1e-12); that entry is taken as 0 here.
template <typename T>
__device__ __forceinline__ T tap_real(const T* __restrict__ h, int ncenter, int tlen, T t, int k, int L,
                                      int order = 1) {
    const T p = (t - (T)k) * (T)L;
    const T fl = floor(p);
    const int n = (int)fl;
    const T alf = p - fl;
    const int i0 = ncenter + n;
    const int i1 = max(min(i0 + 1, tlen - 1), 0);
    const T h0 = i0 >= 0 ? h[i0] : (T)0;
    if (order == 0) return h0;                       // table_2d_forward.jinja:61-67
    return ((T)1 - alf) * h0 + alf * h[i1];
}

template <typename T>
__device__ __forceinline__ cplx_t<T> tap_cplx(const cplx_t<T>* __restrict__ h, int ncenter, int tlen, T t, int k, int L,
                                              int order = 1) {
    const T p = (t - (T)k) * (T)L;
    const T fl = floor(p);
    const int n = (int)fl;
    const T alf = p - fl;
    const int i0 = ncenter + n;
    const int i1 = max(min(i0 + 1, tlen - 1), 0);
    const cplx_t<T> a = i0 >= 0 ? h[i0] : make_c<T>(0, 0), b = h[i1];
    if (order == 0) return a;
    return make_c<T>(((T)1 - alf) * a.x + alf * b.x, ((T)1 - alf) * a.y + alf * b.y);
}

__device__ __forceinline__ void atomic_add_c(float2* p, float2 v) {
    atomicAdd(p, v);   // REDG.E.ADD.F32x2 on sm_90+
}
__device__ __forceinline__ void atomic_add_c(double2* p, double2 v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename C> __device__ __forceinline__ C cmul_conj(C a, C b) {   // a * conj(b)
    C r;
    r.x = a.x * b.x + a.y * b.y;
    r.y = a.y * b.x - a.x * b.y;
    return r;
}

// Packed FP32x2 arithmetic (Blackwell FFMA2 / FMUL2, PTX fma.rn.f32x2): a complex value
// times a real weight is one instruction instead of two.  Each half is an IEEE fma, so the
// result is bit-identical to two scalar fmaf calls.
__device__ __forceinline__ float2 fma_w(float w, float2 v, float2 acc) {   // acc + w * v
    float2 ww = make_float2(w, w);
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&ww);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&v);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&acc);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul_w(float w, float2 v) {                // w * v
    float2 ww = make_float2(w, w);
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&ww);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&v);
    unsigned long long rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ double2 fma_w(double w, double2 v, double2 acc) {
    return make_double2(fma(w, v.x, acc.x), fma(w, v.y, acc.y));
}
__device__ __forceinline__ double2 mul_w(double w, double2 v) { return make_double2(w * v.x, w * v.y); }

// weight algebra: W is T (real table) or cplx_t<T> (complex table)
template <typename T, bool CT> struct WeightT { using type = T; };
template <typename T> struct WeightT<T, true> { using type = cplx_t<T>; };
__device__ __forceinline__ float2 w_mul(float w, float2 v) { return make_float2(w * v.x, w * v.y); }
__device__ __forceinline__ double2 w_mul(double w, double2 v) { return make_double2(w * v.x, w * v.y); }
__device__ __forceinline__ float2 w_mul(float2 w, float2 v) { return cmul(w, v); }
__device__ __forceinline__ double2 w_mul(double2 w, double2 v) { return cmul(w, v); }
// acc + w * v (forward) and acc + conj(w) * v (adjoint) for real or complex weights
__device__ __forceinline__ float2 wfma(float w, float2 v, float2 acc) { return fma_w(w, v, acc); }
__device__ __forceinline__ double2 wfma(double w, double2 v, double2 acc) { return fma_w(w, v, acc); }
__device__ __forceinline__ float2 wfma(float2 w, float2 v, float2 acc) {
    return make_float2(fmaf(w.x, v.x, fmaf(-w.y, v.y, acc.x)), fmaf(w.x, v.y, fmaf(w.y, v.x, acc.y)));
}
__device__ __forceinline__ double2 wfma(double2 w, double2 v, double2 acc) {
    return make_double2(fma(w.x, v.x, fma(-w.y, v.y, acc.x)), fma(w.x, v.y, fma(w.y, v.x, acc.y)));
}
__device__ __forceinline__ float2 wfma_conj(float w, float2 v, float2 acc) { return fma_w(w, v, acc); }
__device__ __forceinline__ double2 wfma_conj(double w, double2 v, double2 acc) { return fma_w(w, v, acc); }
__device__ __forceinline__ float2 wfma_conj(float2 w, float2 v, float2 acc) {
    return make_float2(fmaf(w.x, v.x, fmaf(w.y, v.y, acc.x)), fmaf(w.x, v.y, fmaf(-w.y, v.x, acc.y)));
}
__device__ __forceinline__ double2 wfma_conj(double2 w, double2 v, double2 acc) {
    return make_double2(fma(w.x, v.x, fma(w.y, v.y, acc.x)), fma(w.x, v.y, fma(-w.y, v.x, acc.y)));
}
// conj(w) * v
__device__ __forceinline__ float2 w_mul_conj(float w, float2 v) { return make_float2(w * v.x, w * v.y); }
__device__ __forceinline__ double2 w_mul_conj(double w, double2 v) { return make_double2(w * v.x, w * v.y); }
__device__ __forceinline__ float2 w_mul_conj(float2 w, float2 v) { return cmul_conj(v, w); }
__device__ __forceinline__ double2 w_mul_conj(double2 w, double2 v) { return cmul_conj(v, w); }

}  // namespace b2n
