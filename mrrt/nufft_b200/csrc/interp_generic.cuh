// Generic table interpolators: one thread per (bin-sorted) sample, grid read / updated
// directly in global memory (L2-resident thanks to the sort).  They cover every
// dimension (1..3), both precisions, real and complex tables and per-axis kernel
// widths, and are the fallback of the tiled / sliding-window kernels.
//
// Arithmetic follows c/nufft_table.template.c: forward = nested partial sums
// (:869-919), adjoint = nested partial products with conjugated kernel (:1122-1163).
#pragma once
#include "common.cuh"
#include "dispatch.h"

namespace b2n {

template <typename T, bool CT>
__device__ __forceinline__ typename WeightT<T, CT>::type
load_tap(const void* __restrict__ h, int ncenter, int tlen, T t, int k, int L, int order) {
    if constexpr (CT) {
        return tap_cplx<T>((const cplx_t<T>*)h, ncenter, tlen, t, k, L, order);
    } else {
        return tap_real<T>((const T*)h, ncenter, tlen, t, k, L, order);
    }
}

// JT > 0: all axes share the compile-time width JT (loops fully unrolled, weights in
// registers); JT == 0: per-axis run-time widths up to kMaxJ.
template <typename T, int NDIM, bool CT, int JT>
__global__ void __launch_bounds__(128)
interp_fwd_generic(Geom g, TablePtrs tabs, const T* __restrict__ tm_s,
                   const int32_t* __restrict__ perm, const cplx_t<T>* __restrict__ grid,
                   cplx_t<T>* __restrict__ out, const cplx_t<T>* __restrict__ phase_s,
                   int nbatch) {
    using C = cplx_t<T>;
    using W = typename WeightT<T, CT>::type;
    constexpr int JM = JT > 0 ? JT : kMaxJ;
    const int64_t M = g.M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        W w[NDIM][JM];
        int off[NDIM][JM];
        int Jd[NDIM];
        int stride = 1;
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            const int J = JT > 0 ? JT : g.J[d];
            Jd[d] = J;   // compile-time constant when JT > 0: the tap loops unroll
            const T t = tm_s[(int64_t)d * M + i];
            const int koff = window_origin<T>(t, J);
#pragma unroll(JT > 0 ? JT : 1)
            for (int j = 0; j < J; j++) {
                w[d][j] = load_tap<T, CT>(tabs.h[d], g.ncenter[d], g.tlen[d], t, koff + j, g.L, g.order);
                off[d][j] = local_index(koff + j, g.Kg[d], g.korg[d]) * stride;
            }
            stride *= g.K[d];
        }
        const int64_t dst = perm[i];
        C ph = make_c<T>(1, 0);
        if (phase_s != nullptr) ph = phase_s[i];
        for (int b = 0; b < nbatch; b++) {
            const C* __restrict__ gb = grid + (int64_t)b * g.PK;
            C s3 = make_c<T>(0, 0);
#pragma unroll(JT > 0 ? JT : 1)
            for (int j3 = 0; j3 < (NDIM > 2 ? Jd[NDIM > 2 ? 2 : 0] : 1); j3++) {
                C s2 = make_c<T>(0, 0);
#pragma unroll(JT > 0 ? JT : 1)
                for (int j2 = 0; j2 < (NDIM > 1 ? Jd[NDIM > 1 ? 1 : 0] : 1); j2++) {
                    int base = 0;
                    if (NDIM > 1) base += off[NDIM > 1 ? 1 : 0][j2];
                    if (NDIM > 2) base += off[NDIM > 2 ? 2 : 0][j3];
                    C s1 = make_c<T>(0, 0);
#pragma unroll(JT > 0 ? JT : 1)
                    for (int j1 = 0; j1 < Jd[0]; j1++) {
                        const C v = __ldg(gb + base + off[0][j1]);
                        const C p = w_mul(w[0][j1], v);
                        s1.x += p.x;
                        s1.y += p.y;
                    }
                    if (NDIM > 1) {
                        const C p = w_mul(w[NDIM > 1 ? 1 : 0][j2], s1);
                        s2.x += p.x;
                        s2.y += p.y;
                    } else {
                        s2 = s1;
                    }
                }
                if (NDIM > 2) {
                    const C p = w_mul(w[NDIM > 2 ? 2 : 0][j3], s2);
                    s3.x += p.x;
                    s3.y += p.y;
                } else {
                    s3 = s2;
                }
            }
            if (phase_s != nullptr) s3 = cmul(s3, ph);
            out[(int64_t)b * M + dst] = s3;
        }
    }
}

// TA: type the grid accumulates in.  The float instantiation is launched with TA = double on a
// scratch grid when the plan can afford one: the products are formed in float exactly as in
// the reference (template.c:1122-1163), but the per-cell sums no longer depend on the order in
// which L2 happens to apply the float atomics (measured on the mid-size 3-D test: 6.4e-6 ..
// 9.7e-6 from the reference between launches with float sums, scripts/diag_generic_noise.py).
__device__ __forceinline__ float2 to_acc(float2 v, float) { return v; }
__device__ __forceinline__ double2 to_acc(double2 v, double) { return v; }
__device__ __forceinline__ double2 to_acc(float2 v, double) { return make_double2((double)v.x, (double)v.y); }

template <typename T, int NDIM, bool CT, int JT, typename TA = T>
__global__ void __launch_bounds__(128)
interp_adj_generic(Geom g, TablePtrs tabs, const T* __restrict__ tm_s,
                   const int32_t* __restrict__ perm, const cplx_t<T>* __restrict__ samples,
                   cplx_t<TA>* __restrict__ grid, const cplx_t<T>* __restrict__ phase_s,
                   int nbatch) {
    using C = cplx_t<T>;
    using W = typename WeightT<T, CT>::type;
    constexpr int JM = JT > 0 ? JT : kMaxJ;
    const int64_t M = g.M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        W w[NDIM][JM];
        int off[NDIM][JM];
        int Jd[NDIM];
        int stride = 1;
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            const int J = JT > 0 ? JT : g.J[d];
            Jd[d] = J;   // compile-time constant when JT > 0: the tap loops unroll
            const T t = tm_s[(int64_t)d * M + i];
            const int koff = window_origin<T>(t, J);
#pragma unroll(JT > 0 ? JT : 1)
            for (int j = 0; j < J; j++) {
                w[d][j] = load_tap<T, CT>(tabs.h[d], g.ncenter[d], g.tlen[d], t, koff + j, g.L, g.order);
                off[d][j] = local_index(koff + j, g.Kg[d], g.korg[d]) * stride;
            }
            stride *= g.K[d];
        }
        const int64_t src = perm[i];
        C ph = make_c<T>(1, 0);
        if (phase_s != nullptr) ph = phase_s[i];
        for (int b = 0; b < nbatch; b++) {
            cplx_t<TA>* __restrict__ gb = grid + (int64_t)b * g.PK;
            C f = samples[(int64_t)b * M + src];
            if (phase_s != nullptr) f = cmul_conj(f, ph);
#pragma unroll(JT > 0 ? JT : 1)
            for (int j3 = 0; j3 < (NDIM > 2 ? Jd[NDIM > 2 ? 2 : 0] : 1); j3++) {
                C v3 = f;
                if (NDIM > 2) v3 = w_mul_conj(w[NDIM > 2 ? 2 : 0][j3], f);
#pragma unroll(JT > 0 ? JT : 1)
                for (int j2 = 0; j2 < (NDIM > 1 ? Jd[NDIM > 1 ? 1 : 0] : 1); j2++) {
                    C v2 = v3;
                    if (NDIM > 1) v2 = w_mul_conj(w[NDIM > 1 ? 1 : 0][j2], v3);
                    int base = 0;
                    if (NDIM > 1) base += off[NDIM > 1 ? 1 : 0][j2];
                    if (NDIM > 2) base += off[NDIM > 2 ? 2 : 0][j3];
#pragma unroll(JT > 0 ? JT : 1)
                    for (int j1 = 0; j1 < Jd[0]; j1++) {
                        atomic_add_c(gb + base + off[0][j1], to_acc(w_mul_conj(w[0][j1], v2), (TA)0));
                    }
                }
            }
        }
    }
}

#ifdef B2N_GENERIC_TU
// out += acc (the double scratch grid of the float adjoint)
static __global__ void add_acc64_kernel(int64_t n, const double2* __restrict__ acc, float2* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double2 a = acc[i];
        float2 o = out[i];
        o.x += (float)a.x;
        o.y += (float)a.y;
        out[i] = o;
    }
}

template <typename T, int NDIM, bool CT, int JT>
static void launch_generic(const Geom& g, const TablePtrs& tabs, const void* tm_s, const int32_t* perm,
                           bool fwd, const void* in, void* out, const void* phase_s, int nbatch,
                           int sm_count, void* acc64, cudaStream_t st) {
    using C = cplx_t<T>;
    int64_t nb = (g.M + 127) / 128;
    if (nb > (int64_t)sm_count * 32) nb = (int64_t)sm_count * 32;
    const int grid = (int)(nb < 1 ? 1 : nb);
    if constexpr (sizeof(T) == 4) {
        if (!fwd && acc64 != nullptr) {
            // float adjoint with per-cell sums in double (scratch grid, zeroed by the caller)
            interp_adj_generic<T, NDIM, CT, JT, double><<<grid, 128, 0, st>>>(
                g, tabs, (const T*)tm_s, perm, (const C*)in, (double2*)acc64, (const C*)phase_s, nbatch);
            const int64_t n = g.PK * nbatch;
            int64_t cb = (n + 255) / 256;
            if (cb > (int64_t)sm_count * 16) cb = (int64_t)sm_count * 16;
            add_acc64_kernel<<<(int)(cb < 1 ? 1 : cb), 256, 0, st>>>(n, (const double2*)acc64, (float2*)out);
            return;
        }
    }
    if (fwd)
        interp_fwd_generic<T, NDIM, CT, JT><<<grid, 128, 0, st>>>(
            g, tabs, (const T*)tm_s, perm, (const C*)in, (C*)out, (const C*)phase_s, nbatch);
    else
        interp_adj_generic<T, NDIM, CT, JT><<<grid, 128, 0, st>>>(
            g, tabs, (const T*)tm_s, perm, (const C*)in, (C*)out, (const C*)phase_s, nbatch);
}

template <typename T, int NDIM, bool CT>
static void dispatch_generic_J(const Geom& g, const TablePtrs& tabs, const void* tm_s,
                               const int32_t* perm, bool fwd, const void* in, void* out,
                               const void* phase_s, int nbatch, int sm_count, void* acc64, cudaStream_t st) {
    int J = g.J[0];
    for (int d = 1; d < g.ndim; d++)
        if (g.J[d] != g.J[0]) J = 0;
    switch (J) {
        case 4: launch_generic<T, NDIM, CT, 4>(g, tabs, tm_s, perm, fwd, in, out, phase_s, nbatch, sm_count, acc64, st); break;
        case 6: launch_generic<T, NDIM, CT, 6>(g, tabs, tm_s, perm, fwd, in, out, phase_s, nbatch, sm_count, acc64, st); break;
        default: launch_generic<T, NDIM, CT, 0>(g, tabs, tm_s, perm, fwd, in, out, phase_s, nbatch, sm_count, acc64, st); break;
    }
}

template <typename T>
static int generic_launch_t(const Geom& g, int cplx_table, const TablePtrs& tabs, const void* tm_s,
                            const int32_t* perm, bool fwd, const void* in, void* out,
                            const void* phase_s, int nbatch, int sm_count, void* acc64, cudaStream_t st) {
#define B2N_GO(ND, CTV) dispatch_generic_J<T, ND, CTV>(g, tabs, tm_s, perm, fwd, in, out, phase_s, nbatch, sm_count, acc64, st)
    if (cplx_table) {
        if (g.ndim == 1) B2N_GO(1, true);
        if (g.ndim == 2) B2N_GO(2, true);
        if (g.ndim == 3) B2N_GO(3, true);
    } else {
        if (g.ndim == 1) B2N_GO(1, false);
        if (g.ndim == 2) B2N_GO(2, false);
        if (g.ndim == 3) B2N_GO(3, false);
    }
#undef B2N_GO
    return (int)cudaGetLastError();
}
#endif  // B2N_GENERIC_TU

}  // namespace b2n
