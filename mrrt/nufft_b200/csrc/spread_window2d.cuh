// Register-window adjoint gridding for 2-D multi-coil batches (real table, uniform J).
//
// Same idea as spread_window.cuh, with the COIL index playing the role of the third
// window axis: a warp walks a contiguous run of cell-sorted samples (forward order: axis 1
// fastest) and keeps, for a group of coils, the J x J window of every coil in registers:
// lane slot <-> (j2, coil), J accumulators along axis 1.  All coils share the sample's
// weights (computed once per sample, lane-parallel per 32-sample batch), each has its own
// sample value.  When the window slides along axis 1 the retiring column (J cells per
// coil) is sent to L2 with vector REDs: J*ncoil reductions per occupied cell instead of
// J*J*ncoil per sample (the one-RED-per-tap fallback is L2-reduction bound: 371 M REDs
// per 32-coil adjoint of BASELINE configs[3]).
//
// Arithmetic per sample and coil follows c/nufft_table.template.c:472-520 (2-D real
// adjoint): v2 = coef2*f, ck += coef1*v2.
#pragma once
#include "common.cuh"
#include "dispatch.h"

namespace b2n {

// RPL: slots per lane; coils per group = 32*RPL / J
template <typename T, int J, int RPL, bool HAVE_WTS>
__global__ void __launch_bounds__(128)
spread_window2d_kernel(Geom g, const T* __restrict__ h1, const T* __restrict__ h2,
                       const T* __restrict__ tm_s, const T* __restrict__ wts,
                       const int32_t* __restrict__ pt_ko, const int32_t* __restrict__ pt_kw,
                       const int32_t* __restrict__ perm, const cplx_t<T>* __restrict__ samples,
                       cplx_t<T>* __restrict__ grid, const cplx_t<T>* __restrict__ phase_s,
                       int pts_per_warp, int nbatch) {
    using C = cplx_t<T>;
    constexpr int NCG = 32 * RPL / J;                 // coils per group
    constexpr int NW = 2 * J + 2 * NCG;               // values per staging record
    constexpr int PITCH = NW % 2 == 1 ? NW : NW + 1;  // odd pitch (elements of T)
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    T* stage = (T*)dyn_smem + (size_t)wib * 32 * PITCH;
    int4* actions = (int4*)((T*)dyn_smem + (size_t)4 * 32 * PITCH + (4 * 32 * PITCH % 4 ? 4 - 4 * 32 * PITCH % 4 : 0)) + wib * 32;
    const int64_t M = g.M;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t begin = warp * pts_per_warp;
    if (begin >= M) return;
    const int64_t end = begin + pts_per_warp < M ? begin + pts_per_warp : M;
    const int coil0 = blockIdx.y * NCG;
    const int K1 = g.K[0], K2 = g.K[1];

    int rj2[RPL], rc[RPL];
    bool rvalid[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        const int r = lane + 32 * s;
        rj2[s] = r % J;
        rc[s] = r / J;
        rvalid[s] = rc[s] < NCG && coil0 + rc[s] < nbatch;
        if (!rvalid[s]) rc[s] = 0;
    }
    C acc[RPL][J];
    C* rowptr[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        rowptr[s] = grid;
#pragma unroll
        for (int j = 0; j < J; j++) acc[s][j] = make_c<T>(0, 0);
    }
    int W1 = 0;
    bool have = false;
    int pk1 = -(1 << 30), pk2 = -1;

    for (int64_t base = begin; base < end; base += 32) {
        const int cnt = (int)(end - base < 32 ? end - base : 32);
        __syncwarp();
        int k1 = 0, k2 = 0;
        if (lane < cnt) {
            const int64_t i = base + lane;
            T* w = stage + lane * PITCH;
            k1 = pt_kw[i];
            k2 = pt_kw[M + i];
            if (HAVE_WTS) {
#pragma unroll
                for (int j = 0; j < J; j++) {
                    w[j] = wts[(int64_t)j * M + i];
                    w[J + j] = wts[(int64_t)(J + j) * M + i];
                }
            } else {
                const T t1 = tm_s[i], t2 = tm_s[M + i];
                const int o1 = pt_ko[i], o2 = pt_ko[M + i];
#pragma unroll
                for (int j = 0; j < J; j++) {
                    w[j] = tap_real<T>(h1, g.ncenter[0], g.tlen[0], t1, o1 + j, g.L);
                    w[J + j] = tap_real<T>(h2, g.ncenter[1], g.tlen[1], t2, o2 + j, g.L);
                }
            }
            const int64_t src = perm[i];
            C ph = make_c<T>(1, 0);
            if (phase_s != nullptr) ph = phase_s[i];
#pragma unroll 4
            for (int c = 0; c < NCG; c++) {
                C f = make_c<T>(0, 0);
                if (coil0 + c < nbatch) {
                    f = samples[(int64_t)(coil0 + c) * M + src];
                    if (phase_s != nullptr) f = cmul_conj(f, ph);
                }
                w[2 * J + 2 * c] = f.x;
                w[2 * J + 2 * c + 1] = f.y;
            }
        }
        {
            int q1 = __shfl_up_sync(FULL, k1, 1), q2 = __shfl_up_sync(FULL, k2, 1);
            if (lane == 0) { q1 = pk1; q2 = pk2; }
            const int d = k1 - q1;
            const int act = (k2 == q2 && d >= 0 && d < J) ? d : -1;
            if (lane < cnt) actions[lane] = make_int4(k1, k2, 0, act);
            pk1 = __shfl_sync(FULL, k1, cnt - 1);
            pk2 = __shfl_sync(FULL, k2, cnt - 1);
        }
        __syncwarp();
        int4 kk_next = actions[0];
        for (int q = 0; q < cnt; q++) {
            const T* w = stage + q * PITCH;
            const int4 kk = kk_next;
            if (q + 1 < cnt) kk_next = actions[q + 1];
            T w1[J];
#pragma unroll
            for (int j = 0; j < J; j++) w1[j] = w[j];
            C v[RPL];
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                const T w2 = w[J + rj2[s]];
                const C f = make_c<T>(w[2 * J + 2 * rc[s]], w[2 * J + 2 * rc[s] + 1]);
                v[s] = mul_w(w2, f);
            }
            if (kk.w < 0) {
                if (have) {
#pragma unroll
                    for (int j = 0; j < J; j++) {
                        int ka = W1 + j;
                        if (ka >= K1) ka -= K1;
#pragma unroll
                        for (int s = 0; s < RPL; s++) {
                            if (rvalid[s]) atomic_add_c(rowptr[s] + ka, acc[s][j]);
                            acc[s][j] = make_c<T>(0, 0);
                        }
                    }
                }
                have = true;
                W1 = kk.x;
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    int kb = kk.y + rj2[s]; if (kb >= K2) kb -= K2;
                    rowptr[s] = grid + (int64_t)(coil0 + rc[s]) * g.PK + (int64_t)kb * K1;
                }
            } else {
#pragma unroll 1
                for (int sft = 0; sft < kk.w; sft++) {
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        if (rvalid[s]) atomic_add_c(rowptr[s] + W1, acc[s][0]);
#pragma unroll
                        for (int j = 0; j + 1 < J; j++) acc[s][j] = acc[s][j + 1];
                        acc[s][J - 1] = make_c<T>(0, 0);
                    }
                    W1++;
                }
            }
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                if (rvalid[s]) {
#pragma unroll
                    for (int j = 0; j < J; j++) acc[s][j] = fma_w(w1[j], v[s], acc[s][j]);
                }
            }
        }
    }
    if (have) {
#pragma unroll
        for (int j = 0; j < J; j++) {
            int ka = W1 + j;
            if (ka >= K1) ka -= K1;
#pragma unroll
            for (int s = 0; s < RPL; s++)
                if (rvalid[s]) atomic_add_c(rowptr[s] + ka, acc[s][j]);
        }
    }
}

template <typename T, int J, int RPL>
static int launch_window2d(const Geom& g, const TablePtrs& tabs, const void* tm_s, const void* wts,
                           const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                           const void* samples, void* grid, const void* phase_s, int nbatch,
                           int pts_per_warp, cudaStream_t st, bool* done) {
    using C = cplx_t<T>;
    constexpr int NCG = 32 * RPL / J;
    constexpr int NW = 2 * J + 2 * NCG;
    constexpr int PITCH = NW % 2 == 1 ? NW : NW + 1;
    const int64_t nwarps = (g.M + pts_per_warp - 1) / pts_per_warp;
    const int64_t nblocks = (nwarps + 3) / 4;
    const int ngroups = (nbatch + NCG - 1) / NCG;
    if (nblocks > 0x7fffffff || ngroups > 65535) return 0;
    size_t rec_bytes = (size_t)4 * 32 * PITCH * sizeof(T);
    rec_bytes = (rec_bytes + 15) / 16 * 16;
    const size_t smem = rec_bytes + (size_t)4 * 32 * sizeof(int4) + 16;
    dim3 gd((unsigned)nblocks, (unsigned)ngroups);
    cudaError_t e;
    if (wts != nullptr) {
        auto k = spread_window2d_kernel<T, J, RPL, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem, st>>>(g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tm_s,
                                 (const T*)wts, pt_ko, pt_kw, perm, (const C*)samples, (C*)grid,
                                 (const C*)phase_s, pts_per_warp, nbatch);
    } else {
        auto k = spread_window2d_kernel<T, J, RPL, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem, st>>>(g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tm_s,
                                 (const T*)wts, pt_ko, pt_kw, perm, (const C*)samples, (C*)grid,
                                 (const C*)phase_s, pts_per_warp, nbatch);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

template <typename T>
static int window2d_adj_t(const Geom& g, const TablePtrs& tabs, const void* tm_s, const void* wts,
                          const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                          const void* samples, void* grid, const void* phase_s, int nbatch,
                          int pts_per_warp, cudaStream_t st, bool* done) {
    *done = false;
    if (g.ndim != 2 || g.J[1] != g.J[0]) return 0;
    if (g.K[0] < g.J[0] || g.K[1] < g.J[0]) return 0;
    if (nbatch < 3) return 0;     // a few coils: the one-thread-per-sample kernel is as good
#define B2N_W2D(JJ, RR)                                                                          \
    return launch_window2d<T, JJ, RR>(g, tabs, tm_s, wts, pt_ko, pt_kw, perm, samples, grid,     \
                                      phase_s, nbatch, pts_per_warp, st, done)
#define B2N_W2D_J(JJ)                                                   \
    {                                                                   \
        if (nbatch * JJ <= 32) B2N_W2D(JJ, 1);                          \
        if (nbatch * JJ <= 64) B2N_W2D(JJ, 2);                          \
        if (nbatch * JJ <= 96) B2N_W2D(JJ, 3);                          \
        B2N_W2D(JJ, 6);                                                 \
    }
    switch (g.J[0]) {
        case 4: B2N_W2D_J(4)
        case 6: B2N_W2D_J(6)
        case 8: B2N_W2D_J(8)
        default: return 0;
    }
#undef B2N_W2D_J
#undef B2N_W2D
}

}  // namespace b2n
