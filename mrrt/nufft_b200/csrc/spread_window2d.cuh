// Register-window adjoint gridding in 2-D (real table, uniform J), one coil or a batch.
//
// Same idea as spread_window.cuh.  A lane group (G = 8, 16 or 32 lanes) walks a contiguous
// run of samples in the ADJOINT sort order (cells ordered axis-2-fastest inside a bin) and
// keeps, for a group of coils, the J x J window of every coil in registers:
// lane slot <-> (j1, coil) with j1 fastest, J accumulators along axis 2, the slide axis.
// All coils share the sample's weights (prepared once per sample, lane-parallel per batch of
// G samples); each coil has its own sample value.  When the window slides along axis 2 the
// retiring row (J CONSECUTIVE cells along axis 1 per coil: 1-2 sectors) goes to L2 with
// vector REDs: J reductions per occupied cell and coil instead of J*J per sample and coil
// (the one-RED-per-tap fallback is L2-reduction bound and carries one float32 rounding per
// tap in the grid; here a cell's partial sums stay in registers while the window covers it).
//
// Lanes per sample follow the batch size: one coil -> 8 lanes (four runs per warp, J <= 8
// face positions each), two coils -> 16 lanes, more -> 32 lanes x RPL slots.
//
// Arithmetic per sample and coil follows c/nufft_table.template.c:472-520 (2-D real
// adjoint): v1 = coef1*f, ck += coef2*v1 (the same two products per tap, grouped by role).
#pragma once
#include "common.cuh"
#include "dispatch.h"

namespace b2n {

// G: lanes per sample; RPL: slots per lane; coils per group = G*RPL / J
// CT: complex table (phasing="complex"): complex plan-time weights, conjugated products
//     (template.c:490-491); HAVE_WTS only.
template <typename T, int J, int G, int RPL, bool HAVE_WTS, bool CT = false>
__global__ void __launch_bounds__(128)
spread_window2d_kernel(Geom g, const T* __restrict__ h1, const T* __restrict__ h2,
                       const T* __restrict__ tm_s, const T* __restrict__ wts,
                       const int32_t* __restrict__ pt_ko, const int32_t* __restrict__ pt_kw,
                       const int32_t* __restrict__ perm, const cplx_t<T>* __restrict__ samples,
                       cplx_t<T>* __restrict__ grid, const cplx_t<T>* __restrict__ phase_s,
                       int pts_per_warp, int nbatch, int max_slide) {
    using C = cplx_t<T>;
    constexpr int NG = 32 / G;                        // sample runs per warp
    constexpr int NCG = G * RPL / J;                  // coils per group
    constexpr int WV = CT ? 2 : 1;                    // values per weight
    constexpr int NW = 2 * J * WV + 2 * NCG;          // values per staging record
    constexpr int PITCH = NW % 2 == 1 ? NW : NW + 1;  // odd pitch (elements of T)
    constexpr unsigned FULL = 0xffffffffu;
    using W = typename WeightT<T, CT>::type;
    static_assert(NCG >= 1, "a lane group must hold at least one coil's row");
    static_assert(!CT || HAVE_WTS, "complex tables: plan-time weights");
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    T* stage = (T*)dyn_smem + (size_t)wib * 32 * PITCH;
    int4* actions = (int4*)((T*)dyn_smem + (size_t)4 * 32 * PITCH + (4 * 32 * PITCH % 4 ? 4 - 4 * 32 * PITCH % 4 : 0)) + wib * 32;
    const int64_t M = g.M;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (warp * pts_per_warp >= M) return;
    const int grp = lane / G;
    const int lg = lane - grp * G;
    const int per_group = pts_per_warp / NG;
    const int64_t begin = warp * pts_per_warp + (int64_t)grp * per_group;
    const int64_t end = begin + per_group < M ? begin + per_group : (begin < M ? M : begin);
    const int coil0 = blockIdx.y * NCG;
    const int K1 = g.K[0], K2 = g.K[1];

    int rj1[RPL], rc[RPL];
    bool rvalid[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        const int r = lg + G * s;
        rj1[s] = r % J;
        rc[s] = r / J;
        rvalid[s] = rc[s] < NCG && coil0 + rc[s] < nbatch;
        if (!rvalid[s]) rc[s] = 0;
    }
    C acc[RPL][J];
    C* colptr[RPL];         // grid address of this slot's (axis-1 position, coil)
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        colptr[s] = grid;
#pragma unroll
        for (int j = 0; j < J; j++) acc[s][j] = make_c<T>(0, 0);
    }
    int W2 = 0;             // wrapped window origin along axis 2 (the slide axis)
    bool have = false;
    int pk1 = -1, pk2 = -(1 << 30);

    for (int it = 0; it < per_group; it += G) {
        const int64_t base = begin + it;
        const int cnt = (int)(base >= end ? 0 : (end - base < G ? end - base : G));
        __syncwarp();
        // ---- batch phase: lane = sample (record index = lane)
        int k1 = 0, k2 = 0;
        if (lg < cnt) {
            const int64_t i = base + lg;
            T* w = stage + lane * PITCH;
            k1 = pt_kw[i];
            k2 = pt_kw[M + i];
            if constexpr (CT) {
                const W* __restrict__ wc_ = (const W*)wts;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const W a = wc_[(int64_t)(J + j) * M + i], bq = wc_[(int64_t)j * M + i];
                    w[2 * j] = a.x; w[2 * j + 1] = a.y;
                    w[2 * (J + j)] = bq.x; w[2 * (J + j) + 1] = bq.y;
                }
            } else if (HAVE_WTS) {
#pragma unroll
                for (int j = 0; j < J; j++) {
                    w[j] = wts[(int64_t)(J + j) * M + i];      // axis 2: along the registers
                    w[J + j] = wts[(int64_t)j * M + i];        // axis 1: across the lanes
                }
            } else {
                const T t1 = tm_s[i], t2 = tm_s[M + i];
                const int o1 = pt_ko[i], o2 = pt_ko[M + i];
#pragma unroll
                for (int j = 0; j < J; j++) {
                    w[j] = tap_real<T>(h2, g.ncenter[1], g.tlen[1], t2, o2 + j, g.L, g.order);
                    w[J + j] = tap_real<T>(h1, g.ncenter[0], g.tlen[0], t1, o1 + j, g.L, g.order);
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
                w[2 * J * WV + 2 * c] = f.x;
                w[2 * J * WV + 2 * c + 1] = f.y;
            }
        }
        {
            int q1 = __shfl_up_sync(FULL, k1, 1, G), q2 = __shfl_up_sync(FULL, k2, 1, G);
            if (lg == 0) { q1 = pk1; q2 = pk2; }
            const int d = k2 - q2;
            const int act = (k1 == q1 && d >= 0 && d <= max_slide) ? d : -1;
            if (lg < cnt) actions[lane] = make_int4(k1, k2, 0, act);
            const int last = cnt > 0 ? cnt - 1 : 0;
            const int n1 = __shfl_sync(FULL, k1, last, G), n2 = __shfl_sync(FULL, k2, last, G);
            if (cnt > 0) { pk1 = n1; pk2 = n2; }
        }
        __syncwarp();
        // ---- sample loop: all lanes of a group work on one sample
        int4 kk_next = actions[grp * G];
        for (int q = 0; q < cnt; q++) {
            const T* w = stage + (grp * G + q) * PITCH;
            const int4 kk = kk_next;
            if (q + 1 < cnt) kk_next = actions[grp * G + q + 1];
            W w2[J];
            C v[RPL];
            if constexpr (CT) {
#pragma unroll
                for (int j = 0; j < J; j++) w2[j] = make_c<T>(w[2 * j], w[2 * j + 1]);
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    const W w1 = make_c<T>(w[2 * (J + rj1[s])], w[2 * (J + rj1[s]) + 1]);
                    const C f = make_c<T>(w[4 * J + 2 * rc[s]], w[4 * J + 2 * rc[s] + 1]);
                    v[s] = w_mul_conj(w1, f);
                }
            } else {
#pragma unroll
                for (int j = 0; j < J; j++) w2[j] = w[j];
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    const T w1 = w[J + rj1[s]];
                    const C f = make_c<T>(w[2 * J + 2 * rc[s]], w[2 * J + 2 * rc[s] + 1]);
                    v[s] = mul_w(w1, f);
                }
            }
            if (kk.w < 0) {
                if (have) {
#pragma unroll
                    for (int j = 0; j < J; j++) {
                        int ka = W2 + j;
                        if (ka >= K2) ka -= K2;
#pragma unroll
                        for (int s = 0; s < RPL; s++) {
                            if (rvalid[s]) atomic_add_c(colptr[s] + (int64_t)ka * K1, acc[s][j]);
                            acc[s][j] = make_c<T>(0, 0);
                        }
                    }
                }
                have = true;
                W2 = kk.y;
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    int kb = kk.x + rj1[s]; if (kb >= K1) kb -= K1;
                    colptr[s] = grid + (int64_t)(coil0 + rc[s]) * g.PK + kb;
                }
            } else if (kk.w > 0) {
#pragma unroll 1
                for (int sft = 0; sft < kk.w - 1; sft++) {
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        if (rvalid[s]) atomic_add_c(colptr[s] + (int64_t)W2 * K1, acc[s][0]);
#pragma unroll
                        for (int j = 0; j + 1 < J; j++) acc[s][j] = acc[s][j + 1];
                        acc[s][J - 1] = make_c<T>(0, 0);
                    }
                    W2++;   // stays < K2: it ends at this sample's wrapped origin
                }
                // last cell of the slide: the FMAs do the shift (destination j, addend j + 1)
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    if (rvalid[s]) {
                        atomic_add_c(colptr[s] + (int64_t)W2 * K1, acc[s][0]);
#pragma unroll
                        for (int j = 0; j + 1 < J; j++) acc[s][j] = wfma_conj(w2[j], v[s], acc[s][j + 1]);
                        acc[s][J - 1] = wfma_conj(w2[J - 1], v[s], make_c<T>(0, 0));
                    }
                }
                W2++;
                continue;
            }
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                if (rvalid[s]) {
#pragma unroll
                    for (int j = 0; j < J; j++) acc[s][j] = wfma_conj(w2[j], v[s], acc[s][j]);
                }
            }
        }
    }
    if (have) {
#pragma unroll
        for (int j = 0; j < J; j++) {
            int ka = W2 + j;
            if (ka >= K2) ka -= K2;
#pragma unroll
            for (int s = 0; s < RPL; s++)
                if (rvalid[s]) atomic_add_c(colptr[s] + (int64_t)ka * K1, acc[s][j]);
        }
    }
}

template <typename T, int J, int G, int RPL>
static int launch_window2d(const Geom& g, bool cplx, const TablePtrs& tabs, const void* tm_s, const void* wts,
                           const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                           const void* samples, void* grid, const void* phase_s, int nbatch,
                           const WindowOpts& wo, cudaStream_t st, bool* done) {
    using C = cplx_t<T>;
    constexpr int NCG = G * RPL / J;
    const int NW = 2 * J * (cplx ? 2 : 1) + 2 * NCG;
    const int PITCH = NW % 2 == 1 ? NW : NW + 1;
    int max_slide = wo.max_slide;
    if (max_slide <= 0 || max_slide > J - 1) max_slide = J - 1;
    const int pts_per_warp = (wo.pts_per_warp + 31) / 32 * 32;
    const int64_t nwarps = (g.M + pts_per_warp - 1) / pts_per_warp;
    const int64_t nblocks = (nwarps + 3) / 4;
    const int ngroups = (nbatch + NCG - 1) / NCG;
    if (nblocks > 0x7fffffff || ngroups > 65535) return 0;
    size_t rec_bytes = (size_t)4 * 32 * PITCH * sizeof(T);
    rec_bytes = (rec_bytes + 15) / 16 * 16;
    const size_t smem = rec_bytes + (size_t)4 * 32 * sizeof(int4) + 16;
    dim3 gd((unsigned)nblocks, (unsigned)ngroups);
    cudaError_t e;
    if (cplx) {
        if (wts == nullptr) return 0;                 // complex tables: plan-time weights only
        auto k = spread_window2d_kernel<T, J, G, RPL, true, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem, st>>>(g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tm_s,
                                 (const T*)wts, pt_ko, pt_kw, perm, (const C*)samples, (C*)grid,
                                 (const C*)phase_s, pts_per_warp, nbatch, max_slide);
    } else if (wts != nullptr) {
        auto k = spread_window2d_kernel<T, J, G, RPL, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem, st>>>(g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tm_s,
                                 (const T*)wts, pt_ko, pt_kw, perm, (const C*)samples, (C*)grid,
                                 (const C*)phase_s, pts_per_warp, nbatch, max_slide);
    } else {
        auto k = spread_window2d_kernel<T, J, G, RPL, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem, st>>>(g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tm_s,
                                 (const T*)wts, pt_ko, pt_kw, perm, (const C*)samples, (C*)grid,
                                 (const C*)phase_s, pts_per_warp, nbatch, max_slide);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

// the sample arrays must be in the ADJOINT sort order (axis 2 fastest inside a bin)
template <typename T>
static int window2d_adj_t(const Geom& g, int Jk, bool cplx, const TablePtrs& tabs, const void* tm_s, const void* wts,
                          const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                          const void* samples, void* grid, const void* phase_s, int nbatch,
                          const WindowOpts& wo, cudaStream_t st, bool* done) {
    *done = false;
    if (g.ndim != 2) return 0;
    for (int d = 0; d < 2; d++) {
        if (g.J[d] != Jk && wts == nullptr) return 0;   // padded windows need the plan-time weights
        if (g.K[d] < Jk) return 0;
    }
#define B2N_W2D(JJ, GG, RR)                                                                      \
    return launch_window2d<T, JJ, GG, RR>(g, cplx, tabs, tm_s, wts, pt_ko, pt_kw, perm, samples, grid, \
                                          phase_s, nbatch, wo, st, done)
#define B2N_W2D_J(JJ)                                                   \
    {                                                                   \
        if (nbatch == 1) B2N_W2D(JJ, 8, 1);                             \
        if (nbatch * JJ <= 16) B2N_W2D(JJ, 16, 1);                      \
        if (nbatch * JJ <= 32) B2N_W2D(JJ, 32, 1);                      \
        if (nbatch * JJ <= 64) B2N_W2D(JJ, 32, 2);                      \
        if (nbatch * JJ <= 96) B2N_W2D(JJ, 32, 3);                      \
        B2N_W2D(JJ, 32, 6);                                             \
    }
    switch (Jk) {
        case 4: B2N_W2D_J(4)
        case 6: B2N_W2D_J(6)
        case 8: B2N_W2D_J(8)
        default: return 0;
    }
#undef B2N_W2D_J
#undef B2N_W2D
}

}  // namespace b2n
