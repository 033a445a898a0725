// Sliding-window adjoint gridding (3-D, real table, uniform compile-time J).
//
// Shared-memory float atomics are CAS loops on sm_100a (ATOMS.CAST.SPIN), so the
// adjoint does not accumulate in shared memory at all.  Instead each warp walks a
// contiguous run of the cell-sorted samples ONE SAMPLE AT A TIME and keeps the sample's
// whole J x J x J window of partial sums in REGISTERS: lane <-> (j2, j3) row of the
// window, J accumulators per row along the fastest axis.  Because consecutive samples
// of the sorted order sit in the same or the next cell along axis 1, the window slides:
// only the column that leaves the window is flushed to HBM/L2 (one vector RED per row,
// REDG.E.ADD.F32x2), everything else stays in registers.  A cell therefore receives
// J*J reductions in total instead of one per contributing sample and tap, and there are
// no write conflicts inside a warp by construction.
//
// Arithmetic per sample follows c/nufft_table.template.c:1122-1163: v3 = coef3*f,
// v2 = coef2*v3, ck += coef1*v2.
#pragma once
#include "common.cuh"
#include "dispatch.h"

namespace b2n {

// staging record of one sample (16-byte aligned so that a lane-uniform read is a few
// LDS.128 broadcasts instead of one shuffle per field)
template <typename T> struct StagePt;
template <> struct alignas(16) StagePt<float> {
    float t[3]; float fx;
    float fy; int ko[3];
    int kw[3]; int pad;
};
template <> struct alignas(16) StagePt<double> {
    double t[3]; double fx;
    double fy; int ko[3]; int kw[3];
};

template <typename T, int J, bool TAB_SMEM>
__global__ void __launch_bounds__(128)
spread_slide3d_kernel(Geom g, const T* __restrict__ h1, const T* __restrict__ h2,
                      const T* __restrict__ h3, const T* __restrict__ tm_s,
                      const int32_t* __restrict__ pt_ko, const int32_t* __restrict__ pt_kw,
                      const int32_t* __restrict__ perm, const cplx_t<T>* __restrict__ samples,
                      cplx_t<T>* __restrict__ grid, const cplx_t<T>* __restrict__ phase_s,
                      int pts_per_warp) {
    using C = cplx_t<T>;
    constexpr int R = J * J;                 // rows of the window
    constexpr int RPL = (R + 31) / 32;       // rows per lane
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ StagePt<T> stage[4][32];
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t M = g.M;
    if (TAB_SMEM) {
        // all axes share one table (checked by the launcher): stage it in shared memory,
        // a gather from L1 costs one tag cycle per distinct sector
        T* stab = (T*)dyn_smem;
        for (int e = threadIdx.x; e < g.tlen[0]; e += blockDim.x) stab[e] = __ldg(h1 + e);
        __syncthreads();
    }
    const int64_t begin = warp * pts_per_warp;
    if (begin >= M) return;
    const int64_t end = begin + pts_per_warp < M ? begin + pts_per_warp : M;
    const int b = blockIdx.y;
    const C* __restrict__ sb = samples + (int64_t)b * M;
    C* __restrict__ gb = grid + (int64_t)b * g.PK;
    const int K1 = g.K[0], K2 = g.K[1], K3 = g.K[2];

    // static role of this lane: its rows (j2, j3) and, for the cooperative weight
    // evaluation, one (axis, tap) pair
    int rj2[RPL], rj3[RPL];
    bool rvalid[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        const int r = lane + 32 * s;
        rvalid[s] = r < R;
        rj2[s] = (r % R) % J;
        rj3[s] = (r % R) / J;
    }
    const int wax = lane < J ? 0 : (lane < 2 * J ? 1 : 2);   // axis of this lane's tap
    const int wj = lane - wax * J;
    const bool wactive = lane < 3 * J;
    const T* __restrict__ wh = TAB_SMEM ? (const T*)dyn_smem : (wax == 0 ? h1 : (wax == 1 ? h2 : h3));
    const int wnc = g.ncenter[wax], wtl = g.tlen[wax];
    const T Lf = (T)g.L;

    C acc[RPL][J];
    int rowbase[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        rowbase[s] = 0;
#pragma unroll
        for (int j = 0; j < J; j++) acc[s][j] = make_c<T>(0, 0);
    }
    int W1 = 0, W2 = -1, W3 = -1;            // wrapped window origin; W2<0: no window yet

    for (int64_t base = begin; base < end; base += 32) {
        const int cnt = (int)(end - base < 32 ? end - base : 32);
        __syncwarp();
        if (lane < cnt) {
            // coalesced loads of up to 32 samples into the staging records
            const int64_t i = base + lane;
            StagePt<T> p;
            p.t[0] = tm_s[i]; p.t[1] = tm_s[M + i]; p.t[2] = tm_s[2 * M + i];
            p.ko[0] = pt_ko[i]; p.ko[1] = pt_ko[M + i]; p.ko[2] = pt_ko[2 * M + i];
            p.kw[0] = pt_kw[i]; p.kw[1] = pt_kw[M + i]; p.kw[2] = pt_kw[2 * M + i];
            C f = sb[perm[i]];
            if (phase_s != nullptr) f = cmul_conj(f, phase_s[i]);
            p.fx = f.x; p.fy = f.y;
            stage[wib][lane] = p;
        }
        __syncwarp();
        StagePt<T> cur = stage[wib][0];
        for (int q = 0; q < cnt; q++) {
            StagePt<T> nxt = cur;
            if (q + 1 < cnt) nxt = stage[wib][q + 1];   // prefetch the next record
            // cooperative weights: lane (axis, tap) evaluates one table coefficient
            // (template.c:870-873)
            T wl = 0;
            if (wactive) {
                const T ta = wax == 0 ? cur.t[0] : (wax == 1 ? cur.t[1] : cur.t[2]);
                const int ka = (wax == 0 ? cur.ko[0] : (wax == 1 ? cur.ko[1] : cur.ko[2])) + wj;
                const T p = (ta - (T)ka) * Lf;
                const T fl = floor(p);
                const T alf = p - fl;
                const int i0 = wnc + (int)fl;
                const int i1 = min(i0 + 1, wtl - 1);
                wl = ((T)1 - alf) * wh[i0] + alf * wh[i1];
            }
            const int d = cur.kw[0] - W1;
            if (cur.kw[1] != W2 || cur.kw[2] != W3 || d < 0 || d >= J) {
                // new row of cells (or a jump): flush the whole window
                if (W2 >= 0) {
#pragma unroll
                    for (int j = 0; j < J; j++) {
                        int k1 = W1 + j;
                        if (k1 >= K1) k1 -= K1;
#pragma unroll
                        for (int s = 0; s < RPL; s++) {
                            if (rvalid[s]) atomic_add_c(gb + rowbase[s] + k1, acc[s][j]);
                            acc[s][j] = make_c<T>(0, 0);
                        }
                    }
                }
                W1 = cur.kw[0]; W2 = cur.kw[1]; W3 = cur.kw[2];
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    int k2 = W2 + rj2[s]; if (k2 >= K2) k2 -= K2;
                    int k3 = W3 + rj3[s]; if (k3 >= K3) k3 -= K3;
                    rowbase[s] = (k3 * K2 + k2) * K1;
                }
            } else {
                for (int sft = 0; sft < d; sft++) {
                    // slide by one cell: retire column 0
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        if (rvalid[s]) atomic_add_c(gb + rowbase[s] + W1, acc[s][0]);
#pragma unroll
                        for (int j = 0; j + 1 < J; j++) acc[s][j] = acc[s][j + 1];
                        acc[s][J - 1] = make_c<T>(0, 0);
                    }
                    W1++;   // < K1: the new origin cur.kw[0] is a wrapped index
                }
            }
            // accumulate this sample into the register window
            T w1[J];
#pragma unroll
            for (int j = 0; j < J; j++) w1[j] = __shfl_sync(FULL, wl, j);
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                const T w2 = __shfl_sync(FULL, wl, J + rj2[s]);
                const T w3 = __shfl_sync(FULL, wl, 2 * J + rj3[s]);
                if (rvalid[s]) {
                    const T v3x = w3 * cur.fx, v3y = w3 * cur.fy;
                    const T v2x = w2 * v3x, v2y = w2 * v3y;
#pragma unroll
                    for (int j = 0; j < J; j++) {
                        acc[s][j].x += w1[j] * v2x;
                        acc[s][j].y += w1[j] * v2y;
                    }
                }
            }
            cur = nxt;
        }
    }
    if (W2 >= 0) {
#pragma unroll
        for (int j = 0; j < J; j++) {
            int k1 = W1 + j;
            if (k1 >= K1) k1 -= K1;
#pragma unroll
            for (int s = 0; s < RPL; s++)
                if (rvalid[s]) atomic_add_c(gb + rowbase[s] + k1, acc[s][j]);
        }
    }
}

template <typename T, int J>
static int launch_slide(const Geom& g, const TablePtrs& tabs, const void* tm_s,
                        const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                        const void* samples, void* grid, const void* phase_s, int nbatch,
                        int pts_per_warp, cudaStream_t st, bool* done) {
    using C = cplx_t<T>;
    const int64_t nwarps = (g.M + pts_per_warp - 1) / pts_per_warp;
    const int64_t nblocks = (nwarps + 3) / 4;
    if (nblocks > 0x7fffffff || nbatch > 65535) return 0;
    dim3 gd((unsigned)nblocks, (unsigned)nbatch);
    const bool tab_smem = tabs.h[0] == tabs.h[1] && tabs.h[1] == tabs.h[2] &&
                          (size_t)g.tlen[0] * sizeof(T) <= 40 * 1024;
    cudaError_t e;
    if (tab_smem) {
        const size_t smem = (size_t)g.tlen[0] * sizeof(T);
        auto k = spread_slide3d_kernel<T, J, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem, st>>>(g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tabs.h[2],
                                 (const T*)tm_s, pt_ko, pt_kw, perm, (const C*)samples, (C*)grid,
                                 (const C*)phase_s, pts_per_warp);
    } else {
        spread_slide3d_kernel<T, J, false><<<gd, 128, 0, st>>>(
            g, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tabs.h[2], (const T*)tm_s, pt_ko,
            pt_kw, perm, (const C*)samples, (C*)grid, (const C*)phase_s, pts_per_warp);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

template <typename T>
static int slide_adj_t(const Geom& g, const TablePtrs& tabs, const void* tm_s,
                       const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                       const void* samples, void* grid, const void* phase_s, int nbatch,
                       int pts_per_warp, cudaStream_t st, bool* done) {
    *done = false;
    if (g.ndim != 3) return 0;
    if (g.J[1] != g.J[0] || g.J[2] != g.J[0]) return 0;
    // the window must not wrap onto itself
    if (g.K[0] < g.J[0] || g.K[1] < g.J[0] || g.K[2] < g.J[0]) return 0;
#define B2N_SLIDE(JJ) \
    return launch_slide<T, JJ>(g, tabs, tm_s, pt_ko, pt_kw, perm, samples, grid, phase_s, nbatch, pts_per_warp, st, done)
    switch (g.J[0]) {
        case 4: B2N_SLIDE(4);
        case 5: B2N_SLIDE(5);
        case 6: B2N_SLIDE(6);
        case 7: B2N_SLIDE(7);
        case 8: B2N_SLIDE(8);
        default: return 0;
    }
#undef B2N_SLIDE
}

}  // namespace b2n
