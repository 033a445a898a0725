// Tiled sliding-window adjoint gridding (3-D, real table, uniform compile-time J).
//
// One CTA = one work item = up to `chunk` cell-sorted samples of one bin (the same work
// list as the tiled forward kernel).  The bin's grid tile (+ J-1 halo) is ACCUMULATED in
// shared memory and flushed to HBM once: by one TMA tensor reduction
// (cp.reduce.async.bulk.tensor ... .add) when the box does not cross the periodic
// boundary, by coalesced vector REDs otherwise.
//
// Measurements behind the design (profiles/r01_notes.md): L2 reductions cap at ~270 G
// complex64 cell-adds/s whatever the path, so the number of cell-adds that reach L2 is
// what matters; one-RED-per-tap costs 216 per sample, the register sliding window alone
// ~26 per sample (and each lane hits its own sector), this kernel ~3 per sample.
//
// Inside the CTA each warp walks a contiguous run of the item's samples one sample at a
// time and keeps the sample's J x J x J window of partial sums in REGISTERS (lane <->
// (j2, j3) row, J accumulators along axis 1).  The window slides with the sorted order;
// only the column leaving the window is added to the shared-memory tile.  Lanes of a warp
// hit distinct rows by construction, so the shared-memory float atomics (a CAS loop on
// sm_100a) only ever contend between different warps, which is rare.
//
// Arithmetic per sample follows c/nufft_table.template.c:1122-1163: v3 = coef3*f,
// v2 = coef2*v3, ck += coef1*v2.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dispatch.h"
#include "interp_tiled.cuh"
#include "spread_slide.cuh"

namespace b2n {

__device__ __forceinline__ void smem_add_c(float2* p, float2 v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}
__device__ __forceinline__ void smem_add_c(double2* p, double2 v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0,
                                                  int c1, int c2, int c3) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group "
        "[%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
        "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

constexpr int kTileWarps = 8;

template <typename T, int J>
__global__ void __launch_bounds__(kTileWarps * 32, 2)
spread_tile3d_kernel(const __grid_constant__ CUtensorMap tmap, Geom g, TileShape ts, int use_tma,
                     const T* __restrict__ h1, const T* __restrict__ h2, const T* __restrict__ h3,
                     const T* __restrict__ tm_s, const int32_t* __restrict__ pt_ko,
                     const int32_t* __restrict__ pt_kw, const int32_t* __restrict__ perm,
                     const int4* __restrict__ items, const cplx_t<T>* __restrict__ samples,
                     cplx_t<T>* __restrict__ grid, const cplx_t<T>* __restrict__ phase_s) {
    using C = cplx_t<T>;
    constexpr int R = J * J;
    constexpr int RPL = (R + 31) / 32;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    C* tile = (C*)smem;
    StagePt<T>* stage_all = (StagePt<T>*)(smem + ts.tile_bytes);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wib = tid >> 5;
    StagePt<T>* stage = stage_all + wib * 32;
    const int4 it = items[blockIdx.x];
    const int b = blockIdx.y;
    const int64_t M = g.M;
    int bin = it.x;
    const int o1 = (bin % g.nbin[0]) * g.tile[0];
    bin /= g.nbin[0];
    const int o2 = (bin % g.nbin[1]) * g.tile[1];
    const int o3 = (bin / g.nbin[1]) * g.tile[2];
    const int E1p = ts.E1p, E2 = ts.E2;

    // zero the shared-memory tile
    {
        const int n16 = ts.tile_bytes / 16;
        int4* t4 = (int4*)smem;
        for (int e = tid; e < n16; e += blockDim.x) t4[e] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();

    // this warp's contiguous run of the item's samples (multiples of 32)
    const int per = (((it.z + kTileWarps - 1) / kTileWarps) + 31) & ~31;
    const int begin = it.y + wib * per;
    const int end = min(begin + per, it.y + it.z);
    const C* __restrict__ sb = samples + (int64_t)b * M;

    int rj2[RPL], rj3[RPL];
    bool rvalid[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        const int r = lane + 32 * s;
        rvalid[s] = r < R;
        rj2[s] = (r % R) % J;
        rj3[s] = (r % R) / J;
    }
    const int wax = lane < J ? 0 : (lane < 2 * J ? 1 : 2);
    const int wj = lane - wax * J;
    const bool wactive = lane < 3 * J;
    const T* __restrict__ wh = wax == 0 ? h1 : (wax == 1 ? h2 : h3);
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
    int W1 = 0, W2 = -1, W3 = -1;   // window origin in tile-local cells; W2 < 0: none yet

    for (int base = begin; base < end; base += 32) {
        const int cnt = min(32, end - base);
        __syncwarp();
        if (lane < cnt) {
            const int64_t i = base + lane;
            StagePt<T> p;
            p.t[0] = tm_s[i]; p.t[1] = tm_s[M + i]; p.t[2] = tm_s[2 * M + i];
            p.ko[0] = pt_ko[i]; p.ko[1] = pt_ko[M + i]; p.ko[2] = pt_ko[2 * M + i];
            p.kw[0] = pt_kw[i] - o1; p.kw[1] = pt_kw[M + i] - o2; p.kw[2] = pt_kw[2 * M + i] - o3;
            C f = sb[perm[i]];
            if (phase_s != nullptr) f = cmul_conj(f, phase_s[i]);
            p.fx = f.x; p.fy = f.y;
            stage[lane] = p;
        }
        __syncwarp();
        for (int q = 0; q < cnt; q++) {
            const StagePt<T> cur = stage[q];
            // cooperative weights: lane (axis, tap) evaluates one table coefficient
            T wl = 0;
            if (wactive) {
                const T ta = wax == 0 ? cur.t[0] : (wax == 1 ? cur.t[1] : cur.t[2]);
                const int ka = (wax == 0 ? cur.ko[0] : (wax == 1 ? cur.ko[1] : cur.ko[2])) + wj;
                const T p = (ta - (T)ka) * Lf;
                const T fl = floor(p);
                const T alf = p - fl;
                const int i0 = wnc + (int)fl;
                const int i1 = min(i0 + 1, wtl - 1);
                wl = ((T)1 - alf) * __ldg(wh + i0) + alf * __ldg(wh + i1);
            }
            const int d = cur.kw[0] - W1;
            if (cur.kw[1] != W2 || cur.kw[2] != W3 || d < 0 || d >= J) {
                if (W2 >= 0) {
#pragma unroll
                    for (int j = 0; j < J; j++) {
#pragma unroll
                        for (int s = 0; s < RPL; s++) {
                            if (rvalid[s]) smem_add_c(tile + rowbase[s] + W1 + j, acc[s][j]);
                            acc[s][j] = make_c<T>(0, 0);
                        }
                    }
                }
                W1 = cur.kw[0]; W2 = cur.kw[1]; W3 = cur.kw[2];
#pragma unroll
                for (int s = 0; s < RPL; s++)
                    rowbase[s] = ((W3 + rj3[s]) * E2 + (W2 + rj2[s])) * E1p;
            } else {
                for (int sft = 0; sft < d; sft++) {
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        if (rvalid[s]) smem_add_c(tile + rowbase[s] + W1, acc[s][0]);
#pragma unroll
                        for (int j = 0; j + 1 < J; j++) acc[s][j] = acc[s][j + 1];
                        acc[s][J - 1] = make_c<T>(0, 0);
                    }
                    W1++;
                }
            }
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
        }
    }
    if (W2 >= 0) {
#pragma unroll
        for (int j = 0; j < J; j++)
#pragma unroll
            for (int s = 0; s < RPL; s++)
                if (rvalid[s]) smem_add_c(tile + rowbase[s] + W1 + j, acc[s][j]);
    }
    __syncthreads();

    // one flush of the tile to HBM
    const bool interior = (o1 + ts.E1 <= g.K[0]) && (o2 + ts.E2 <= g.K[1]) && (o3 + ts.E3 <= g.K[2]);
    if (use_tma && interior) {
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_reduce_add_4d(&tmap, tile, 2 * o1, o2, o3, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        C* __restrict__ gb = grid + (int64_t)b * g.PK;
        const int n = ts.E1 * ts.E2 * ts.E3;
        for (int e = tid; e < n; e += blockDim.x) {
            const int i1 = e % ts.E1;
            const int r = e / ts.E1;
            const int i2 = r % ts.E2;
            const int i3 = r / ts.E2;
            const C v = tile[(i3 * ts.E2 + i2) * ts.E1p + i1];
            if (v.x != (T)0 || v.y != (T)0) {
                const int k1 = (o1 + i1) % g.K[0];
                const int k2 = (o2 + i2) % g.K[1];
                const int k3 = (o3 + i3) % g.K[2];
                atomic_add_c(gb + ((int64_t)k3 * g.K[1] + k2) * g.K[0] + k1, v);
            }
        }
    }
}

template <typename T, int J>
static int launch_tile_adj(const Geom& g, const TablePtrs& tabs, const void* tm_s,
                           const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                           const int4* items, int64_t n_items, const void* samples, void* grid,
                           const void* phase_s, int nbatch, int use_tma, cudaStream_t st,
                           bool* done) {
    using C = cplx_t<T>;
    *done = false;
    TileShape ts;
    ts.E1 = g.tile[0] + J - 1;
    ts.E2 = g.tile[1] + J - 1;
    ts.E3 = g.tile[2] + J - 1;
    const int align = 16 / (int)sizeof(C) > 1 ? 16 / (int)sizeof(C) : 1;
    ts.E1p = (ts.E1 + align - 1) / align * align;
    const size_t tb = (size_t)ts.E1p * ts.E2 * ts.E3 * sizeof(C);
    ts.tile_bytes = (int)((tb + 127) / 128 * 128);
    const size_t smem = ts.tile_bytes + (size_t)kTileWarps * 32 * sizeof(StagePt<T>);
    int dev = 0;
    cudaGetDevice(&dev);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_smem) return 0;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const bool tma_ok = use_tma && make_grid_tmap<T, 3>(&map, g, ts, grid, nbatch);
    auto k = spread_tile3d_kernel<T, J>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 gd((unsigned)n_items, (unsigned)nbatch);
    k<<<gd, kTileWarps * 32, smem, st>>>(map, g, ts, tma_ok ? 1 : 0, (const T*)tabs.h[0],
                                         (const T*)tabs.h[1], (const T*)tabs.h[2], (const T*)tm_s,
                                         pt_ko, pt_kw, perm, items, (const C*)samples, (C*)grid,
                                         (const C*)phase_s);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

template <typename T>
static int tile_adj_t(const Geom& g, const TablePtrs& tabs, const void* tm_s, const int32_t* pt_ko,
                      const int32_t* pt_kw, const int32_t* perm, const int4* items, int64_t n_items,
                      const void* samples, void* grid, const void* phase_s, int nbatch, int use_tma,
                      cudaStream_t st, bool* done) {
    *done = false;
    if (g.ndim != 3 || n_items == 0 || n_items > 0x7fffffff || nbatch > 65535) return 0;
    if (g.J[1] != g.J[0] || g.J[2] != g.J[0]) return 0;
#define B2N_TILEADJ(JJ)                                                                          \
    return launch_tile_adj<T, JJ>(g, tabs, tm_s, pt_ko, pt_kw, perm, items, n_items, samples, grid, \
                                  phase_s, nbatch, use_tma, st, done)
    switch (g.J[0]) {
        case 4: B2N_TILEADJ(4);
        case 5: B2N_TILEADJ(5);
        case 6: B2N_TILEADJ(6);
        case 7: B2N_TILEADJ(7);
        case 8: B2N_TILEADJ(8);
        default: return 0;
    }
#undef B2N_TILEADJ
}

}  // namespace b2n
