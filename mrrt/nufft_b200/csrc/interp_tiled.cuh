// Tiled forward interpolation (2-D / 3-D, real table, uniform compile-time J).
//
// One CTA = one work item = up to `chunk` bin-sorted samples of one bin.  The bin's
// grid tile plus its J-1 halo is staged in shared memory -- by one TMA box load
// (cp.async.bulk.tensor, mbarrier completion) when the box does not cross the periodic
// boundary, by cooperative wrapped loads otherwise -- the Kaiser-Bessel lookup table is
// staged next to it, and each thread then gathers its samples' J^d taps from shared
// memory.  Samples inside a bin are sorted by cell (first axis fastest), so the lanes
// of a warp read neighbouring or identical shared-memory words (broadcast).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dispatch.h"

namespace b2n {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (++spins > (1u << 26)) __trap();   // never hang the device on a bad descriptor
    }
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct TileShape {
    int E1, E2, E3;   // tile + halo extents (cells)
    int E1p;          // padded row pitch (cells)
    int tile_bytes;   // E1p*E2*E3*sizeof(C) rounded up to 128
};

// TAB: 0 table in global memory, 1 table staged in shared memory, 2 plan-time weights
// PAIR: the work item ranges over SLOTS (one sample, or two consecutive sorted samples of
//       the same cell, slots[s] = (first sample << 1) | has_partner): both samples of a
//       slot share the window, so every tap is read from shared memory once for the two of
//       them -- the kernel is shared-memory-bandwidth bound and 39 % of the bench
//       trajectory's samples are the second one of such a pair.  PAIR 1 reads the
//       sample-ordered arrays through slots[]; PAIR 2 (plan-time weights only) reads
//       slot-ordered copies (weights of both samples as one vector load), which stay
//       coalesced when the slots of a bin are stored column-interleaved.
// CT:   complex table (phasing="complex"): complex plan-time weights (TAB 2, no pairs);
//       arithmetic of template.c:623-709 / :710-821 (full complex products).
template <typename T, int NDIM, int J, int TAB, int PAIR, bool CT = false>
__global__ void __launch_bounds__(256)
interp_fwd_tiled_kernel(const __grid_constant__ CUtensorMap tmap, Geom g, TileShape ts, int use_tma,
                        const T* __restrict__ tab, const T* __restrict__ wts,
                        const T* __restrict__ tm_s,
                        const int32_t* __restrict__ pt_ko, const int32_t* __restrict__ pt_kw,
                        const int32_t* __restrict__ perm, const int4* __restrict__ items,
                        SlotArgs sa,
                        const cplx_t<T>* __restrict__ grid, cplx_t<T>* __restrict__ out,
                        const cplx_t<T>* __restrict__ phase_s) {
    using C = cplx_t<T>;
    using W = typename WeightT<T, CT>::type;
    static_assert(!CT || (TAB == 2 && PAIR == 0), "complex tables: plan-time weights, no pairs");
    constexpr bool TAB_SMEM = TAB == 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    C* tile = (C*)smem;
    T* stab = (T*)(smem + ts.tile_bytes);
    const int tid = threadIdx.x;
    const int4 it = items[blockIdx.x];
    const int b = blockIdx.y;
    int bin = it.x;
    const int o1 = (bin % g.nbin[0]) * g.tile[0];
    bin /= g.nbin[0];
    const int o2 = NDIM > 1 ? (bin % g.nbin[1]) * g.tile[1] : 0;
    const int o3 = NDIM > 2 ? (bin / g.nbin[1]) * g.tile[2] : 0;
    const bool interior = (o1 + ts.E1 <= g.K[0]) && (NDIM < 2 || o2 + ts.E2 <= g.K[1]) &&
                          (NDIM < 3 || o3 + ts.E3 <= g.K[2]);
    const bool tma = use_tma && interior;
    if (tma) {
        if (tid == 0) mbar_init(&mbar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&mbar, (uint32_t)(ts.E1p * ts.E2 * ts.E3 * (int)sizeof(C)));
            if (NDIM == 2) tma_load_3d(tile, &tmap, &mbar, 2 * o1, o2, b);
            else tma_load_4d(tile, &tmap, &mbar, 2 * o1, o2, o3, b);
        }
    } else {
        const C* __restrict__ gb = grid + (int64_t)b * g.PK;
        const int n = ts.E1 * ts.E2 * ts.E3;
        for (int e = tid; e < n; e += blockDim.x) {
            const int i1 = e % ts.E1;
            const int r = e / ts.E1;
            const int i2 = r % ts.E2;
            const int i3 = r / ts.E2;
            int k1 = o1 + i1; if (k1 >= g.K[0]) k1 -= g.K[0];
            int k2 = o2 + i2; if (k2 >= g.K[1]) k2 -= g.K[1];
            int k3 = o3 + i3; if (k3 >= g.K[2]) k3 -= g.K[2];
            // K may be smaller than tile+halo on tiny grids
            k1 %= g.K[0]; k2 %= g.K[1]; k3 %= g.K[2];
            tile[(i3 * ts.E2 + i2) * ts.E1p + i1] = __ldg(gb + ((int64_t)k3 * g.K[1] + k2) * g.K[0] + k1);
        }
    }
    if (TAB_SMEM) {
        for (int e = tid; e < g.tlen[0]; e += blockDim.x) stab[e] = __ldg(tab + e);
    }
    __syncthreads();
    if (tma) mbar_wait(&mbar, 0);
    const T* __restrict__ h = TAB_SMEM ? stab : tab;
    const int64_t M = g.M;
    const int end = it.y + it.z;
    const uint32_t* __restrict__ slots = sa.slots;
    const int64_t ns = sa.ns;
    using T2 = typename Cplx<T>::type;      // (weight, partner's weight)
    const T2* __restrict__ wts2 = (const T2*)sa.wts2;
    for (int sidx = it.y + tid; sidx < end; sidx += blockDim.x) {
        int i = sidx;
        bool pair = false;
        if (PAIR == 1) {
            const uint32_t u = slots[sidx];
            i = (int)(u >> 1);
            pair = (u & 1u) != 0;
        }
        // output positions first: their latency hides behind the gather
        int pa = 0, pb = -1;
        if (PAIR == 2) {
            pa = __ldg(sa.perm + sidx);
            pb = __ldg(sa.perm + ns + sidx);
        }
        W w[NDIM][J];
        T wq[PAIR ? NDIM : 1][PAIR ? J : 1];   // partner's weights (zero when there is none)
        int c[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            const int od = d == 0 ? o1 : (d == 1 ? o2 : o3);
            if constexpr (PAIR == 2) {
                c[d] = __ldg(sa.kw + (int64_t)d * ns + sidx) - od;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const T2 ww = wts2[(int64_t)(d * J + j) * ns + sidx];
                    w[d][j] = ww.x;
                    wq[d][j] = ww.y;
                }
                continue;
            }
            c[d] = pt_kw[(int64_t)d * M + i] - od;           // wrapped origin inside the tile
            if constexpr (CT) {
#pragma unroll
                for (int j = 0; j < J; j++) w[d][j] = ((const W*)wts)[(int64_t)(d * J + j) * M + i];
            } else if (TAB == 2) {
#pragma unroll
                for (int j = 0; j < J; j++) w[d][j] = wts[(int64_t)(d * J + j) * M + i];
                if (PAIR) {
#pragma unroll
                    for (int j = 0; j < J; j++)
                        wq[d][j] = pair ? wts[(int64_t)(d * J + j) * M + i + 1] : (T)0;
                }
            } else if constexpr (!CT) {
                const T t = tm_s[(int64_t)d * M + i];
                const int koff = pt_ko[(int64_t)d * M + i];  // 1 + floor(t - J/2.), plan time
#pragma unroll
                for (int j = 0; j < J; j++)
                    w[d][j] = tap_real<T>(h, g.ncenter[0], g.tlen[0], t, koff + j, g.L, g.order);
                if (PAIR) {
                    const T tq = tm_s[(int64_t)d * M + (pair ? i + 1 : i)];
                    // same wrapped cell, but possibly whole periods away: its own origin
                    const int koq = pt_ko[(int64_t)d * M + (pair ? i + 1 : i)];
#pragma unroll
                    for (int j = 0; j < J; j++)
                        wq[d][j] = pair ? tap_real<T>(h, g.ncenter[0], g.tlen[0], tq, koq + j, g.L, g.order)
                                        : (T)0;
                }
            }
        }
        C s3 = make_c<T>(0, 0), q3 = make_c<T>(0, 0);
#pragma unroll
        for (int j3 = 0; j3 < (NDIM > 2 ? J : 1); j3++) {
            C s2 = make_c<T>(0, 0), q2 = make_c<T>(0, 0);
#pragma unroll
            for (int j2 = 0; j2 < (NDIM > 1 ? J : 1); j2++) {
                int row = c[0];
                if (NDIM == 2) row += (c[1] + j2) * ts.E1p;
                if (NDIM == 3) row += ((c[NDIM > 2 ? 2 : 0] + j3) * ts.E2 + (c[1] + j2)) * ts.E1p;
                const C* __restrict__ pr = tile + row;
                C s1 = make_c<T>(0, 0), q1 = make_c<T>(0, 0);
#pragma unroll
                for (int j1 = 0; j1 < J; j1++) {
                    const C v = pr[j1];
                    s1 = wfma(w[0][j1], v, s1);                        // FFMA2 (real weights)
                    if (PAIR) q1 = fma_w(wq[0][j1], v, q1);
                }
                s2 = wfma(w[NDIM > 1 ? 1 : 0][j2], s1, s2);
                if (PAIR) q2 = fma_w(wq[NDIM > 1 ? 1 : 0][j2], q1, q2);
            }
            if (NDIM > 2) {
                s3 = wfma(w[NDIM > 2 ? 2 : 0][j3], s2, s3);
                if (PAIR) q3 = fma_w(wq[NDIM > 2 ? 2 : 0][j3], q2, q3);
            } else {
                s3 = s2;
                q3 = q2;
            }
        }
        if (PAIR == 2) {
            if (sa.phase2 != nullptr) {
                const C* __restrict__ ph2 = (const C*)sa.phase2;
                s3 = cmul(s3, ph2[2 * (int64_t)sidx]);
                q3 = cmul(q3, ph2[2 * (int64_t)sidx + 1]);
            }
            out[(int64_t)b * M + pa] = s3;
            if (pb >= 0) out[(int64_t)b * M + pb] = q3;
            continue;
        }
        if (phase_s != nullptr) s3 = cmul(s3, phase_s[i]);
        out[(int64_t)b * M + perm[i]] = s3;
        if (PAIR && pair) {
            if (phase_s != nullptr) q3 = cmul(q3, phase_s[i + 1]);
            out[(int64_t)b * M + perm[i + 1]] = q3;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// tensor map over the grid viewed as REAL elements: dims (2*K1, K2[, K3], nbatch)
template <typename T, int NDIM>
static bool make_grid_tmap(CUtensorMap* map, const Geom& g, const TileShape& ts, const void* grid,
                           int nbatch) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) return false;
    const size_t cs = 2 * sizeof(T);
    if ((g.K[0] * cs) % 16 != 0 || ((uintptr_t)grid % 16) != 0) return false;
    if ((ts.E1p * cs) % 16 != 0 || 2 * ts.E1p > 256 || ts.E2 > 256 || ts.E3 > 256) return false;
    cuuint64_t dims[4];
    cuuint64_t strides[3];
    cuuint32_t box[4];
    cuuint32_t estr[4] = {1, 1, 1, 1};
    dims[0] = 2 * (cuuint64_t)g.K[0];
    box[0] = 2 * ts.E1p;
    dims[1] = g.K[1];
    box[1] = ts.E2;
    strides[0] = g.K[0] * cs;
    if (NDIM == 2) {
        dims[2] = nbatch;
        box[2] = 1;
        strides[1] = (cuuint64_t)g.PK * cs;
    } else {
        dims[2] = g.K[2];
        box[2] = ts.E3;
        strides[1] = (cuuint64_t)g.K[0] * g.K[1] * cs;
        dims[3] = nbatch;
        box[3] = 1;
        strides[2] = (cuuint64_t)g.PK * cs;
    }
    CUresult r = enc(map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                     NDIM + 1, const_cast<void*>(grid), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <typename T, int NDIM, int J>
static int launch_fwd_tiled(const Geom& g, bool cplx, const TablePtrs& tabs, const void* tm_s,
                            const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm, const int4* items, int64_t n_items,
                            const SlotArgs& sa, const void* grid, void* out, const void* phase_s, int nbatch,
                            const FwdOpts& fo, cudaStream_t st, bool* done) {
    using C = cplx_t<T>;
    *done = false;
    const int fwd_pitch = fo.pitch;
    const int use_tma = fo.use_tma;
    TileShape ts;
    ts.E1 = g.tile[0] + J - 1;
    ts.E2 = NDIM > 1 ? g.tile[1] + J - 1 : 1;
    ts.E3 = NDIM > 2 ? g.tile[2] + J - 1 : 1;
    // TMA box rows must be a multiple of 16 bytes.  A row pitch of a whole number of
    // 128-byte bank rows makes the bank depend on the axis-1 position only, which is what
    // the cell-sorted (axis 1 fastest) lanes of a warp differ in: measured/simulated
    // bank-conflict wavefronts drop from +31 % (pitch 22) to +18 % (pitch 32) on the
    // bench trajectory (profiles/r01_notes.md).
    const int align = 16 / (int)sizeof(C) > 1 ? 16 / (int)sizeof(C) : 1;
    ts.E1p = (ts.E1 + align - 1) / align * align;
    if (fwd_pitch > 0 && fwd_pitch >= ts.E1 && fwd_pitch % align == 0) ts.E1p = fwd_pitch;
    else if (fwd_pitch == 0 && sizeof(C) == 8 && ts.E1 <= 32 && ts.E1 > 16) ts.E1p = 32;
    const size_t tb = (size_t)ts.E1p * ts.E2 * ts.E3 * sizeof(C);
    ts.tile_bytes = (int)((tb + 127) / 128 * 128);
    const size_t tab_bytes = (size_t)g.tlen[0] * sizeof(T);
    int dev = 0;
    cudaGetDevice(&dev);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    bool tab_smem = wts == nullptr && !sa.packed;
    size_t smem = ts.tile_bytes + (tab_smem ? tab_bytes : 0);
    if (smem > (size_t)max_smem || smem > 100 * 1024) {
        tab_smem = false;
        smem = ts.tile_bytes;
    }
    if (smem > (size_t)max_smem) return 0;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const bool tma_ok = use_tma && make_grid_tmap<T, NDIM>(&map, g, ts, grid, nbatch);
    dim3 gridDim((unsigned)n_items, (unsigned)nbatch);
    cudaError_t e;
#define B2N_LAUNCH_FWD_P(TABV, PAIRV)                                                              \
    {                                                                                              \
        auto k = interp_fwd_tiled_kernel<T, NDIM, J, TABV, PAIRV>;                                 \
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (e != cudaSuccess) return (int)e;                                                       \
        k<<<gridDim, 256, smem, st>>>(map, g, ts, tma_ok ? 1 : 0, (const T*)tabs.h[0],             \
                                      (const T*)wts, (const T*)tm_s, pt_ko, pt_kw, perm, items,    \
                                      sa, (const C*)grid, (C*)out, (const C*)phase_s);             \
    }
    if (cplx) {
        // complex table: plan-time complex weights only
        if (wts == nullptr) return 0;
        auto k = interp_fwd_tiled_kernel<T, NDIM, J, 2, 0, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<gridDim, 256, smem, st>>>(map, g, ts, tma_ok ? 1 : 0, (const T*)tabs.h[0], (const T*)wts,
                                      (const T*)tm_s, pt_ko, pt_kw, perm, items, sa, (const C*)grid,
                                      (C*)out, (const C*)phase_s);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        *done = true;
        return 0;
    }
    const int pairv = sa.slots == nullptr ? 0 : (sa.packed ? 2 : 1);
#define B2N_LAUNCH_FWD(TABV)                                                                       \
    {                                                                                              \
        if (pairv) B2N_LAUNCH_FWD_P(TABV, 1) else B2N_LAUNCH_FWD_P(TABV, 0)                        \
    }
    if (pairv == 2) B2N_LAUNCH_FWD_P(2, 2)
    else if (wts != nullptr) B2N_LAUNCH_FWD(2)
    else if (tab_smem) B2N_LAUNCH_FWD(1)
    else B2N_LAUNCH_FWD(0)
#undef B2N_LAUNCH_FWD
#undef B2N_LAUNCH_FWD_P
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

// returns 0 or a cudaError_t; *done tells whether the tiled kernel took the call
template <typename T>
static int tiled_fwd_t(const Geom& g, int Jk, bool cplx, bool tables_equal, const TablePtrs& tabs, const void* tm_s,
                       const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm, const int4* items, int64_t n_items, const SlotArgs& sa, const void* grid,
                       void* out, const void* phase_s, int nbatch, const FwdOpts& fo, cudaStream_t st,
                       bool* done) {
    *done = false;
    if (g.ndim < 2 || (!tables_equal && wts == nullptr && !sa.packed) || n_items == 0 || n_items > 0x7fffffff ||
        nbatch > 65535)
        return 0;
    // a window wider than an axis' own J needs the zero-padded plan-time weights
    for (int d = 0; d < g.ndim; d++) {
        if (g.J[d] != Jk && wts == nullptr && !sa.packed) return 0;
        if (g.K[d] < Jk) return 0;
    }
#define B2N_TILED(ND, JJ)                                                                      \
    return launch_fwd_tiled<T, ND, JJ>(g, cplx, tabs, tm_s, wts, pt_ko, pt_kw, perm, items, n_items, sa, grid, out, phase_s, \
                                       nbatch, fo, st, done)
    if (g.ndim == 2) {
        switch (Jk) {
            case 4: B2N_TILED(2, 4);
            case 6: B2N_TILED(2, 6);
            case 8: B2N_TILED(2, 8);
            default: return 0;
        }
    }
    switch (Jk) {
        case 4: B2N_TILED(3, 4);
        case 6: B2N_TILED(3, 6);
        case 8: B2N_TILED(3, 8);
        default: return 0;
    }
#undef B2N_TILED
}

}  // namespace b2n
