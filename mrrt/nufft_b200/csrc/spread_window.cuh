// Register-window adjoint gridding (3-D, real table, uniform J).
//
// Each half-warp walks a contiguous run of cell-sorted samples and keeps the current
// sample's J x J x J window of partial sums in registers: lane <-> one (b, c) position of
// the window face, J accumulators along the SLIDE axis a.  Consecutive samples of the
// sorted order sit in the same or the next cell along a, so the window slides and only the
// retiring J x J face is sent to L2 (vector REDs).  Axis roles are a run-time
// permutation: with the adjoint sort order (cells ordered axis-3-fastest inside a bin)
// a = axis 3 and the lane face is (axis 1, axis 2), so that the lanes of one retiring
// face write runs of J CONSECUTIVE grid cells -- 2-3 sectors per run instead of one
// sector per lane (profiles/r01_notes.md: strided REDs were the L1TEX bottleneck).
//
// Per 16-sample batch the weights are prepared LANE-PARALLEL (lane = sample) together
// with a window action code (slide distance or "new window") and written to a per-warp
// staging record; the per-sample loop then only does broadcast shared-memory reads and
// packed FMAs.  The last cell of a slide is folded into the FMAs of the sliding sample
// (destination j, addend j + 1): no register moves.
//
// Arithmetic per sample follows c/nufft_table.template.c:1122-1163:
// v3 = coef3*f, v2 = coef2*v3, ck += coef1*v2 (the same products, grouped by axis role).
//
// Variants measured and dropped in round 2 (lane-ring and fixed-ring windows, 8 / 32 lanes
// per sample, plan-time window records with and without cp.async, the shared-memory tile
// adjoint and the first sliding-window generation) are described in profiles/r01_notes.md;
// their code is in the history up to commit 4d956cd.
#pragma once
#include "common.cuh"
#include "dispatch.h"

namespace b2n {

struct WindowAxes {
    int ax[3];          // ax[0] = slide axis a, ax[1] = fast lane axis b, ax[2] = slow lane axis c
    int K[3];           // grid size along (a, b, c)
    int stride[3];      // grid stride (cells) along (a, b, c)
};

constexpr int kWinLanes = 16;   // lanes per sample: two register windows per warp

// Staging of one batch (16 samples per half-warp): per sample a record of kW values (3J
// weights, fx, fy) with an ODD pitch in elements, so that the lanes' scalar stores (lane =
// sample) fall in different banks, plus a separate int4 (kA, kB, kC, action) per sample.
template <typename T, int J, bool FW = false, bool CT = false> struct WinRec {
    static constexpr int kW = (CT ? 6 : 3) * J + 2;              // weights (complex: re, im) + fx, fy
    // odd pitch in units of sizeof(T): float records spread over all 32 banks, double
    // records over all 16 bank pairs, and every element stays naturally aligned
    static constexpr int kPitchElems = kW % 2 == 1 ? kW : kW + 1;
    static constexpr int kPitch = kPitchElems * (int)sizeof(T);  // bytes between records
    // per-warp staging: 32 records (padded to 16 bytes) + 32 action codes
    static constexpr int kRecBytes = ((32 * kPitch + 15) / 16) * 16;
    static constexpr int kBytes = kRecBytes + 32 * 16;           // per warp
};

// FW (face weights) staging.  Per sample a 16-byte aligned HEAD = {J slide-axis weights, fx,
// fy, (kA, kB, kC, action)} written by the batch phase with 16-byte vector stores and read
// by the sample loop with 16-byte broadcast loads (3 for float J=6 instead of 9 scalar
// ones), and a FACE record of the J*J products wb[jb]*wc[jc] in face order r = jb + J*jc,
// so a lane fetches its face weights from ONE lane-dependent base (+16 per slot) instead
// of two table-indexed reads and one multiply per slot.  Head pitch = 16 bytes x odd: the
// 8 lanes of a quarter-warp store phase then cover all 32 banks exactly once.
template <typename T, int J> struct WinRec<T, J, true, false> {
    static constexpr int kHeadInt = (((J + 2) * (int)sizeof(T) + 15) / 16) * 16;   // offset of the int4
    static constexpr int kHeadMin = kHeadInt + 16;
    static constexpr int kHead = (kHeadMin / 16) % 2 == 1 ? kHeadMin : kHeadMin + 16;
    static constexpr int kFaceElems = (J * J) % 2 == 1 ? J * J : J * J + 1;          // odd pitch
    // slots past the face (r >= J*J) read up to 31 elements past a record: keep them inside
    static constexpr int kFaceBytes = ((32 * kFaceElems + 32) * (int)sizeof(T) + 15) / 16 * 16;
    static constexpr int kBytes = 32 * kHead + kFaceBytes;       // per warp
};

__device__ __forceinline__ void store16(void* dst, const float* v) {
    *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store16(void* dst, const double* v) {
    *(double2*)dst = make_double2(v[0], v[1]);
}
__device__ __forceinline__ void load16(const void* src, float* v) {
    const float4 t = *(const float4*)src;
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load16(const void* src, double* v) {
    const double2 t = *(const double2*)src;
    v[0] = t.x; v[1] = t.y;
}

// conj(coef_b) * (conj(coef_c) * f): the face value every slide-axis tap is multiplied with
// (face-weight staging: wb already holds the product coef_b * coef_c)
template <typename T, bool FW, bool CT, typename WT>
__device__ __forceinline__ cplx_t<T> face_value(WT wb, WT wc, T fx, T fy) {
    if constexpr (CT) return w_mul_conj(wb, w_mul_conj(wc, make_c<T>(fx, fy)));
    else if constexpr (FW) return mul_w(wb, make_c<T>(fx, fy));
    else return mul_w(wb, mul_w(wc, make_c<T>(fx, fy)));
}

// TAB: 0 table in global memory, 1 table staged in shared memory, 2 plan-time weights.
// FWV: 0 = scalar staging records; 1 = face-weight staging (plan-time weights only);
//      2 = the same compiled for 5 CTAs per SM (float J<=6: 96 registers with a 48-byte
//      spill instead of 110).
// CT:  complex table (phasing="complex"): complex plan-time weights in the scalar records,
//      conjugated products (template.c:1021-1022); TAB 2, FWV 0 only.
template <typename T, int J, int TAB, int FWV = 0, bool CT = false>
__global__ void __launch_bounds__(128, FWV == 2 ? 5 : 1)
spread_window3d_kernel(Geom g, WindowAxes wa, const T* __restrict__ h1, const T* __restrict__ h2,
                       const T* __restrict__ h3, const T* __restrict__ tm_s,
                       const T* __restrict__ wts,
                       const int32_t* __restrict__ pt_ko, const int32_t* __restrict__ pt_kw,
                       const int32_t* __restrict__ perm, const cplx_t<T>* __restrict__ samples,
                       cplx_t<T>* __restrict__ grid, const cplx_t<T>* __restrict__ phase_s,
                       int pts_per_warp, int max_slide) {
    using C = cplx_t<T>;
    constexpr int G = kWinLanes;
    constexpr bool FW = FWV != 0;
    constexpr int R = J * J;
    constexpr int RPL = (R + G - 1) / G;
    constexpr int NG = 32 / G;                                    // sample groups per warp
    static_assert(!FW || TAB == 2, "face-weight staging needs the plan-time weights");
    static_assert(!CT || (TAB == 2 && !FW), "complex tables: plan-time weights, scalar records");
    using W = typename WeightT<T, CT>::type;
    constexpr int RB = WinRec<T, J, false, CT>::kPitch;           // (not used with FW)
    constexpr int WB = FW ? WinRec<T, J, true, false>::kBytes : WinRec<T, J, false, CT>::kBytes;
    constexpr int HP = WinRec<T, J, true>::kHead;                 // FW: head pitch (bytes)
    constexpr int HI = WinRec<T, J, true>::kHeadInt;              // FW: offset of (kA, kB, kC, act)
    constexpr int FP = WinRec<T, J, true>::kFaceElems;            // FW: face pitch (elements)
    constexpr int HV = HI / (int)sizeof(T);                       // FW: head values incl. padding
    constexpr int VPC = 16 / (int)sizeof(T);                      // values per 16-byte chunk
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const T* stab = (const T*)(dyn_smem + 4 * WB);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char* stage = dyn_smem + wib * WB;                   // this warp's records
    int4* actions = (int4*)(stage + WinRec<T, J, false, CT>::kRecBytes);  // this warp's action codes (not FW)
    T* face = (T*)(stage + 32 * HP);                              // FW: this warp's face records
    const int64_t M = g.M;
    constexpr bool TAB_SMEM = TAB == 1;
    if (TAB_SMEM) {
        T* st = (T*)(dyn_smem + 4 * WB);
        for (int e = threadIdx.x; e < g.tlen[0]; e += blockDim.x) st[e] = __ldg(h1 + e);
        __syncthreads();
    }
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (warp * pts_per_warp >= M) return;
    // this lane group's contiguous run of samples
    const int grp = lane / G;
    const int lg = lane - grp * G;
    const int per_group = pts_per_warp / NG;
    const int64_t begin = warp * pts_per_warp + (int64_t)grp * per_group;
    const int64_t end = begin + per_group < M ? begin + per_group : (begin < M ? M : begin);
    const int b = blockIdx.y;
    const C* __restrict__ sb = samples + (int64_t)b * M;
    C* __restrict__ gb = grid + (int64_t)b * g.PK;
    const int aA = wa.ax[0], aB = wa.ax[1], aC = wa.ax[2];
    const int KA = wa.K[0], KB = wa.K[1], KC = wa.K[2];
    const int sA = wa.stride[0], sB = wa.stride[1], sC = wa.stride[2];
    const T* __restrict__ tabA = TAB_SMEM ? stab : (aA == 0 ? h1 : (aA == 1 ? h2 : h3));
    const T* __restrict__ tabB = TAB_SMEM ? stab : (aB == 0 ? h1 : (aB == 1 ? h2 : h3));
    const T* __restrict__ tabC = TAB_SMEM ? stab : (aC == 0 ? h1 : (aC == 1 ? h2 : h3));
    const int ncA = aA == 0 ? g.ncenter[0] : (aA == 1 ? g.ncenter[1] : g.ncenter[2]);
    const int ncB = aB == 0 ? g.ncenter[0] : (aB == 1 ? g.ncenter[1] : g.ncenter[2]);
    const int ncC = aC == 0 ? g.ncenter[0] : (aC == 1 ? g.ncenter[1] : g.ncenter[2]);
    const int tlA = aA == 0 ? g.tlen[0] : (aA == 1 ? g.tlen[1] : g.tlen[2]);
    const int tlB = aB == 0 ? g.tlen[0] : (aB == 1 ? g.tlen[1] : g.tlen[2]);
    const int tlC = aC == 0 ? g.tlen[0] : (aC == 1 ? g.tlen[1] : g.tlen[2]);

    // this lane's positions on the window face: r = jb + J*jc
    int rjb[RPL], rjc[RPL];
    bool rvalid[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        const int r = lg + G * s;
        rvalid[s] = r < R;
        rjb[s] = (r % R) % J;
        rjc[s] = (r % R) / J;
    }
    C acc[RPL][J];
    C* faceptr[RPL];        // grid address of this lane's face position (axes b, c)
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        faceptr[s] = gb;   // (overwritten by the first window)
#pragma unroll
        for (int j = 0; j < J; j++) acc[s][j] = make_c<T>(0, 0);
    }
    int WA = 0;             // wrapped window origin along the slide axis
    bool have = false;
    int pkA = -1 << 30, pkB = -1, pkC = -1;   // previous sample's wrapped origin

    for (int it = 0; it < per_group; it += G) {
        // (uniform trip count for the whole warp; groups past their end idle)
        const int64_t base = begin + it;
        const int cnt = (int)(base >= end ? 0 : (end - base < G ? end - base : G));
        __syncwarp();
        // ---- batch phase: lane = sample (record index = lane)
        int kA = 0, kB = 0, kC = 0;
        if (lg < cnt) {
            const int64_t i = base + lg;
            T* w = (T*)(stage + lane * RB);
            kA = pt_kw[(int64_t)aA * M + i];
            kB = pt_kw[(int64_t)aB * M + i];
            kC = pt_kw[(int64_t)aC * M + i];
            if constexpr (FW) {
                T hv[HV], wB[J], wC[J];
#pragma unroll
                for (int e = 0; e < HV; e++) hv[e] = (T)0;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    hv[j] = wts[(int64_t)(aA * J + j) * M + i];
                    wB[j] = wts[(int64_t)(aB * J + j) * M + i];
                    wC[j] = wts[(int64_t)(aC * J + j) * M + i];
                }
                C f = sb[perm[i]];
                if (phase_s != nullptr) f = cmul_conj(f, phase_s[i]);
                hv[J] = f.x;
                hv[J + 1] = f.y;
                unsigned char* hb = stage + lane * HP;
#pragma unroll
                for (int c = 0; c < HV / VPC; c++) store16(hb + 16 * c, hv + VPC * c);
                T* fr = face + lane * FP;
#pragma unroll
                for (int jc = 0; jc < J; jc++)
#pragma unroll
                    for (int jb = 0; jb < J; jb++) fr[jb + J * jc] = wB[jb] * wC[jc];
            } else if constexpr (CT) {
                const W* __restrict__ wc_ = (const W*)wts;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const W a = wc_[(int64_t)(aA * J + j) * M + i], bq = wc_[(int64_t)(aB * J + j) * M + i],
                            cq = wc_[(int64_t)(aC * J + j) * M + i];
                    w[2 * j] = a.x; w[2 * j + 1] = a.y;
                    w[2 * (J + j)] = bq.x; w[2 * (J + j) + 1] = bq.y;
                    w[2 * (2 * J + j)] = cq.x; w[2 * (2 * J + j) + 1] = cq.y;
                }
            } else if (TAB == 2) {
#pragma unroll
                for (int j = 0; j < J; j++) {
                    w[j] = wts[(int64_t)(aA * J + j) * M + i];
                    w[J + j] = wts[(int64_t)(aB * J + j) * M + i];
                    w[2 * J + j] = wts[(int64_t)(aC * J + j) * M + i];
                }
            } else if constexpr (!CT) {
                const T tA = tm_s[(int64_t)aA * M + i], tB = tm_s[(int64_t)aB * M + i],
                        tC = tm_s[(int64_t)aC * M + i];
                const int oA = pt_ko[(int64_t)aA * M + i], oB = pt_ko[(int64_t)aB * M + i],
                          oC = pt_ko[(int64_t)aC * M + i];
#pragma unroll
                for (int j = 0; j < J; j++) {
                    w[j] = tap_real<T>(tabA, ncA, tlA, tA, oA + j, g.L, g.order);
                    w[J + j] = tap_real<T>(tabB, ncB, tlB, tB, oB + j, g.L, g.order);
                    w[2 * J + j] = tap_real<T>(tabC, ncC, tlC, tC, oC + j, g.L, g.order);
                }
            }
            if constexpr (!FW) {
                C f = sb[perm[i]];
                if (phase_s != nullptr) f = cmul_conj(f, phase_s[i]);
                w[(CT ? 6 : 3) * J] = f.x;
                w[(CT ? 6 : 3) * J + 1] = f.y;
            }
        }
        // window action: slide distance along a, or -1 = new window
        {
            int qA = __shfl_up_sync(FULL, kA, 1, G), qB = __shfl_up_sync(FULL, kB, 1, G),
                qC = __shfl_up_sync(FULL, kC, 1, G);
            if (lg == 0) { qA = pkA; qB = pkB; qC = pkC; }
            const int d = kA - qA;
            // slides longer than max_slide cells cost more instructions than flushing the
            // whole window and starting a new one
            const int act = (kB == qB && kC == qC && d >= 0 && d <= max_slide) ? d : -1;
            if (lg < cnt) {
                if constexpr (FW) *(int4*)(stage + lane * HP + HI) = make_int4(kA, kB, kC, act);
                else actions[lane] = make_int4(kA, kB, kC, act);
            }
            const int last = cnt > 0 ? cnt - 1 : 0;
            const int nA = __shfl_sync(FULL, kA, last, G), nB = __shfl_sync(FULL, kB, last, G),
                      nC = __shfl_sync(FULL, kC, last, G);
            if (cnt > 0) { pkA = nA; pkB = nB; pkC = nC; }
        }
        __syncwarp();
        // ---- sample loop: all lanes of a group work on one sample
        // FW: only the action code is prefetched (one register); the origin (kA, kB, kC) is read
        // by the new-window path itself, and the record pointers advance by a constant
        int4 kk_next = make_int4(0, 0, 0, 0);
        const unsigned char* recf = stage + (grp * G) * HP;
        const T* wff = face + (grp * G) * FP + lg;
        if constexpr (FW) kk_next.w = *(const int*)(recf + HI + 12);
        else kk_next = actions[grp * G];
        for (int q = 0; q < cnt; q++, recf += HP, wff += FP) {
            const unsigned char* rec = FW ? recf : stage + (grp * G + q) * RB;
            const int4 kk = kk_next;
            if (q + 1 < cnt) {
                if constexpr (FW) kk_next.w = *(const int*)(rec + HP + HI + 12);
                else kk_next = actions[grp * G + q + 1];
            }
            // operands of this sample are fetched before the window update so that their
            // shared-memory latency overlaps it
            const T* w = (const T*)rec;
            W wA[J];
            T fx, fy;
            W wb[RPL], wc[RPL];
            if constexpr (CT) {
#pragma unroll
                for (int j = 0; j < J; j++) wA[j] = make_c<T>(w[2 * j], w[2 * j + 1]);
                fx = w[6 * J];
                fy = w[6 * J + 1];
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    wb[s] = make_c<T>(w[2 * (J + rjb[s])], w[2 * (J + rjb[s]) + 1]);
                    wc[s] = make_c<T>(w[2 * (2 * J + rjc[s])], w[2 * (2 * J + rjc[s]) + 1]);
                }
            } else if constexpr (FW) {
                T hv[HV];
#pragma unroll
                for (int c = 0; c < HV / VPC; c++) load16(rec + 16 * c, hv + VPC * c);
#pragma unroll
                for (int j = 0; j < J; j++) wA[j] = hv[j];
                fx = hv[J];
                fy = hv[J + 1];
                // one lane-dependent base, slots at a fixed distance (r = lg + G*s).  Slots past
                // the face (r >= J*J) read up to G*RPL - J*J elements past the record: still
                // inside this warp's face area (next record / tail padding); never flushed.
                const T* wf = wff;
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    wb[s] = wf[G * s];
                    wc[s] = wb[s];
                }
            } else {
#pragma unroll
                for (int j = 0; j < J; j++) wA[j] = w[j];
                fx = w[3 * J];
                fy = w[3 * J + 1];
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    wb[s] = w[J + rjb[s]];
                    wc[s] = w[2 * J + rjc[s]];
                }
            }
            if (FW && kk.w == 0) {
                // same cell as the previous sample (the common case in the dense centre):
                // straight to the FMAs
            } else
            if (kk.w < 0) {
                if (have) {
#pragma unroll
                    for (int j = 0; j < J; j++) {
                        int ka = WA + j;
                        if (ka >= KA) ka -= KA;
#pragma unroll
                        for (int s = 0; s < RPL; s++) {
                            if (rvalid[s]) atomic_add_c(faceptr[s] + (int64_t)ka * sA, acc[s][j]);
                            acc[s][j] = make_c<T>(0, 0);
                        }
                    }
                }
                have = true;
                const int4 ko = FW ? *(const int4*)(rec + HI) : kk;     // this sample's origin
                WA = ko.x;
#pragma unroll
                for (int s = 0; s < RPL; s++) {
                    int kb = ko.y + rjb[s]; if (kb >= KB) kb -= KB;
                    int kc = ko.z + rjc[s]; if (kc >= KC) kc -= KC;
                    faceptr[s] = gb + ((int64_t)kb * sB + (int64_t)kc * sC);
                }
            } else {
#pragma unroll 1
                for (int sft = 0; sft < kk.w - 1; sft++) {
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        if (rvalid[s]) atomic_add_c(faceptr[s] + (int64_t)WA * sA, acc[s][0]);
#pragma unroll
                        for (int j = 0; j + 1 < J; j++) acc[s][j] = acc[s][j + 1];
                        acc[s][J - 1] = make_c<T>(0, 0);
                    }
                    WA++;   // stays < KA: it ends at this sample's wrapped origin
                }
                if (kk.w > 0) {
                    // last cell of the slide: flush register 0 and let the FMAs themselves
                    // do the shift (destination j, addend j + 1) -- no register moves.
                    // (Fusing slides by 2..J-1 cells the same way was measured: 124
                    // registers and 5.25 ms instead of 4.79 ms: more divergence between the two
                    // half-warps and one CTA per SM fewer.)
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        if (rvalid[s]) {
                            atomic_add_c(faceptr[s] + (int64_t)WA * sA, acc[s][0]);
                            const C v2 = face_value<T, FW, CT>(wb[s], wc[s], fx, fy);
#pragma unroll
                            for (int j = 0; j + 1 < J; j++)
                                acc[s][j] = wfma_conj(wA[j], v2, acc[s][j + 1]);
                            acc[s][J - 1] = wfma_conj(wA[J - 1], v2, make_c<T>(0, 0));
                        }
                    }
                    WA++;
                    continue;
                }
            }
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                if (rvalid[s]) {
                    // (coef_c * f) * coef_b, then * coef_a per cell; packed re/im arithmetic
                    const C v2 = face_value<T, FW, CT>(wb[s], wc[s], fx, fy);
#pragma unroll
                    for (int j = 0; j < J; j++) acc[s][j] = wfma_conj(wA[j], v2, acc[s][j]);
                }
            }
        }
    }
    if (have) {
#pragma unroll
        for (int j = 0; j < J; j++) {
            int ka = WA + j;
            if (ka >= KA) ka -= KA;
#pragma unroll
            for (int s = 0; s < RPL; s++)
                if (rvalid[s]) atomic_add_c(faceptr[s] + (int64_t)ka * sA, acc[s][j]);
        }
    }
}

template <typename T, int J>
static int launch_window(const Geom& g, bool cplx, const TablePtrs& tabs, const WindowOpts& wo, const void* tm_s,
                         const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                         const void* samples, void* grid, const void* phase_s, int nbatch,
                         cudaStream_t st, bool* done) {
    int max_slide = wo.max_slide;
    if (max_slide <= 0 || max_slide > J - 1) max_slide = J - 1;
    const int pts_per_warp = (wo.pts_per_warp + 31) / 32 * 32;
    using C = cplx_t<T>;
    const int64_t nwarps = (g.M + pts_per_warp - 1) / pts_per_warp;
    const int64_t nblocks = (nwarps + 3) / 4;
    if (nblocks > 0x7fffffff || nbatch > 65535) return 0;
    WindowAxes wa;
    if (wo.slide_axis == 2) { wa.ax[0] = 2; wa.ax[1] = 0; wa.ax[2] = 1; }
    else { wa.ax[0] = 0; wa.ax[1] = 1; wa.ax[2] = 2; }
    const int strides[3] = {1, g.K[0], g.K[0] * g.K[1]};
    for (int r = 0; r < 3; r++) { wa.K[r] = g.K[wa.ax[r]]; wa.stride[r] = strides[wa.ax[r]]; }
    dim3 gd((unsigned)nblocks, (unsigned)nbatch);
    const bool tab_smem = tabs.h[0] == tabs.h[1] && tabs.h[1] == tabs.h[2] &&
                          (size_t)g.tlen[0] * sizeof(T) <= 56 * 1024;
    cudaError_t e;
#define B2N_LAUNCH_WIN(TABV, FWV, SMEM)                                                            \
    {                                                                                              \
        auto k = spread_window3d_kernel<T, J, TABV, FWV>;                                          \
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM));     \
        if (e != cudaSuccess) return (int)e;                                                       \
        k<<<gd, 128, (SMEM), st>>>(g, wa, (const T*)tabs.h[0], (const T*)tabs.h[1],                \
                                   (const T*)tabs.h[2], (const T*)tm_s, (const T*)wts, pt_ko, pt_kw, \
                                   perm, (const C*)samples, (C*)grid, (const C*)phase_s,           \
                                   pts_per_warp, max_slide);                                       \
    }
    const size_t stage_bytes = (size_t)4 * WinRec<T, J>::kBytes;
    const size_t stage_fw = (size_t)4 * WinRec<T, J, true>::kBytes;
    if (cplx) {
        if (wts == nullptr) return 0;                 // complex tables: plan-time weights only
        const size_t smem_c = (size_t)4 * WinRec<T, J, false, true>::kBytes;
        auto k = spread_window3d_kernel<T, J, 2, 0, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
        if (e != cudaSuccess) return (int)e;
        k<<<gd, 128, smem_c, st>>>(g, wa, (const T*)tabs.h[0], (const T*)tabs.h[1], (const T*)tabs.h[2],
                                   (const T*)tm_s, (const T*)wts, pt_ko, pt_kw, perm, (const C*)samples,
                                   (C*)grid, (const C*)phase_s, pts_per_warp, max_slide);
    } else
    if (wts != nullptr && wo.facew == 2 && sizeof(T) == 4 && J <= 6) B2N_LAUNCH_WIN(2, 2, stage_fw)
    else if (wts != nullptr && wo.facew != 0) B2N_LAUNCH_WIN(2, 1, stage_fw)
    else if (wts != nullptr) B2N_LAUNCH_WIN(2, 0, stage_bytes)
    else if (tab_smem) B2N_LAUNCH_WIN(1, 0, stage_bytes + (size_t)g.tlen[0] * sizeof(T))
    else B2N_LAUNCH_WIN(0, 0, stage_bytes)
#undef B2N_LAUNCH_WIN
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

template <typename T>
static int window_adj_t(const Geom& g, int Jk, bool cplx, const TablePtrs& tabs, const WindowOpts& wo, const void* tm_s,
                        const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                        const void* samples, void* grid, const void* phase_s, int nbatch,
                        cudaStream_t st, bool* done) {
    *done = false;
    if (g.ndim != 3) return 0;
    for (int d = 0; d < 3; d++) {
        if (g.J[d] != Jk && wts == nullptr) return 0;   // padded windows need the plan-time weights
        if (g.K[d] < Jk) return 0;
    }
#define B2N_WIN(JJ)                                                                          \
    return launch_window<T, JJ>(g, cplx, tabs, wo, tm_s, wts, pt_ko, pt_kw, perm, samples, grid,  \
                                phase_s, nbatch, st, done)
    switch (Jk) {
        case 4: B2N_WIN(4);
        case 5: B2N_WIN(5);
        case 6: B2N_WIN(6);
        case 7: B2N_WIN(7);
        case 8: B2N_WIN(8);
        default: return 0;
    }
#undef B2N_WIN
}

}  // namespace b2n
