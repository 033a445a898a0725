// Column-group register-window adjoint gridding (3-D, real table, plan-time weights).
//
// One WARP keeps an 8 x FC x J window of partial sums in registers (FC = 4 or 8): lane <->
// (b, c) position of the window FACE (b = lane & 7 along grid axis 1, c = lane >> 3 (+4 per
// slot) along grid axis 2), J accumulators along the slide axis a (grid axis 3).  The face is
// WIDER than the J x J footprint of a sample: it covers every window whose origin lies in a
// group of GB x GC = (9 - J) x (FC + 1 - J) neighbouring grid columns (3 x 3 for J = 6), so
// the samples of all those columns share one window.  Samples are sorted by (column group,
// origin along a) -- the "column order" of the plan -- and a warp walks a contiguous run of
// them: a sample whose origin sits off the face's corner just has its weights shifted on
// the face (zero elsewhere).  What this buys over one J x J face per half-warp
// (spread_window.cuh):
//   * the whole warp works on ONE sample: no divergence between two half-warps that disagree
//     about sliding (both FMA blocks were issued in 46 % of the warp iterations);
//   * the window slides once per (column group, a) instead of once per occupied cell:
//     ~16 samples per slide on the bench trajectory instead of 2.6, and one 64-cell face per
//     slide instead of nine 36-cell faces -> 3.4x fewer L2 reductions;
//   * 24 accumulator registers instead of 36 -> more resident warps.
// The price: only J*J of the 32*RPL face slots carry a non-zero weight (56 % at J = 6).
//
// Arithmetic per sample follows c/nufft_table.template.c:1122-1163 (products of the three
// axis coefficients with the sample value, accumulated per cell); the association is
// ((coef_b * f) * coef_c) * coef_a.
#pragma once
#include "common.cuh"
#include "dispatch.h"
#include "spread_window.cuh"   // WindowAxes, store16 / load16

namespace b2n {

template <int J> struct ColumnShape {
    static constexpr int FB = 8;                      // face cells along b (lanes & 7)
    static constexpr int RPL = (J + 3) / 4;           // face slots per lane
    static constexpr int FC = 4 * RPL;                // face cells along c
    static constexpr int GB = FB - J + 1;             // grid columns per group along b
    static constexpr int GC = FC - J + 1;             // ... along c
};

template <typename T, int J> struct ColumnRec {
    using S = ColumnShape<J>;
    static constexpr int VPC = 16 / (int)sizeof(T);                        // values per 16-byte chunk
    static constexpr int HV = (J + 2 + VPC - 1) / VPC * VPC;               // head values incl. padding
    static constexpr int kHeadInt = HV * (int)sizeof(T);                   // offset of (kA, ab, ac, act)
    static constexpr int kHeadMin = kHeadInt + 16;
    static constexpr int kHead = (kHeadMin / 16) % 2 == 1 ? kHeadMin : kHeadMin + 16;   // 16 B x odd
    static constexpr int kFaceElems = (S::FB + S::FC) % 2 == 1 ? S::FB + S::FC : S::FB + S::FC + 1;
    static constexpr int kBytes = 32 * kHead + (32 * kFaceElems * (int)sizeof(T) + 15) / 16 * 16;   // per warp
};

// Plan-time COLUMN RECORDS (built by column_records_kernel below): everything the batch phase
// reads per sample, blocked by 32 samples of the column order so that one lane-pass reads
// rows of 32 consecutive values from ONE base address with immediate offsets (no per-load
// address arithmetic, which was 2/3 of the batch phase):
//   rows 0 .. 3J-1   weights along a (axis 3), b (axis 1), c (axis 2)     [32] x T each
//   4 int rows       origin along a | face corner ab + (ob << 16) | ac + (oc << 16) | acquisition index
template <typename T, int J> struct ColumnBlock {
    static constexpr int kRow = 32 * (int)sizeof(T);
    static constexpr int kInts = 3 * J * kRow;                 // byte offset of the int rows
    static constexpr int kBytes = kInts + 4 * 128;             // per block of 32 samples
};

template <typename T, int J>
__global__ void column_records_kernel(Geom g, const void* h0, const void* h1, const void* h2,
                                      const T* __restrict__ tm_s, const int32_t* __restrict__ pt_ko,
                                      const int32_t* __restrict__ pt_kw, const int32_t* __restrict__ perm,
                                      unsigned char* __restrict__ rec) {
    using S = ColumnShape<J>;
    using B = ColumnBlock<T, J>;
    const int64_t M = g.M;
    const void* tabs[3] = {h0, h1, h2};
    const int axes[3] = {2, 0, 1};                             // a, b, c
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        unsigned char* blk = rec + (i >> 5) * B::kBytes;
        const int l = (int)(i & 31);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int d = axes[r];
            const T t = tm_s[(int64_t)d * M + i];
            const int ko = pt_ko[(int64_t)d * M + i];
#pragma unroll
            for (int j = 0; j < J; j++)
                ((T*)(blk + (r * J + j) * B::kRow))[l] =
                    j < g.J[d] ? tap_real<T>((const T*)tabs[d], g.ncenter[d], g.tlen[d], t, ko + j, g.L, g.order) : (T)0;
        }
        const int kB = pt_kw[i], kC = pt_kw[M + i];
        const int ab = kB / S::GB * S::GB, ac = kC / S::GC * S::GC;
        int32_t* ir = (int32_t*)(blk + B::kInts);
        ir[l] = pt_kw[2 * M + i];
        ir[32 + l] = ab | ((kB - ab) << 16);
        ir[64 + l] = ac | ((kC - ac) << 16);
        ir[96 + l] = perm[i];
    }
}

template <typename T, int J, int MINB>
__global__ void __launch_bounds__(128, MINB)
spread_column3d_kernel(Geom g, WindowAxes wa, const unsigned char* __restrict__ records,
                       const cplx_t<T>* __restrict__ samples, cplx_t<T>* __restrict__ grid,
                       const cplx_t<T>* __restrict__ phase_s, int pts_per_warp, int max_slide) {
    using C = cplx_t<T>;
    using S = ColumnShape<J>;
    using R = ColumnRec<T, J>;
    using B = ColumnBlock<T, J>;
    constexpr int RPL = S::RPL, FB = S::FB, FC = S::FC;
    constexpr int HP = R::kHead, HI = R::kHeadInt, HV = R::HV, VPC = R::VPC, FP = R::kFaceElems;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char* heads = dyn_smem + wib * R::kBytes;            // this warp's head records
    T* faces = (T*)(heads + 32 * HP);                             // ... and face weight rows
    const int64_t M = g.M;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t begin = warp * pts_per_warp;
    if (begin >= M) return;
    const int64_t end = begin + pts_per_warp < M ? begin + pts_per_warp : M;
    const int bt = blockIdx.y;
    const C* __restrict__ sb = samples + (int64_t)bt * M;
    C* __restrict__ gb = grid + (int64_t)bt * g.PK;
    const int KA = wa.K[0], KB = wa.K[1], KC = wa.K[2];
    const int sA = wa.stride[0], sB = wa.stride[1], sC = wa.stride[2];
    const int lb = lane & 7, lc0 = lane >> 3;

    C acc[RPL][J];
    C* faceptr[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        faceptr[s] = gb;
#pragma unroll
        for (int j = 0; j < J; j++) acc[s][j] = make_c<T>(0, 0);
    }
    int WA = 0;
    bool have = false;
    int pkA = -1 << 30, pab = -1, pac = -1;   // previous sample: origin along a, face corner
    // this lane's column of the current block of 32 records
    const unsigned char* blk = records + (begin >> 5) * B::kBytes + lane * (int)sizeof(T);
    const unsigned char* blki = records + (begin >> 5) * B::kBytes + B::kInts + lane * 4;

    for (int64_t base = begin; base < end; base += 32, blk += B::kBytes, blki += B::kBytes) {
        const int cnt = (int)(end - base < 32 ? end - base : 32);
        __syncwarp();
        // ---- batch phase: lane = sample
        int kA = 0, ab = 0, ac = 0;
        // the next block of this run on its way from HBM to L2 while this one is worked on
        if (base + 32 < end && lane * 128 < B::kBytes)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + B::kBytes + lane * (128 - (int)sizeof(T))));
        if (lane < cnt) {
            kA = *(const int*)blki;
            const int pb = *(const int*)(blki + 128), pc = *(const int*)(blki + 256);
            C f = sb[*(const int*)(blki + 384)];
            ab = pb & 0xffff;
            ac = pc & 0xffff;
            const int ob = pb >> 16, oc = pc >> 16;
            T hv[HV];
#pragma unroll
            for (int e = 0; e < HV; e++) hv[e] = (T)0;
#pragma unroll
            for (int j = 0; j < J; j++) hv[j] = *(const T*)(blk + j * B::kRow);
            T* frb = faces + lane * FP + ob;
            T* frc = faces + lane * FP + FB + oc;
#pragma unroll
            for (int j = 0; j < J; j++) {
                frb[j] = *(const T*)(blk + (J + j) * B::kRow);
                frc[j] = *(const T*)(blk + (2 * J + j) * B::kRow);
            }
            // the face positions this sample's window does not reach
#pragma unroll
            for (int z = 0; z < FB - J; z++) frb[z < ob ? z - ob : z + J - ob] = (T)0;
#pragma unroll
            for (int z = 0; z < FC - J; z++) frc[z < oc ? z - oc : z + J - oc] = (T)0;
            if (phase_s != nullptr) f = cmul_conj(f, phase_s[base + lane]);
            hv[J] = f.x;
            hv[J + 1] = f.y;
            unsigned char* hb = heads + lane * HP;
#pragma unroll
            for (int c = 0; c < HV / VPC; c++) store16(hb + 16 * c, hv + VPC * c);
        }
        {
            int qA = __shfl_up_sync(FULL, kA, 1), qb = __shfl_up_sync(FULL, ab, 1),
                qc = __shfl_up_sync(FULL, ac, 1);
            if (lane == 0) { qA = pkA; qb = pab; qc = pac; }
            const int d = kA - qA;
            const int act = (ab == qb && ac == qc && d >= 0 && d <= max_slide) ? d : -1;
            if (lane < cnt) *(int4*)(heads + lane * HP + HI) = make_int4(kA, ab, ac, act);
            pkA = __shfl_sync(FULL, kA, cnt - 1);
            pab = __shfl_sync(FULL, ab, cnt - 1);
            pac = __shfl_sync(FULL, ac, cnt - 1);
        }
        __syncwarp();
        // ---- sample loop: the whole warp works on one sample.  Only the action code is
        // prefetched (one register; past the last record it reads the first face row: unused)
        const unsigned char* rec = heads;
        const unsigned char* const rec_end = heads + cnt * HP;
        const T* fwb = faces + lb;
        const T* fwc = faces + FB + lc0;
        int act_next = *(const int*)(rec + HI + 12);
#pragma unroll 2
        for (; rec != rec_end; rec += HP, fwb += FP, fwc += FP) {
            const int act = act_next;
            act_next = *(const int*)(rec + HP + HI + 12);
            T hv[HV];
#pragma unroll
            for (int c = 0; c < HV / VPC; c++) load16(rec + 16 * c, hv + VPC * c);
            const C vb = mul_w(fwb[0], make_c<T>(hv[J], hv[J + 1]));
            C v[RPL];
#pragma unroll
            for (int s = 0; s < RPL; s++) v[s] = mul_w(fwc[4 * s], vb);
            if (act != 0) {
                if (act < 0) {
                    if (have) {
#pragma unroll
                        for (int j = 0; j < J; j++) {
                            int ka = WA + j;
                            if (ka >= KA) ka -= KA;
#pragma unroll
                            for (int s = 0; s < RPL; s++) {
                                atomic_add_c(faceptr[s] + (int64_t)ka * sA, acc[s][j]);
                                acc[s][j] = make_c<T>(0, 0);
                            }
                        }
                    }
                    have = true;
                    const int4 ko = *(const int4*)(rec + HI);
                    WA = ko.x;
                    int kb = ko.y + lb;
                    if (kb >= KB) kb -= KB;
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        int kc = ko.z + lc0 + 4 * s;
                        if (kc >= KC) kc -= KC;
                        faceptr[s] = gb + ((int64_t)kb * sB + (int64_t)kc * sC);
                    }
                } else {
                    // slide by `act` cells: one face per cell goes to L2, the registers shift
                    // (once per ~16 samples on the bench trajectory: not worth a second FMA block)
#pragma unroll 1
                    for (int sft = 0; sft < act; sft++) {
#pragma unroll
                        for (int s = 0; s < RPL; s++) {
                            atomic_add_c(faceptr[s] + (int64_t)WA * sA, acc[s][0]);
#pragma unroll
                            for (int j = 0; j + 1 < J; j++) acc[s][j] = acc[s][j + 1];
                            acc[s][J - 1] = make_c<T>(0, 0);
                        }
                        WA++;   // stays < KA: it ends at this sample's wrapped origin
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < RPL; s++)
#pragma unroll
                for (int j = 0; j < J; j++) acc[s][j] = fma_w(hv[j], v[s], acc[s][j]);
        }
    }
    if (have) {
#pragma unroll
        for (int j = 0; j < J; j++) {
            int ka = WA + j;
            if (ka >= KA) ka -= KA;
#pragma unroll
            for (int s = 0; s < RPL; s++) atomic_add_c(faceptr[s] + (int64_t)ka * sA, acc[s][j]);
        }
    }
}

// whether the column kernel can serve a plan: every axis at least as long as the face
template <int J> static bool column_fits(const Geom& g) {
    using S = ColumnShape<J>;
    return g.K[0] >= S::FB && g.K[1] >= S::FC && g.K[2] >= J;
}

template <typename T, int J>
static int launch_column(const Geom& g, const WindowOpts& wo, const void* records, const void* samples,
                         void* grid, const void* phase_s, int nbatch, cudaStream_t st, bool* done) {
    using C = cplx_t<T>;
    if (!column_fits<J>(g) || records == nullptr) return 0;
    int max_slide = wo.max_slide;
    if (max_slide <= 0 || max_slide > J - 1) max_slide = J - 1;
    const int pts_per_warp = (wo.pts_per_warp + 31) / 32 * 32;
    const int64_t nwarps = (g.M + pts_per_warp - 1) / pts_per_warp;
    const int64_t nblocks = (nwarps + 3) / 4;
    if (nblocks > 0x7fffffff || nbatch > 65535) return 0;
    WindowAxes wa;
    wa.ax[0] = 2; wa.ax[1] = 0; wa.ax[2] = 1;
    const int strides[3] = {1, g.K[0], g.K[0] * g.K[1]};
    for (int r = 0; r < 3; r++) { wa.K[r] = g.K[wa.ax[r]]; wa.stride[r] = strides[wa.ax[r]]; }
    constexpr int MINB = sizeof(T) == 4 ? (J <= 6 ? 7 : 5) : 3;
    auto k = spread_column3d_kernel<T, J, MINB>;
    const size_t smem = (size_t)4 * ColumnRec<T, J>::kBytes;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 gd((unsigned)nblocks, (unsigned)nbatch);
    k<<<gd, 128, smem, st>>>(g, wa, (const unsigned char*)records, (const C*)samples, (C*)grid,
                             (const C*)phase_s, pts_per_warp, max_slide);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
}

// bytes of the column records of M samples at width J (0: no column kernel at this width)
template <typename T> static size_t column_record_bytes_t(int Jk, int64_t M) {
    const size_t nblk = (size_t)((M + 31) / 32);
    switch (Jk) {
        case 4: return nblk * ColumnBlock<T, 4>::kBytes;
        case 5: return nblk * ColumnBlock<T, 5>::kBytes;
        case 6: return nblk * ColumnBlock<T, 6>::kBytes;
        case 7: return nblk * ColumnBlock<T, 7>::kBytes;
        case 8: return nblk * ColumnBlock<T, 8>::kBytes;
        default: return 0;
    }
}

template <typename T>
static int column_build_t(const Geom& g, int Jk, const TablePtrs& tabs, const void* tm_s, const int32_t* pt_ko,
                          const int32_t* pt_kw, const int32_t* perm, void* records, int nblocks,
                          cudaStream_t st) {
#define B2N_COLB(JJ)                                                                              \
    case JJ:                                                                                      \
        column_records_kernel<T, JJ><<<nblocks, 256, 0, st>>>(g, tabs.h[0], tabs.h[1], tabs.h[2], \
            (const T*)tm_s, pt_ko, pt_kw, perm, (unsigned char*)records);                         \
        break;
    switch (Jk) {
        B2N_COLB(4) B2N_COLB(5) B2N_COLB(6) B2N_COLB(7) B2N_COLB(8)
        default: return (int)cudaErrorInvalidValue;
    }
#undef B2N_COLB
    return (int)cudaGetLastError();
}

template <typename T>
static int column_adj_t(const Geom& g, int Jk, const WindowOpts& wo, const void* records, const void* samples,
                        void* grid, const void* phase_s, int nbatch, cudaStream_t st, bool* done) {
    *done = false;
    if (g.ndim != 3) return 0;
#define B2N_COL(JJ) return launch_column<T, JJ>(g, wo, records, samples, grid, phase_s, nbatch, st, done)
    switch (Jk) {
        case 4: B2N_COL(4);
        case 5: B2N_COL(5);
        case 6: B2N_COL(6);
        case 7: B2N_COL(7);
        case 8: B2N_COL(8);
        default: return 0;
    }
#undef B2N_COL
}

}  // namespace b2n
