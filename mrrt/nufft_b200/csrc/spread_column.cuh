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

// Per-warp staging of one batch of 32 samples (written lane-parallel by the batch phase, read
// by the sample loop).  The sample loop is bound by shared-memory WAVEFRONTS (one per cycle per
// SM; a 16-byte broadcast load costs two, a scalar or 8-byte one costs one), so it reads as
// few bytes per sample as possible -- 6 wavefronts at J = 6:
//   head  (16-byte chunks, pitch 16 B x odd: conflict-free vector stores)
//         [wA[0..J-1] | origin along a + ((action + 1) << 16) | pad]   coef_a: the only broadcast
//                              reads of the loop; the packed action is read only when the window moves
//         [wC'[0..FC-1]]       coef_c on the face, slots of one lane adjacent: 1 read / lane
//   fB    [8] complex          coef_b on the face times the sample value: 1 read / lane
// Which samples move the window travels in a register (ballot mask), not through shared memory.
template <typename T, int J> struct ColumnRec {
    using S = ColumnShape<J>;
    static constexpr int VPC = 16 / (int)sizeof(T);                        // values per 16-byte chunk
    static constexpr int NA = (J * (int)sizeof(T) + 4 + 15) / 16;          // chunks of coef_a + the packed action
    static constexpr int NC = S::FC / VPC;                                 // chunks of coef_c
    static constexpr int kTailA = J * (int)sizeof(T) - 16 * (NA - 1);      // coef_a bytes in the last chunk
    static constexpr int kAct = J * (int)sizeof(T);                        // byte offset of the packed action
    static constexpr int kC = 16 * NA;                                     // byte offset of wC'
    static constexpr int kHead = 16 * ((NA + NC) % 2 == 1 ? NA + NC : NA + NC + 1);   // 16 B x odd
    static constexpr int kFB = S::FB + 1;                                  // fB pitch in complex elements (odd)
    static constexpr int kFBOff = 32 * kHead;
    static constexpr int kBytes = (kFBOff + 32 * kFB * 2 * (int)sizeof(T) + 15) / 16 * 16;   // per warp
};

__device__ __forceinline__ float int_bits_as(int v, float) { return __int_as_float(v); }
__device__ __forceinline__ double int_bits_as(int v, double) { return __hiloint2double(0, v); }

// Plan-time COLUMN RECORDS (built by column_records_kernel below): everything the batch phase
// reads per sample, blocked by 32 samples of the column order so that one lane-pass reads
// rows of 32 consecutive values from ONE base address with immediate offsets (no per-load
// address arithmetic), already shifted to the sample's place on the face:
//   J rows     coef_a (axis 3)                                                  [32] x T each
//   8 rows     coef_b (axis 1) at face position p = origin offset + tap, zero elsewhere
//   FC rows    coef_c (axis 2) likewise, row index (c & 3) * RPL + (c >> 2)
//   4 int rows origin along a | (action + 1) << 16; face corner ab; ac; acquisition index
// action = cells the window slides before this sample (0 .. max_slide), or -1 = new window.
template <typename T, int J> struct ColumnBlock {
    using S = ColumnShape<J>;
    static constexpr int kRow = 32 * (int)sizeof(T);
    static constexpr int kRows = J + S::FB + S::FC;
    static constexpr int kInts = kRows * kRow;                 // byte offset of the int rows
    static constexpr int kBytes = kInts + 4 * 128;             // per block of 32 samples
};

template <typename T, int J>
__global__ void column_records_kernel(Geom g, const void* h0, const void* h1, const void* h2,
                                      const T* __restrict__ tm_s, const int32_t* __restrict__ pt_ko,
                                      const int32_t* __restrict__ pt_kw, const int32_t* __restrict__ perm,
                                      int max_slide, unsigned char* __restrict__ rec) {
    using S = ColumnShape<J>;
    using B = ColumnBlock<T, J>;
    const int64_t M = g.M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        unsigned char* blk = rec + (i >> 5) * B::kBytes;
        const int l = (int)(i & 31);
        const int kB = pt_kw[i], kC = pt_kw[M + i], kA = pt_kw[2 * M + i];
        const int ab = kB / S::GB * S::GB, ac = kC / S::GC * S::GC;
        const int ob = kB - ab, oc = kC - ac;
        {   // a: axis 3
            const T t = tm_s[2 * M + i];
            const int ko = pt_ko[2 * M + i];
#pragma unroll
            for (int j = 0; j < J; j++)
                ((T*)(blk + j * B::kRow))[l] =
                    j < g.J[2] ? tap_real<T>((const T*)h2, g.ncenter[2], g.tlen[2], t, ko + j, g.L, g.order) : (T)0;
        }
        {   // b: axis 1, shifted to the face
            const T t = tm_s[i];
            const int ko = pt_ko[i];
#pragma unroll
            for (int q = 0; q < S::FB; q++) {
                const int j = q - ob;
                ((T*)(blk + (J + q) * B::kRow))[l] =
                    (j >= 0 && j < g.J[0]) ? tap_real<T>((const T*)h0, g.ncenter[0], g.tlen[0], t, ko + j, g.L, g.order) : (T)0;
            }
        }
        {   // c: axis 2, shifted to the face, slots of one lane adjacent
            const T t = tm_s[M + i];
            const int ko = pt_ko[M + i];
#pragma unroll
            for (int q = 0; q < S::FC; q++) {
                const int j = q - oc;
                ((T*)(blk + (J + S::FB + (q & 3) * S::RPL + (q >> 2)) * B::kRow))[l] =
                    (j >= 0 && j < g.J[1]) ? tap_real<T>((const T*)h1, g.ncenter[1], g.tlen[1], t, ko + j, g.L, g.order) : (T)0;
            }
        }
        int act = -1;
        if (i > 0) {
            const int qB = pt_kw[i - 1], qC = pt_kw[M + i - 1], d = kA - pt_kw[2 * M + i - 1];
            if (qB / S::GB * S::GB == ab && qC / S::GC * S::GC == ac && d >= 0 && d <= max_slide) act = d;
        }
        int32_t* ir = (int32_t*)(blk + B::kInts);
        ir[l] = kA | ((act + 1) << 16);
        ir[32 + l] = ab;
        ir[64 + l] = ac;
        ir[96 + l] = perm[i];
    }
}

template <typename T, int J, int MINB>
__global__ void __launch_bounds__(128, MINB)
spread_column3d_kernel(Geom g, WindowAxes wa, const unsigned char* __restrict__ records,
                       const cplx_t<T>* __restrict__ samples, cplx_t<T>* __restrict__ grid,
                       const cplx_t<T>* __restrict__ phase_s, int pts_per_warp) {
    using C = cplx_t<T>;
    using S = ColumnShape<J>;
    using R = ColumnRec<T, J>;
    using B = ColumnBlock<T, J>;
    constexpr int RPL = S::RPL, FB = S::FB, FC = S::FC;
    constexpr int HP = R::kHead, VPC = R::VPC, NA = R::NA, NC = R::NC, FBP = R::kFB;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char* heads = dyn_smem + wib * R::kBytes;            // this warp's staging
    C* fbs = (C*)(heads + R::kFBOff);
    const int64_t M = g.M;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t begin = warp * pts_per_warp;
    if (begin >= M) return;
    const int64_t end = begin + pts_per_warp < M ? begin + pts_per_warp : M;
    const int bt = blockIdx.y;
    const C* __restrict__ sb = samples + (int64_t)bt * M;
    C* __restrict__ gb = grid + (int64_t)bt * g.PK;
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
    // 32-bit bookkeeping (M < 2^31): block index of the run, samples left in it
    const int blk0 = (int)(begin >> 5);
    const int nsamp = (int)(end - begin);

    for (int done = 0; done < nsamp; done += 32) {
        const int cnt = nsamp - done < 32 ? nsamp - done : 32;
        // this lane's column of the current block of 32 records
        const unsigned char* blk0p = records + (size_t)(blk0 + (done >> 5)) * B::kBytes;
        const unsigned char* blk = blk0p + lane * (int)sizeof(T);
        const unsigned char* blki = blk0p + B::kInts + lane * 4;
        __syncwarp();
        // the next block of this run on its way from HBM to L2 while this one is worked on
        if (done + 32 < nsamp && lane == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(blk0p + B::kBytes), "n"(B::kBytes));
        // ---- batch phase: lane = sample
        int ka_act = 1 << 16;
        if (lane < cnt) {
            ka_act = *(const int*)blki;
            C f = sb[*(const int*)(blki + 384)];
            if (phase_s != nullptr) f = cmul_conj(f, phase_s[((int64_t)(blk0 + (done >> 5)) << 5) + lane]);
            unsigned char* hb = heads + lane * HP;
            {
                T hv[NA * VPC];
#pragma unroll
                for (int e = 0; e < NA * VPC; e++) hv[e] = e < J ? *(const T*)(blk + e * B::kRow) : (T)0;
                hv[J] = int_bits_as(ka_act, (T)0);
#pragma unroll
                for (int c = 0; c < NA; c++) store16(hb + 16 * c, hv + VPC * c);
            }
            {
                T cv[FC];
#pragma unroll
                for (int e = 0; e < FC; e++) cv[e] = *(const T*)(blk + (J + FB + e) * B::kRow);
#pragma unroll
                for (int c = 0; c < NC; c++) store16(hb + R::kC + 16 * c, cv + VPC * c);
            }
            C* frb = fbs + lane * FBP;
#pragma unroll
            for (int q = 0; q < FB; q++) frb[q] = mul_w(*(const T*)(blk + (J + q) * B::kRow), f);
        }
        // samples that move the window (slide or new window); the first of the run always does
        unsigned nz = __ballot_sync(FULL, (ka_act >> 16) != 1);
        if (!have) nz |= 1u;
        __syncwarp();
        // ---- sample loop: the whole warp works on one sample
        const unsigned char* rec = heads;
        const unsigned char* const rec_end = heads + cnt * HP;
        const C* fwb = fbs + lb;
        const unsigned char* fwc = heads + R::kC + lc0 * RPL * (int)sizeof(T);
#pragma unroll 2
        for (; rec != rec_end; rec += HP, fwb += FBP, fwc += HP, nz >>= 1) {
            T hv[NA * VPC];
#pragma unroll
            for (int c = 0; c < NA; c++) {
                if (c == NA - 1 && R::kTailA <= 0) {
                    // (the last chunk holds only the packed action)
                } else if (c == NA - 1 && R::kTailA <= 8 && sizeof(T) == 4) {
                    const float2 t2 = *(const float2*)(rec + 16 * c);
                    hv[VPC * c] = t2.x;
                    hv[VPC * c + 1] = t2.y;
                } else {
                    load16(rec + 16 * c, hv + VPC * c);
                }
            }
            const C vb = fwb[0];
            C v[RPL];
            if constexpr (RPL == 2) {
                const C wc2 = *(const C*)fwc;
                v[0] = mul_w(wc2.x, vb);
                v[1] = mul_w(wc2.y, vb);
            } else {
                v[0] = mul_w(*(const T*)fwc, vb);
            }
            if (nz & 1u) {
                const int ka = *(const int*)(rec + R::kAct);
                const int act_q = have ? (ka >> 16) - 1 : -1;
                if (act_q < 0) {
                    const int KA = wa.K[0], sA = wa.stride[0];
                    if (have) {
#pragma unroll
                        for (int j = 0; j < J; j++) {
                            int kj = WA + j;
                            if (kj >= KA) kj -= KA;
#pragma unroll
                            for (int s = 0; s < RPL; s++) {
                                atomic_add_c(faceptr[s] + (int64_t)kj * sA, acc[s][j]);
                                acc[s][j] = make_c<T>(0, 0);
                            }
                        }
                    }
                    have = true;
                    WA = ka & 0xffff;
                    // face corner of this sample's column group: from the record block (rare path)
                    const int* io = (const int*)(blk0p + B::kInts) + (int)((rec - heads) / HP);
                    int kb = io[32] + lb;
                    if (kb >= wa.K[1]) kb -= wa.K[1];
#pragma unroll
                    for (int s = 0; s < RPL; s++) {
                        int kc = io[64] + lc0 + 4 * s;
                        if (kc >= wa.K[2]) kc -= wa.K[2];
                        faceptr[s] = gb + ((int64_t)kb * wa.stride[1] + (int64_t)kc * wa.stride[2]);
                    }
                } else {
                    // slide by `act_q` cells: one face per cell goes to L2, the registers shift
                    // (once per ~16 samples on the bench trajectory: not worth a second FMA block)
                    const int sA = wa.stride[0];
#pragma unroll 1
                    for (int sft = 0; sft < act_q; sft++) {
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
        const int KA = wa.K[0], sA = wa.stride[0];
#pragma unroll
        for (int j = 0; j < J; j++) {
            int kj = WA + j;
            if (kj >= KA) kj -= KA;
#pragma unroll
            for (int s = 0; s < RPL; s++) atomic_add_c(faceptr[s] + (int64_t)kj * sA, acc[s][j]);
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
                             (const C*)phase_s, pts_per_warp);
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
                          const int32_t* pt_kw, const int32_t* perm, int max_slide, void* records,
                          int nblocks, cudaStream_t st) {
    if (max_slide <= 0 || max_slide > Jk - 1) max_slide = Jk - 1;
#define B2N_COLB(JJ)                                                                              \
    case JJ:                                                                                      \
        column_records_kernel<T, JJ><<<nblocks, 256, 0, st>>>(g, tabs.h[0], tabs.h[1], tabs.h[2], \
            (const T*)tm_s, pt_ko, pt_kw, perm, max_slide, (unsigned char*)records);                         \
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
