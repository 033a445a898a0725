// Host-side entry points of the per-precision kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace b2n {

struct TablePtrs {
    const void* h[3];
};

// forward "slots" (one or two same-cell samples per thread, see aux_kernels.cuh).
// slots == nullptr: no pairing.  packed != 0: the per-slot arrays below are laid out in
// slot order (coalesced whatever the slot order is); otherwise the kernel goes through
// slots[] into the sample-ordered arrays.
struct SlotArgs {
    const uint32_t* slots = nullptr;   // [ns] (first sample << 1) | has_partner
    int64_t ns = 0;
    int packed = 0;
    const void* wts2 = nullptr;        // [sum J][ns] pairs (weight, partner's weight or 0)
    const int32_t* kw = nullptr;       // [ndim][ns] wrapped window origin
    const int32_t* perm = nullptr;     // [2][ns] output position (partner: -1 when none)
    const void* phase2 = nullptr;      // [ns] pairs of sample phases, or nullptr
};

// launch options of the tiled forward kernel
struct FwdOpts {
    int use_tma = 1;         // 0 = cooperative tile loads instead of the TMA box load
    int pitch = 0;           // shared-memory row pitch of the tile in cells (0 = automatic)
};

// launch options of the register-window adjoint kernels (2-D and 3-D)
struct WindowOpts {
    int slide_axis = 0;      // 3-D: grid axis the window slides along (2 with the adjoint sort order, else 0)
    int pts_per_warp = 256;  // samples per warp (two half-warp runs)
    int max_slide = 0;       // longest slide in cells before a new window is started (0 = J - 1)
    int facew = 0;           // 0 scalar staging records, 1 face-weight staging, 2 ... at 5 CTAs / SM
};

// run-time radix schedule of the general own axis-3 FFT pass (fft_axis3.cuh)
struct Axis3Plan {
    int L;            // transform length K3
    int npass;
    int radix[12];
};

// radices 8.., 4.., 2.., 3.. with one (2, 3) pair merged into a radix-6 pass (384 = 8 * 8 * 6:
// three passes instead of four); returns false when L has another prime factor
static inline bool axis3_factor(int L, Axis3Plan* ap) {
    ap->L = L;
    ap->npass = 0;
    if (L < 2) return false;
    int n2 = 0, n3 = 0;
    while (L % 8 == 0) { ap->radix[ap->npass++] = 8; L /= 8; }
    while (L % 4 == 0) { ap->radix[ap->npass++] = 4; L /= 4; }
    while (L % 2 == 0) { n2++; L /= 2; }
    while (L % 3 == 0) { n3++; L /= 3; }
    if (L != 1) return false;
    const bool six = n2 > 0 && n3 > 0;
    if (six) { n2--; n3--; }
    while (n2-- > 0 && ap->npass < 12) ap->radix[ap->npass++] = 2;
    while (n3-- > 0 && ap->npass < 12) ap->radix[ap->npass++] = 3;
    if (six && ap->npass < 12) ap->radix[ap->npass++] = 6;
    return ap->npass <= 11;
}

// Face of the column-group adjoint kernel (spread_column.cuh: ColumnShape<J>) for window width J
// in 4..8: FB x FC grid cells shared by the windows of GB x GC neighbouring grid columns.  The
// plan builds its column sort order from GB, GC and checks that the grid is as large as the face.
inline bool column_shape(int J, int* FB, int* FC, int* GB, int* GC) {
    if (J < 4 || J > 8) return false;
    *FB = 8;
    *FC = 4 * ((J + 3) / 4);
    *GB = *FB - J + 1;
    *GC = *FC - J + 1;
    return true;
}

// generic_launch: acc64 = zeroed complex128 scratch grid [nbatch][PK] for the float adjoint's
// per-cell sums (nullptr: float atomics straight into `out`; ignored by the forward and by f64).
// each returns 0 or a cudaError_t; *done tells whether the kernel family took the call.
// Jk: compile-time window width to run (>= every g.J[d]; wider than g.J[d] only with plan-time
// weights, whose extra taps are zero: aux_kernels.cuh:point_weights_kernel)
#define B2N_DECLARE(SUF)                                                                          \
    int generic_launch_##SUF(const Geom& g, int cplx_table, const TablePtrs& tabs, const void* tm_s, \
                             const int32_t* perm, bool fwd, const void* in, void* out,           \
                             const void* phase_s, int nbatch, int sm_count, void* acc64,         \
                             cudaStream_t st);                                                   \
    int tiled_fwd_##SUF(const Geom& g, int Jk, bool cplx, bool tables_equal, const TablePtrs& tabs, const void* tm_s, \
                        const void* wts, const int32_t* pt_ko, const int32_t* pt_kw,             \
                        const int32_t* perm, const int4* items, int64_t n_items,                 \
                        const SlotArgs& sa, const void* grid, void* out, const void* phase_s,    \
                        int nbatch, const FwdOpts& fo, cudaStream_t st, bool* done);             \
    int window_adj_##SUF(const Geom& g, int Jk, bool cplx, const TablePtrs& tabs, const WindowOpts& wo,             \
                         const void* tm_s, const void* wts, const int32_t* pt_ko,                \
                         const int32_t* pt_kw, const int32_t* perm, const void* samples,         \
                         void* grid, const void* phase_s, int nbatch, cudaStream_t st,           \
                         bool* done);                                                            \
    int column_adj_##SUF(const Geom& g, int Jk, const WindowOpts& wo, const void* records,       \
                         const void* samples, void* grid, const void* phase_s, int nbatch,       \
                         cudaStream_t st, bool* done);                                           \
    size_t column_record_bytes_##SUF(int Jk, int64_t M);                                         \
    int column_build_##SUF(const Geom& g, int Jk, const TablePtrs& tabs, const void* tm_s,       \
                           const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,      \
                           int max_slide, void* records, int nblocks, cudaStream_t st);          \
    int window2d_adj_##SUF(const Geom& g, int Jk, bool cplx, const TablePtrs& tabs, const void* tm_s, const void* wts, \
                           const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,      \
                           const void* samples, void* grid, const void* phase_s, int nbatch,     \
                           const WindowOpts& wo, cudaStream_t st, bool* done);
B2N_DECLARE(f32)
B2N_DECLARE(f64)
#undef B2N_DECLARE

}  // namespace b2n
