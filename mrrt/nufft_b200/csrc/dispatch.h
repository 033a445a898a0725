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

// each returns 0 or a cudaError_t
#define B2N_DECLARE(SUF)                                                                          \
    int generic_launch_##SUF(const Geom& g, int cplx_table, const TablePtrs& tabs, const void* tm_s, \
                             const int32_t* perm, bool fwd, const void* in, void* out,           \
                             const void* phase_s, int nbatch, int sm_count, cudaStream_t st);    \
    int tiled_fwd_##SUF(const Geom& g, bool tables_equal, const TablePtrs& tabs, const void* tm_s, \
                        const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm, const int4* items, int64_t n_items, const SlotArgs& sa, const void* grid, \
                        void* out, const void* phase_s, int nbatch, int use_tma, cudaStream_t st, \
                        bool* done);                                                             \
    int slide_adj_##SUF(const Geom& g, const TablePtrs& tabs, const void* tm_s,                 \
                        const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,        \
                        const void* samples, void* grid, const void* phase_s, int nbatch,        \
                        int pts_per_warp, cudaStream_t st, bool* done);
#define B2N_DECLARE2(SUF)                                                                        \
    int tile_adj_##SUF(const Geom& g, const TablePtrs& tabs, const void* tm_s, const int32_t* pt_ko, \
                       const int32_t* pt_kw, const int32_t* perm, const int4* items,             \
                       int64_t n_items, const void* samples, void* grid, const void* phase_s,    \
                       int nbatch, int use_tma, cudaStream_t st, bool* done);
#define B2N_DECLARE3(SUF)                                                                        \
    int window_adj_##SUF(const Geom& g, const TablePtrs& tabs, int slide_axis, const void* tm_s, \
                         const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,        \
                         const void* samples, void* grid, const void* phase_s, int nbatch,       \
                         int pts_per_warp, cudaStream_t st, bool* done);                         \
    size_t window_record_bytes_##SUF(int J);                                                     \
    int window_records_build_##SUF(const Geom& g, int slide_axis, const void* wts,               \
                                   const int32_t* pt_kw, int pts_per_warp, int max_slide,        \
                                   void* recs, int sm_count, cudaStream_t st);
#define B2N_DECLARE4(SUF)                                                                        \
    int window2d_adj_##SUF(const Geom& g, const TablePtrs& tabs, const void* tm_s, const void* wts, \
                           const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,      \
                           const void* samples, void* grid, const void* phase_s, int nbatch,     \
                           int pts_per_warp, cudaStream_t st, bool* done);
B2N_DECLARE4(f32)
B2N_DECLARE4(f64)
#undef B2N_DECLARE4
B2N_DECLARE3(f32)
B2N_DECLARE3(f64)
#undef B2N_DECLARE3
B2N_DECLARE2(f32)
B2N_DECLARE2(f64)
#undef B2N_DECLARE2
B2N_DECLARE(f32)
B2N_DECLARE(f64)
#undef B2N_DECLARE

}  // namespace b2n
