// tiled forward interpolation, double instantiations
#include <cstring>
#include "interp_tiled.cuh"
namespace b2n {
int tiled_fwd_f64(const Geom& g, int Jk, bool cplx, bool tables_equal, const TablePtrs& tabs, const void* tm_s,
                  const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                  const int4* items, int64_t n_items, const SlotArgs& sa, const void* grid, void* out,
                  const void* phase_s, int nbatch, const FwdOpts& fo, cudaStream_t st, bool* done) {
    return tiled_fwd_t<double>(g, Jk, cplx, tables_equal, tabs, tm_s, wts, pt_ko, pt_kw, perm, items, n_items,
                           sa, grid, out, phase_s, nbatch, fo, st, done);
}
}  // namespace b2n
