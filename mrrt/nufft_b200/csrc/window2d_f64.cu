// 2-D register-window adjoint gridding, double instantiations
#include "spread_window2d.cuh"
namespace b2n {
int window2d_adj_f64(const Geom& g, int Jk, bool cplx, const TablePtrs& tabs, const void* tm_s, const void* wts,
                     const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                     const void* samples, void* grid, const void* phase_s, int nbatch,
                     const WindowOpts& wo, cudaStream_t st, bool* done) {
    return window2d_adj_t<double>(g, Jk, cplx, tabs, tm_s, wts, pt_ko, pt_kw, perm, samples, grid, phase_s,
                              nbatch, wo, st, done);
}
}  // namespace b2n
