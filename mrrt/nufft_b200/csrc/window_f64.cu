// register-window adjoint gridding, double instantiations
#include "spread_window.cuh"
namespace b2n {
int window_adj_f64(const Geom& g, int Jk, bool cplx, const TablePtrs& tabs, const WindowOpts& wo, const void* tm_s,
                   const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                   const void* samples, void* grid, const void* phase_s, int nbatch,
                   cudaStream_t st, bool* done) {
    return window_adj_t<double>(g, Jk, cplx, tabs, wo, tm_s, wts, pt_ko, pt_kw, perm, samples, grid, phase_s,
                            nbatch, st, done);
}
}  // namespace b2n
