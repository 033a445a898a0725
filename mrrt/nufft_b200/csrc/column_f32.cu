// column-group register-window adjoint gridding, float instantiations
#include "spread_column.cuh"
namespace b2n {
int column_adj_f32(const Geom& g, int Jk, const WindowOpts& wo, const void* wts, const int32_t* pt_kw,
                   const int32_t* perm, const void* samples, void* grid, const void* phase_s,
                   int nbatch, cudaStream_t st, bool* done) {
    return column_adj_t<float>(g, Jk, wo, wts, pt_kw, perm, samples, grid, phase_s, nbatch, st, done);
}
}  // namespace b2n
