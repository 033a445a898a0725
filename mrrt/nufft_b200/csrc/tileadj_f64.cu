// tiled sliding-window adjoint gridding, double instantiations
#include <cstring>
#include "spread_tile.cuh"
namespace b2n {
int tile_adj_f64(const Geom& g, const TablePtrs& tabs, const void* tm_s, const int32_t* pt_ko,
                 const int32_t* pt_kw, const int32_t* perm, const int4* items, int64_t n_items,
                 const void* samples, void* grid, const void* phase_s, int nbatch, int use_tma,
                 cudaStream_t st, bool* done) {
    return tile_adj_t<double>(g, tabs, tm_s, pt_ko, pt_kw, perm, items, n_items, samples, grid, phase_s,
                          nbatch, use_tma, st, done);
}
}  // namespace b2n
