// column-group register-window adjoint gridding and its plan-time records, float instantiations
#include "spread_column.cuh"
namespace b2n {
int column_adj_f32(const Geom& g, int Jk, const WindowOpts& wo, const void* records, const void* samples,
                   void* grid, const void* phase_s, int nbatch, cudaStream_t st, bool* done) {
    return column_adj_t<float>(g, Jk, wo, records, samples, grid, phase_s, nbatch, st, done);
}
size_t column_record_bytes_f32(int Jk, int64_t M) { return column_record_bytes_t<float>(Jk, M); }
int column_build_f32(const Geom& g, int Jk, const TablePtrs& tabs, const void* tm_s, const int32_t* pt_ko,
                     const int32_t* pt_kw, const int32_t* perm, int max_slide, void* records, int nblocks, cudaStream_t st) {
    return column_build_t<float>(g, Jk, tabs, tm_s, pt_ko, pt_kw, perm, max_slide, records, nblocks, st);
}
}  // namespace b2n
