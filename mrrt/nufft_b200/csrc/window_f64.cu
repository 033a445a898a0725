// register-window adjoint gridding (second generation), double instantiations
#include "spread_window.cuh"
namespace b2n {
int window_adj_f64(const Geom& g, const TablePtrs& tabs, int slide_axis, const void* tm_s,
                   const void* wts, const int32_t* pt_ko, const int32_t* pt_kw, const int32_t* perm,
                   const void* samples, void* grid, const void* phase_s, int nbatch,
                   int pts_per_warp, cudaStream_t st, bool* done) {
    return window_adj_t<double>(g, tabs, slide_axis, tm_s, wts, pt_ko, pt_kw, perm, samples, grid,
                            phase_s, nbatch, pts_per_warp, st, done);
}
size_t window_record_bytes_f64(int J) { return window_record_bytes_t<double>(J); }
int window_records_build_f64(const Geom& g, int slide_axis, const void* wts, const int32_t* pt_kw,
                             int pts_per_warp, int max_slide, void* recs, int sm_count,
                             cudaStream_t st) {
    return window_records_build_t<double>(g, slide_axis, wts, pt_kw, pts_per_warp, max_slide, recs,
                                      sm_count, st);
}
}  // namespace b2n
