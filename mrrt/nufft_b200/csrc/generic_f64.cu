// generic interpolators, double instantiations
#define B2N_GENERIC_TU
#include "interp_generic.cuh"
namespace b2n {
int generic_launch_f64(const Geom& g, int cplx_table, const TablePtrs& tabs, const void* tm_s,
                       const int32_t* perm, bool fwd, const void* in, void* out,
                       const void* phase_s, int nbatch, int sm_count, void* acc64, cudaStream_t st) {
    return generic_launch_t<double>(g, cplx_table, tabs, tm_s, perm, fwd, in, out, phase_s, nbatch, sm_count, acc64, st);
}
}  // namespace b2n
