// Plan-time preprocessing (tm, window origins, bin ids, sort keys), the fused
// scale / zero-pad / phase / crop kernels around cuFFT, and the fixed-width (ELL)
// sparse-mode kernels.
#pragma once
#include "common.cuh"

namespace b2n {

// ---------------------------------------------------------------------------------
// trajectory preprocessing (new step; integer results are bit-exact vs
// oracle/nufft_oracle.py:bin_sort)
// ---------------------------------------------------------------------------------
template <typename T> struct Gam { T g[3]; };

template <typename T>
__global__ void prep_points_kernel(Geom g, Gam<T> gam, int kind, int jmax, const T* __restrict__ coords,
                                   T* __restrict__ tm, uint64_t* __restrict__ keys,
                                   uint64_t* __restrict__ keys_b, int32_t* __restrict__ bin_ids,
                                   int32_t* __restrict__ iota, int* __restrict__ nonfinite) {
    const int64_t M = g.M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t bin = 0, cell = 0, bin_b = 0, cell_b = 0;
        int64_t bstride = 1, cstride = 1, bstride_b = 1, cstride_b = 1;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < kMaxDim; d++) {
            if (d < g.ndim) {
                T t = coords[(int64_t)d * M + i];
                // tm = omega / gam in the precision dtype (_nufft.py:338-342)
                if (kind == 1) t = div_rn(t, gam.g[d]);
                ok = ok && is_finite(t);
                tm[(int64_t)d * M + i] = t;
                int kw = 0;
                if (ok) kw = local_index(window_origin<T>(t, g.J[d]), g.Kg[d], g.korg[d]);
                // slab plans: the whole window must lie inside the rows this plan holds
                if (ok && g.Kg[d] != g.K[d] && kw + jmax > g.K[d]) { atomicExch(nonfinite, 2); kw = 0; }
                bin += (int64_t)(kw / g.tile[d]) * bstride;
                cell += (int64_t)(kw % g.tile[d]) * cstride;
                // adjoint order: its own (longer) bins, LAST axis fastest inside the bin
                bin_b += (int64_t)(kw / g.tile_b[d]) * bstride_b;
                if (g.colmode) cell_b += (int64_t)(kw % g.tile_b[d]) * cstride_b;   // last axis slowest
                else cell_b = cell_b * g.tile_b[d] + (kw % g.tile_b[d]);
                bstride_b *= g.nbin_b[d];
                cstride_b *= g.tile_b[d];
                bstride *= g.nbin[d];
                cstride *= g.tile[d];
            }
        }
        if (!ok) atomicMax(nonfinite, 3);
        keys[i] = (uint64_t)(bin * cstride + cell);
        if (keys_b != nullptr) keys_b[i] = (uint64_t)(bin_b * cstride_b + cell_b);
        bin_ids[i] = (int32_t)bin;
        iota[i] = (int32_t)i;
    }
}

template <typename T>
__global__ void gather_points_kernel(int ndim, int64_t M, const int32_t* __restrict__ perm,
                                     const T* __restrict__ tm, T* __restrict__ tm_s) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = perm[i];
        for (int d = 0; d < ndim; d++) tm_s[(int64_t)d * M + i] = tm[(int64_t)d * M + m];
    }
}

// per sorted sample: unwrapped window origin ko = 1 + floor(t - J/2.) (double arithmetic,
// template.c:865-867) and its periodic wrap kw, so the hot kernels do neither double
// arithmetic nor integer division
template <typename T>
__global__ void point_windows_kernel(Geom g, const T* __restrict__ tm_s, int32_t* __restrict__ pt_ko,
                                     int32_t* __restrict__ pt_kw) {
    const int64_t M = g.M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        for (int d = 0; d < g.ndim; d++) {
            const int ko = window_origin<T>(tm_s[(int64_t)d * M + i], g.J[d]);
            pt_ko[(int64_t)d * M + i] = ko;
            pt_kw[(int64_t)d * M + i] = local_index(ko, g.Kg[d], g.korg[d]);
        }
    }
}

struct TabArgs {
    const void* h[3];
};

// per sorted sample the 3J interpolation weights themselves (real tables): the table is
// consulted once at plan time -- same expression as in the kernels, so the values are
// bit-identical -- and the hot kernels stream the weights instead of gathering from the
// table (a shared-memory gather costs several bank-conflict wavefronts per tap and the
// forward kernel is shared-memory-bandwidth bound).  Layout [sum(J)][M], sample fastest.
// Jk: kernel width the hot kernels are compiled for (>= every J[d]); an axis with J[d] < Jk gets
// Jk - J[d] trailing taps of weight ZERO, so unequal / odd widths run on the equal-width kernels
// (the extra taps multiply grid cells by 0 in the forward and add 0 in the adjoint).
// CT: complex table (phasing="complex"): the weights are complex (template.c:97-143, :623-709)
template <typename T, bool CT = false>
__global__ void point_weights_kernel(Geom g, TabArgs tabs, int Jk, const T* __restrict__ tm_s,
                                     const int32_t* __restrict__ pt_ko,
                                     typename WeightT<T, CT>::type* __restrict__ wts) {
    const int64_t M = g.M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        int row = 0;
        for (int d = 0; d < g.ndim; d++) {
            const T t = tm_s[(int64_t)d * M + i];
            const int ko = pt_ko[(int64_t)d * M + i];
            for (int j = 0; j < Jk; j++, row++) {
                if constexpr (CT) {
                    wts[(int64_t)row * M + i] =
                        j < g.J[d] ? tap_cplx<T>((const cplx_t<T>*)tabs.h[d], g.ncenter[d], g.tlen[d], t, ko + j, g.L, g.order)
                                   : make_c<T>(0, 0);
                } else {
                    wts[(int64_t)row * M + i] =
                        j < g.J[d] ? tap_real<T>((const T*)tabs.h[d], g.ncenter[d], g.tlen[d], t, ko + j, g.L, g.order)
                                   : (T)0;
                }
            }
        }
    }
}

template <typename C>
__global__ void gather_c_kernel(int64_t M, const int32_t* __restrict__ perm,
                                const C* __restrict__ src, C* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[perm[i]];
}

// first sorted position of every non-empty bin (bin_start pre-filled with -1)
__global__ void bin_start_kernel(int64_t M, int cells_per_tile, const uint64_t* __restrict__ keys_s,
                                 int32_t* __restrict__ bin_start) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = (int64_t)(keys_s[i] / (uint64_t)cells_per_tile);
        if (i == 0 || (int64_t)(keys_s[i - 1] / (uint64_t)cells_per_tile) != b)
            bin_start[b] = (int32_t)i;
    }
}

// ---- forward "slots": one or two consecutive sorted samples that sit in the SAME cell.
// A thread of the tiled forward kernel then reads every tap of the shared window once for
// both samples.  Pairs are formed greedily inside each run of equal sort keys (positions
// 0|1, 2|3, ... of the run).
// head[i] = i where a run of equal keys starts, else 0 (max-scanned into the run start)
__global__ void slot_heads_kernel(int64_t M, const uint64_t* __restrict__ keys_s,
                                  int32_t* __restrict__ head) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        head[i] = (i == 0 || keys_s[i] != keys_s[i - 1]) ? (int32_t)i : 0;
}
// isslot[i] = 1 when sample i opens a slot (even position inside its run)
__global__ void slot_flags_kernel(int64_t M, const int32_t* __restrict__ runstart,
                                  int32_t* __restrict__ isslot) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        isslot[i] = (((int32_t)i - runstart[i]) & 1) == 0 ? 1 : 0;
}
// slots[slotidx[i]] = (i << 1) | has_partner
__global__ void slot_write_kernel(int64_t M, const uint64_t* __restrict__ keys_s,
                                  const int32_t* __restrict__ isslot,
                                  const int32_t* __restrict__ slotidx, uint32_t* __restrict__ slots) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (isslot[i]) {
            const uint32_t pair = (i + 1 < M && keys_s[i + 1] == keys_s[i]) ? 1u : 0u;
            slots[slotidx[i]] = ((uint32_t)i << 1) | pair;
        }
    }
}
// ---- column-interleaved slot order inside a bin.  The forward tile has a row pitch of whole
// bank rows, so the bank of a tap depends on the axis-1 position ("column") of the window
// only; a half-warp is conflict-free when its 16 lanes sit in 16 different columns.  Slots
// are therefore reordered inside their bin by (rank within column, column).
// key = bin * T0 + column
__global__ void slot_colkey_kernel(int64_t ns, int cells_per_tile, int T0,
                                   const uint64_t* __restrict__ keys_s,
                                   const uint32_t* __restrict__ slots, uint64_t* __restrict__ key) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ns;
         s += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys_s[slots[s] >> 1];
        const uint64_t bin = k / (uint64_t)cells_per_tile;
        const uint64_t col = (k % (uint64_t)cells_per_tile) % (uint64_t)T0;
        key[s] = bin * (uint64_t)T0 + col;
    }
}
// head[s] = s where a (bin, column) group starts in the column-sorted slot list, else 0
__global__ void slot_grouphead_kernel(int64_t ns, const uint64_t* __restrict__ key,
                                      int32_t* __restrict__ head) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ns;
         s += (int64_t)gridDim.x * blockDim.x)
        head[s] = (s == 0 || key[s] != key[s - 1]) ? (int32_t)s : 0;
}
// final key = (bin << 40) | (rank << 8) | column
__global__ void slot_rankkey_kernel(int64_t ns, int T0, const uint64_t* __restrict__ key,
                                    const int32_t* __restrict__ groupstart,
                                    uint64_t* __restrict__ key_out) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ns;
         s += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t bin = key[s] / (uint64_t)T0, col = key[s] % (uint64_t)T0;
        const uint64_t rank = (uint64_t)((int32_t)s - groupstart[s]);
        key_out[s] = (bin << 40) | (rank << 8) | col;
    }
}

// ---- slot-ordered copies of what the forward kernel reads per slot
__global__ void slot_pack_kernel(int64_t ns, int ndim, int64_t M, const uint32_t* __restrict__ slots,
                                 const int32_t* __restrict__ pt_kw, const int32_t* __restrict__ perm,
                                 int32_t* __restrict__ slot_kw, int32_t* __restrict__ slot_perm) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ns;
         s += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = slots[s];
        const int64_t i = u >> 1;
        for (int d = 0; d < ndim; d++) slot_kw[(int64_t)d * ns + s] = pt_kw[(int64_t)d * M + i];
        slot_perm[s] = perm[i];
        slot_perm[ns + s] = (u & 1u) ? perm[i + 1] : -1;
    }
}
template <typename C>
__global__ void slot_phase_kernel(int64_t ns, const uint32_t* __restrict__ slots,
                                  const int32_t* __restrict__ perm, const C* __restrict__ phase,
                                  C* __restrict__ phase2) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ns;
         s += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = slots[s];
        const int64_t i = u >> 1;
        phase2[2 * s] = phase[perm[i]];
        C z;
        z.x = 0;
        z.y = 0;
        phase2[2 * s + 1] = (u & 1u) ? phase[perm[i + 1]] : z;
    }
}
// (weight, partner's weight or 0) per slot, same expression as point_weights_kernel
template <typename T>
__global__ void slot_weights_kernel(Geom g, TabArgs tabs, int Jk, int64_t ns, const uint32_t* __restrict__ slots,
                                    const T* __restrict__ tm_s, const int32_t* __restrict__ pt_ko,
                                    typename Cplx<T>::type* __restrict__ wts2) {
    const int64_t M = g.M;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ns;
         s += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = slots[s];
        const int64_t i = u >> 1;
        const bool pair = (u & 1u) != 0;
        int row = 0;
        for (int d = 0; d < g.ndim; d++) {
            const T t = tm_s[(int64_t)d * M + i];
            const T tq = tm_s[(int64_t)d * M + (pair ? i + 1 : i)];
            const int ko = pt_ko[(int64_t)d * M + i];
            // same wrapped cell, but the partner may sit whole periods away: its own origin
            const int koq = pt_ko[(int64_t)d * M + (pair ? i + 1 : i)];
            for (int j = 0; j < Jk; j++, row++) {
                typename Cplx<T>::type ww;
                ww.x = j < g.J[d] ? tap_real<T>((const T*)tabs.h[d], g.ncenter[d], g.tlen[d], t, ko + j, g.L, g.order) : (T)0;
                ww.y = (pair && j < g.J[d])
                           ? tap_real<T>((const T*)tabs.h[d], g.ncenter[d], g.tlen[d], tq, koq + j, g.L, g.order)
                           : (T)0;
                wts2[(int64_t)row * ns + s] = ww;
            }
        }
    }
}

// slot index of the first sample of every non-empty bin (bin starts always open a slot)
__global__ void bin_slot_start_kernel(int64_t nbins, const int32_t* __restrict__ bin_start,
                                      const int32_t* __restrict__ slotidx,
                                      int32_t* __restrict__ bin_slot_start) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nbins;
         b += (int64_t)gridDim.x * blockDim.x)
        bin_slot_start[b] = bin_start[b] >= 0 ? slotidx[bin_start[b]] : -1;
}

__global__ void keys_to_i64_kernel(int64_t M, const uint64_t* __restrict__ k, int64_t* __restrict__ o) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        o[i] = (int64_t)k[i];
}

// ---------------------------------------------------------------------------------
// fused kernels around the oversampled FFT
// ---------------------------------------------------------------------------------
struct AxisPtrs {
    const double* sn[3];   // per-axis deapodization factors (double)
    const void* pb[3];     // per-axis phase_before angle, precision dtype
};

// grid[k] = x[k] * sn[k] * fwd_scale inside the Nd corner, 0 elsewhere
// (_nufft.py:1325-1331: `x * sn` then zero-padded FFT).  sn is the reference's dense
// array re-formed on the fly: ((s1*s2)*s3) in double, then cast (:737-748).
// One thread handles VEC consecutive cells of a grid row, so the row decode (integer
// divisions) is amortised and the stores are 16-32 bytes wide.
//
// SENSE = true (b2n_sense_fwd): ONE image is encoded by nbatch coil sensitivity maps,
// grid[b] = pad((x * smaps[b]) * sn): the coil images x * smaps[b] that a caller of the
// reference forms before NufftBase.fft (mrrt.operators' MRI_Operator, _nufft.py:3-5) are
// never written to memory.  Same rounding order as the unfused sequence.
template <typename T, int VEC, bool SENSE = false>
__global__ void pre_scale_pad_kernel(Geom g, AxisPtrs ax, T fwd_scale, int apply_scale,
                                     const cplx_t<T>* __restrict__ image,
                                     cplx_t<T>* __restrict__ grid, int nbatch,
                                     const cplx_t<T>* __restrict__ smaps = nullptr) {
    using C = cplx_t<T>;
    const int cpr = (g.K[0] + VEC - 1) / VEC;               // chunks per row
    const int64_t rows = (g.PK / g.K[0]) * nbatch;
    const int64_t total = rows * cpr;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cpr;
        const int k1s = (int)(idx - row * cpr) * VEC;
        const int k2 = g.ndim > 1 ? (int)(row % g.K[1]) : 0;
        const int64_t r2 = g.ndim > 1 ? row / g.K[1] : row;
        const int k3 = g.ndim > 2 ? (int)(r2 % g.K[2]) : 0;
        const int64_t b = g.ndim > 2 ? r2 / g.K[2] : r2;
        const bool row_in = (g.ndim < 2 || k2 < g.N[1]) && (g.ndim < 3 || k3 < g.N[2]);
        double s23 = 1.0;
        int64_t nrow = SENSE ? 0 : b * g.PN;
        if (row_in) {
            if (g.ndim > 1) nrow += (int64_t)k2 * g.N[0];
            if (g.ndim > 2) nrow += (int64_t)k3 * g.N[0] * g.N[1];
        }
        C v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) {
            const int k1 = k1s + e;
            v[e] = make_c<T>(0, 0);
            if (row_in && k1 < g.N[0]) {
                double s = ax.sn[0][k1];
                if (g.ndim > 1) s *= ax.sn[1][k2];
                if (g.ndim > 2) s *= ax.sn[2][k3];
                const T st = (T)s;
                C x = image[nrow + k1];
                if (SENSE) {
                    const C c = smaps[b * g.PN + nrow + k1];
                    x = make_c<T>(x.x * c.x - x.y * c.y, x.x * c.y + x.y * c.x);
                }
                v[e] = make_c<T>(x.x * st, x.y * st);
                if (apply_scale) { v[e].x *= fwd_scale; v[e].y *= fwd_scale; }
            }
        }
        (void)s23;
        C* dst = grid + row * g.K[0] + k1s;
        if (k1s + VEC <= g.K[0] && (g.K[0] % VEC) == 0) {
            // rows start VEC-aligned: vector stores
            if (sizeof(C) * VEC == 32) {
                ((int4*)dst)[0] = ((const int4*)v)[0];
                ((int4*)dst)[1] = ((const int4*)v)[1];
            } else if (sizeof(C) * VEC == 16) {
                ((int4*)dst)[0] = ((const int4*)v)[0];
            } else {
#pragma unroll
                for (int e = 0; e < VEC; e++) dst[e] = v[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++)
                if (k1s + e < g.K[0]) dst[e] = v[e];
        }
    }
}

__device__ __forceinline__ void sincos_t(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincos_t(double a, double* s, double* c) { sincos(a, s, c); }

// grid[k] *= exp(+-i*((p1[k1]+p2[k2])+p3[k3])): phase_before with the angle summed in
// the precision dtype in the reference's order (_nufft.py:703-715, :1370-1371, :1519-1520)
template <typename T, int VEC>
__global__ void phase_before_kernel(Geom g, AxisPtrs ax, int conj, cplx_t<T>* __restrict__ grid,
                                    int nbatch) {
    using C = cplx_t<T>;
    const int cpr = (g.K[0] + VEC - 1) / VEC;
    const int64_t rows = (g.PK / g.K[0]) * nbatch;
    const int64_t total = rows * cpr;
    const bool vec_ok = (g.K[0] % VEC) == 0;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cpr;
        const int k1s = (int)(idx - row * cpr) * VEC;
        const int k2 = g.ndim > 1 ? (int)(row % g.K[1]) : 0;
        const int64_t r2 = g.ndim > 1 ? row / g.K[1] : row;
        const int k3 = g.ndim > 2 ? (int)(r2 % g.K[2]) : 0;
        const T a2 = g.ndim > 1 ? ((const T*)ax.pb[1])[k2] : (T)0;
        const T a3 = g.ndim > 2 ? ((const T*)ax.pb[2])[k3] : (T)0;
        C* p = grid + row * g.K[0] + k1s;
        C v[VEC];
        const bool full = vec_ok && k1s + VEC <= g.K[0];
        if (full) {
            if (sizeof(C) * VEC == 32) {
                ((int4*)v)[0] = ((const int4*)p)[0];
                ((int4*)v)[1] = ((const int4*)p)[1];
            } else if (sizeof(C) * VEC == 16) {
                ((int4*)v)[0] = ((const int4*)p)[0];
            } else {
#pragma unroll
                for (int e = 0; e < VEC; e++) v[e] = p[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++)
                if (k1s + e < g.K[0]) v[e] = p[e];
        }
#pragma unroll
        for (int e = 0; e < VEC; e++) {
            const int k1 = min(k1s + e, g.K[0] - 1);
            T ang = ((const T*)ax.pb[0])[k1];
            if (g.ndim > 1) ang = ang + a2;
            if (g.ndim > 2) ang = ang + a3;
            T s, c;
            sincos_t(ang, &s, &c);
            if (conj) s = -s;
            const C u = v[e];
            v[e] = make_c<T>(u.x * c - u.y * s, u.x * s + u.y * c);
        }
        if (full) {
            if (sizeof(C) * VEC == 32) {
                ((int4*)p)[0] = ((const int4*)v)[0];
                ((int4*)p)[1] = ((const int4*)v)[1];
            } else if (sizeof(C) * VEC == 16) {
                ((int4*)p)[0] = ((const int4*)v)[0];
            } else {
#pragma unroll
                for (int e = 0; e < VEC; e++) p[e] = v[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++)
                if (k1s + e < g.K[0]) p[e] = v[e];
        }
    }
}

// image[n] = grid[n (corner)] * adj_scale * sn[n]   (_nufft.py:1560-1572)
template <typename T>
__global__ void post_crop_scale_kernel(Geom g, AxisPtrs ax, T adj_scale, int apply_scale,
                                       const cplx_t<T>* __restrict__ grid,
                                       cplx_t<T>* __restrict__ image, int nbatch) {
    using C = cplx_t<T>;
    const int64_t total = g.PN * nbatch;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / g.PN;
        int64_t r = idx - b * g.PN;
        const int n1 = (int)(r % g.N[0]);
        r /= g.N[0];
        const int n2 = g.ndim > 1 ? (int)(r % g.N[1]) : 0;
        const int n3 = g.ndim > 2 ? (int)(r / g.N[1]) : 0;
        double s = ax.sn[0][n1];
        int64_t k = n1;
        if (g.ndim > 1) { s *= ax.sn[1][n2]; k += (int64_t)n2 * g.K[0]; }
        if (g.ndim > 2) { s *= ax.sn[2][n3]; k += (int64_t)n3 * g.K[0] * g.K[1]; }
        const T st = (T)s;
        C v = grid[b * g.PK + k];
        if (apply_scale) { v.x *= adj_scale; v.y *= adj_scale; }
        image[idx] = make_c<T>(v.x * st, v.y * st);
    }
}

// image[n] = sum_b conj(smaps[b][n]) * (grid[b][n (corner)] * adj_scale * sn[n]):
// the crop/scale of every coil (as post_crop_scale_kernel) and the coil combination a
// caller of the reference performs after NufftBase.adj, without writing the coil images.
// Coils are summed in index order in the precision dtype, like the unfused sequence.
template <typename T>
__global__ void sense_crop_combine_kernel(Geom g, AxisPtrs ax, T adj_scale, int apply_scale,
                                          const cplx_t<T>* __restrict__ grid,
                                          const cplx_t<T>* __restrict__ smaps,
                                          cplx_t<T>* __restrict__ image, int ncoil) {
    using C = cplx_t<T>;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < g.PN;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx;
        const int n1 = (int)(r % g.N[0]);
        r /= g.N[0];
        const int n2 = g.ndim > 1 ? (int)(r % g.N[1]) : 0;
        const int n3 = g.ndim > 2 ? (int)(r / g.N[1]) : 0;
        double s = ax.sn[0][n1];
        int64_t k = n1;
        if (g.ndim > 1) { s *= ax.sn[1][n2]; k += (int64_t)n2 * g.K[0]; }
        if (g.ndim > 2) { s *= ax.sn[2][n3]; k += (int64_t)n3 * g.K[0] * g.K[1]; }
        const T st = (T)s;
        C acc = make_c<T>(0, 0);
        for (int b = 0; b < ncoil; b++) {
            C v = grid[b * g.PK + k];
            if (apply_scale) { v.x *= adj_scale; v.y *= adj_scale; }
            v = make_c<T>(v.x * st, v.y * st);
            const C c = smaps[b * g.PN + idx];
            acc.x += c.x * v.x + c.y * v.y;
            acc.y += c.x * v.y - c.y * v.x;
        }
        image[idx] = acc;
    }
}

// grid[b][k] *= kern[k] (complex): the spectrum of the Toeplitz kernel applied to the
// padded image's spectrum (b2n_grid_multiply); two cells per thread where the element
// count allows, the kernel array is read once per batch entry from L2.
template <typename T>
__global__ void grid_multiply_kernel(int64_t PK, const cplx_t<T>* __restrict__ kern,
                                     cplx_t<T>* __restrict__ grid, int nbatch) {
    using C = cplx_t<T>;
    const int64_t total = PK * nbatch;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const C t = kern[idx % PK];
        const C v = grid[idx];
        grid[idx] = make_c<T>(v.x * t.x - v.y * t.y, v.x * t.y + v.y * t.x);
    }
}

// ---------------------------------------------------------------------------------
// sparse mode: fixed-width rows (exactly prod(Jd) entries per sample)
// ---------------------------------------------------------------------------------
struct SparseSrc {
    const void* coef[3];      // [J_d, M] double or complex double, tap fastest
    const int32_t* kidx[3];   // [J_d, M]
};

// values formed as the reference forms them (_nufft.py:812-858): products in double in
// axis order, conjugate, optional row phase, cast to the matrix dtype
template <typename T, bool CT>
__global__ void build_ell_kernel(Geom g, SparseSrc src, const int32_t* __restrict__ perm,
                                 const double2* __restrict__ row_phase, int nnzr,
                                 typename WeightT<T, CT>::type* __restrict__ vals,
                                 int32_t* __restrict__ cols) {
    const int64_t total = g.M * nnzr;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / nnzr;
        int j = (int)(e - i * nnzr);
        const int64_t m = perm[i];
        int col = 0, kstride = 1;
        double2 u = make_double2(1.0, 0.0);
        for (int d = 0; d < g.ndim; d++) {
            const int jd = j % g.J[d];
            j /= g.J[d];
            const int64_t a = m * g.J[d] + jd;
            col += src.kidx[d][a] * kstride;
            kstride *= g.K[d];
            if (CT) {
                const double2 c = ((const double2*)src.coef[d])[a];
                u = d == 0 ? c : cmul(u, c);
            } else {
                const double c = ((const double*)src.coef[d])[a];
                u.x = d == 0 ? c : u.x * c;
            }
        }
        if constexpr (CT) {
            u.y = -u.y;
            if (row_phase != nullptr) u = cmul(u, row_phase[m]);
            vals[e] = make_c<T>((T)u.x, (T)u.y);
        } else {
            vals[e] = (T)u.x;
        }
        cols[e] = col;
    }
}

template <typename T, bool CT>
__global__ void __launch_bounds__(256)
spmv_fwd_kernel(Geom g, int nnzr, const typename WeightT<T, CT>::type* __restrict__ vals,
                const int32_t* __restrict__ cols, const int32_t* __restrict__ perm,
                const cplx_t<T>* __restrict__ grid, cplx_t<T>* __restrict__ out,
                const cplx_t<T>* __restrict__ phase_s, int nbatch) {
    using C = cplx_t<T>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < g.M; i += nwarp) {
        const int64_t dst = perm[i];
        for (int b = 0; b < nbatch; b++) {
            const C* __restrict__ gb = grid + (int64_t)b * g.PK;
            C acc = make_c<T>(0, 0);
            for (int j = lane; j < nnzr; j += 32) {
                const C p = w_mul(vals[i * nnzr + j], __ldg(gb + cols[i * nnzr + j]));
                acc.x += p.x;
                acc.y += p.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            }
            if (lane == 0) {
                if (phase_s != nullptr) acc = cmul(acc, phase_s[i]);
                out[(int64_t)b * g.M + dst] = acc;
            }
        }
    }
}

template <typename T, bool CT>
__global__ void __launch_bounds__(256)
spmv_adj_kernel(Geom g, int nnzr, const typename WeightT<T, CT>::type* __restrict__ vals,
                const int32_t* __restrict__ cols, const int32_t* __restrict__ perm,
                const cplx_t<T>* __restrict__ samples, cplx_t<T>* __restrict__ grid,
                const cplx_t<T>* __restrict__ phase_s, int nbatch) {
    using C = cplx_t<T>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < g.M; i += nwarp) {
        const int64_t srcm = perm[i];
        for (int b = 0; b < nbatch; b++) {
            C* __restrict__ gb = grid + (int64_t)b * g.PK;
            C f = samples[(int64_t)b * g.M + srcm];
            if (phase_s != nullptr) f = cmul_conj(f, phase_s[i]);
            for (int j = lane; j < nnzr; j += 32)
                atomic_add_c(gb + cols[i * nnzr + j], w_mul_conj(vals[i * nnzr + j], f));
        }
    }
}

// un-permute the ELL rows back to acquisition order for the copy-out used by tests
template <typename V>
__global__ void ell_unpermute_kernel(int64_t M, int nnzr, const int32_t* __restrict__ perm,
                                     const V* __restrict__ vals_s, const int32_t* __restrict__ cols_s,
                                     V* __restrict__ vals, int32_t* __restrict__ cols) {
    const int64_t total = M * nnzr;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / nnzr;
        const int j = (int)(e - i * nnzr);
        const int64_t m = perm[i];
        if (vals != nullptr) vals[m * nnzr + j] = vals_s[e];
        if (cols != nullptr) cols[m * nnzr + j] = cols_s[e];
    }
}

}  // namespace b2n
