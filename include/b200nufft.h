/*
 * b200nufft.h -- C ABI of libb200nufft.so (sm_100a), the B200-native replacement for
 * the native layer of mrrt.nufft's NUFFT hot path.
 *
 * Plain pointers and sizes only; no torch/cupy types.  Device arrays are exchanged by
 * raw device pointer (obtained from a DLPack / __cuda_array_interface__ producer on the
 * Python side).  Every entry point returns 0 on success or a B2N_E* code; the message
 * is available from b2n_last_error() (thread-local).  All launches go to the
 * cudaStream_t passed as `stream` (0 = legacy default stream).
 *
 * What each entry point replaces in the reference (paths under mrrt/nufft/):
 *   b2n_plan_create / _destroy   NufftBase._init_gpu (_nufft.py:362-390): cuFFT plan +
 *                                NVRTC RawKernel compilation (cuda/cupy.py:70-183) +
 *                                launch config (_cupy.py:16-44).  Nothing is compiled
 *                                at run time here.
 *   b2n_plan_set_tables          upload of NufftBase.h (_nufft.py:880-935); values are
 *                                computed on the host exactly as the reference does.
 *   b2n_plan_set_points          tm = omega/gam (_nufft.py:338-342) + the new bin-sort
 *                                step (no reference counterpart).
 *   b2n_interp_fwd               _interp{1,2,3}_table_forward (_nufft_table.pyx:29,215,
 *                                444) -> TYPE_interp{n}_table1_{real,complex}_forward
 *                                (c/nufft_table.template.c:43,97,538,623,710,822) and
 *                                the interp{n}_table*_per_GPUkernel RawKernels
 *                                (cuda/jinja/table_{1,2,3}d_forward.jinja).
 *   b2n_interp_adj               _interp{1,2,3}_table_adj (pyx:120,328,574) ->
 *                                *_adj / *_adj_inner (template.c:144-531, 924-1202) and
 *                                the *_per_adj_GPUkernel RawKernels
 *                                (cuda/jinja/table_{1,2,3}d_adjoint.jinja).
 *   b2n_nufft_fwd / b2n_nufft_adj  nufft_forward / nufft_adj (_nufft.py:1275-1397,
 *                                1459-1578): sn scaling, zero-pad, cuFFT, phase_before,
 *                                interpolation, phase_after / the mirror image.
 *   b2n_plan_set_sparse, b2n_spmv_fwd / _adj
 *                                _init_sparsemat's matrix (_nufft.py:751-877) held in
 *                                fixed-width (ELL) form, `obj.p * xk` (:1384) and
 *                                `obj.p.H * xk` (:1505) (scipy / cuSPARSE SpMV).
 */
#ifndef B200NUFFT_H
#define B200NUFFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2N_OK 0
#define B2N_EINVAL 1   /* bad argument (maps to ValueError) */
#define B2N_ECUDA 2    /* CUDA runtime / cuFFT failure (maps to RuntimeError) */
#define B2N_ESTATE 3   /* call sequence error, e.g. points not set (RuntimeError) */
#define B2N_ENONFINITE 4 /* NaN/Inf sample coordinate (ValueError) */

#define B2N_SINGLE 0   /* float / complex64 */
#define B2N_DOUBLE 1   /* double / complex128 */

#define B2N_COORD_TM 0     /* coordinates already in grid units (tm) */
#define B2N_COORD_OMEGA 1  /* radians; tm = omega / (2*pi/K) in the precision dtype */

typedef struct b2n_plan b2n_plan;

int b2n_version(void);
const char *b2n_last_error(void);

/* ndim in 1..3; Nd, Kd, Jd: ndim ints; L: table oversampling (table length J*L+1);
 * precision: B2N_SINGLE/B2N_DOUBLE; table_is_complex: 0 real table (phasing="real"),
 * 1 complex table (phasing="complex"); device: CUDA device ordinal. */
int b2n_plan_create(int ndim, const int *Nd, const int *Kd, const int *Jd, int L,
                    int precision, int table_is_complex, int device, b2n_plan **out);
int b2n_plan_destroy(b2n_plan *plan);

/* Integer options.  Layout options must precede b2n_plan_set_points; launch options may
 * change between transforms.  Every non-default value is a tested variant (tests/).
 *   layout:  "tile1","tile2","tile3" bin shape in grid cells; "tileb1".."tileb3" bin shape of
 *            the adjoint sort order (ignored by plans in column order); "chunk" max samples per
 *            forward work item; "order_b" 1 = also build the adjoint-side sort order;
 *            "adj_column" 1 (default) = 3-D real-table plans build the COLUMN order and the
 *            column records of csrc/spread_column.cuh (0 = the per-cell register-window kernel
 *            on the round-1 adjoint order); "win_maxslide" longest window slide in cells before a
 *            new window is started (0 = J-1; baked into the column records, so it is a layout
 *            option for those plans); "precomp_weights" 1 = plan-time interpolation weights
 *            (0 = table lookups in the kernels, the table staged in shared memory);
 *            "table_order" 1 = linear interpolation between table entries (default, the
 *            reference's CPU behaviour), 0 = the entry at floor((t-k)*L) (order 0 of the
 *            reference's GPU templates); "fwd_pair" 0/1/2 = same-cell sample pairs in the forward
 *            kernel off / automatic / on; "fwd_interleave" column-interleaved slot order;
 *            "slab_kglobal2", "slab_origin2" (slab plans, below).
 *   launch:  "force_generic" 1 = one-thread-per-sample kernels only; "use_tma" 0 = cooperative
 *            tile loads; "fwd_pitch" shared-memory row pitch of the forward tile (0 = auto);
 *            "slide_pts" samples per warp of the register-window adjoint kernels (0 = automatic:
 *            256, column kernel 512); "win_facew" -1 auto / 0 off / 1,2 face-weight staging of
 *            the per-cell 3-D window adjoint; "pruned_fft" 1 = skip the all-zero planes of the
 *            padded FFT; "own_fft3" 1 (default) = the axis-3 pass of the pruned FFT by the fused
 *            kernels of csrc/fft_axis3.cuh (zero padding, phase_before and crop inside the pass;
 *            compile-time radix schedule for K3 in {128,192,256,384,512,768,1024}), 2 = always the
 *            run-time schedule, 0 = cuFFT strided pass + phase kernel; "own_fft12" 1 (default) =
 *            also the two in-plane passes by that kernel, with the scale / zero-pad and crop /
 *            scale fused into the axis-1 pass (single 3-D volume, every Kd in the list above),
 *            0 = scale/pad kernel + cuFFT 2-D + crop kernel; "profile" 1 = CUDA events around the
 *            interpolation kernels; "sparse_mode" (set by the host for mode="sparse").
 *   read-only (b2n_plan_get_option): "last_fwd_kernel" (0 one thread per sample, 1 tiled),
 *            "last_adj_kernel" (0 one RED per tap, 3 3-D per-cell register window, 4 2-D register
 *            window, 5 3-D column-group register window), "n_items", "n_slots", "lib_calls",
 *            "axis3_fused" (1 when a fused axis-3 kernel serves this plan: b2n_axis3_fwd then
 *            ignores the planes >= Nd[2] of its input instead of requiring them to be zero),
 *            "inplane_own" (1 once the own in-plane passes have served a transform).
 *
 * Threading / streams: a plan is NOT re-entrant.  It owns one scratch grid and its cuFFT
 * handles are re-pointed at the stream of each call, so transforms on one plan must be issued
 * by one host thread at a time and on one stream (or be ordered by the caller with events).
 * Plan-time arrays (sorted points, weights) are complete when set_points / set_tables return.
 * Every entry point runs on the plan's device and restores the caller's current device. */
int b2n_plan_set_option(b2n_plan *plan, const char *name, long value);
long b2n_plan_get_option(b2n_plan *plan, const char *name);

/* h_host[d]: HOST pointer to Jd[d]*L+1 entries, real (or interleaved complex) in the
 * precision dtype. */
int b2n_plan_set_tables(b2n_plan *plan, const void *const *h_host);

/* Host-side plan constants for the full transforms:
 *  sn1d[d]      HOST double[Nd[d]]   per-axis deapodization factor (sn = product)
 *  pb_angle[d]  HOST real[Kd[d]] in the precision dtype: per-axis phase_before angle
 *               (2*pi/K*n_mid)*k, or NULL for a complex table (no phase_before)
 *  fwd_scale    multiplies the gridded spectrum in the forward transform (1/sqrt(prod K)
 *               if ortho else 1)
 *  adj_scale    multiplies the UNNORMALISED inverse FFT in the adjoint
 *               (adjoint_scalefactor, or adjoint_scalefactor/sqrt(prod K) if ortho) */
int b2n_plan_set_scaling(b2n_plan *plan, const double *const *sn1d,
                         const void *const *pb_angle, double fwd_scale,
                         double adj_scale);

/* coords_dev: DEVICE [M, ndim] column-major (axis d at offset d*M), precision dtype.
 * Computes tm (if kind==B2N_COORD_OMEGA), window origins, bin ids, sort keys, the stable
 * sort permutation and the sorted coordinate copy.  Synchronises `stream`. */
int b2n_plan_set_points(b2n_plan *plan, const void *coords_dev, int64_t M, int kind,
                        void *stream);

/* Optional per-sample unit phasor multiplied into the forward output and (conjugated)
 * into the adjoint input: phase_after (_nufft.py:717-724) or phase_shift (:898-903).
 * phase_dev: DEVICE complex[M] in the precision dtype, acquisition order; NULL clears. */
int b2n_plan_set_sample_phase(b2n_plan *plan, const void *phase_dev, void *stream);

int64_t b2n_plan_num_points(b2n_plan *plan);
int64_t b2n_plan_num_bins(b2n_plan *plan);
/* copy-outs for the bit-exact tests; any pointer may be NULL.
 * tm_dev real[M*ndim]; bin_ids_dev int32[M] (acquisition order); keys_dev int64[M]
 * (acquisition order); perm_dev int32[M] (sorted position -> acquisition index). */
int b2n_plan_get_points(b2n_plan *plan, void *tm_dev, int32_t *bin_ids_dev,
                        int64_t *keys_dev, int32_t *perm_dev, void *stream);

/* Forward "slots" (no reference counterpart; part of the trajectory preprocessing): one or
 * two sorted samples of the same grid cell that a thread of the forward kernel handles
 * together.  slot = (sorted position of the first sample << 1) | has_partner (the partner
 * is the next sorted position); pairs are formed greedily inside each run of equal sort
 * keys, and the slots of a bin are stored ordered by (rank inside the bin's axis-1 column,
 * column).  num_slots is 0 when the plan has none (1-D, complex table, option fwd_pair=0). */
int64_t b2n_plan_num_slots(b2n_plan *plan);
int b2n_plan_get_slots(b2n_plan *plan, uint32_t *slots_dev, void *stream);

/* Interpolation only (the reference's grid_only switches).
 * grid_dev: complex[prod(Kd) * nbatch], first axis fastest, batch slowest.
 * samples_dev: complex[M * nbatch], acquisition order, batch slowest.
 * apply_phase != 0 applies the sample phase set by b2n_plan_set_sample_phase. */
int b2n_interp_fwd(b2n_plan *plan, const void *grid_dev, void *samples_dev, int nbatch,
                   int apply_phase, void *stream);
/* zeroes grid_dev first, like the reference (template.c:965-966, :1107-1108) */
int b2n_interp_adj(b2n_plan *plan, const void *samples_dev, void *grid_dev, int nbatch,
                   int apply_phase, void *stream);

/* The two halves of a full transform, for callers that pipeline host transfers of sample
 * chunks against the interpolation (several plans over disjoint sample ranges sharing one
 * grid): image -> oversampled spectrum (x*sn, zero-pad, FFT, phase_before) into grid_dev,
 * and gridded spectrum -> image (conj phase_before, inverse FFT, crop, scale; grid_dev is
 * overwritten).  b2n_interp_adj accepts apply_phase | 2 to ACCUMULATE into grid_dev
 * instead of zeroing it first. */
int b2n_grid_fwd(b2n_plan *plan, const void *image_dev, void *grid_dev, int nbatch,
                 void *stream);
int b2n_grid_adj(b2n_plan *plan, void *grid_dev, void *image_dev, int nbatch, void *stream);

/* grid_dev[b][k] *= kernel_dev[k] (complex, precision dtype) for b < nbatch: the middle
 * step of the Toeplitz normal operator (the reference only hints at
 * it with return_psf, _nufft.py:1459,1495,1517-1518): on a plan with Kd = 2*Nd, unit
 * deapodization and no phase_before, b2n_grid_fwd (zero-pad + FFT), this call with the
 * spectrum of the point-spread function, and b2n_grid_adj (inverse FFT + crop + scale)
 * apply A^H W A without touching the non-uniform samples. */
int b2n_grid_multiply(b2n_plan *plan, void *grid_dev, const void *kernel_dev, int nbatch,
                      void *stream);

/* Full transforms.  image_dev: complex[prod(Nd) * nbatch] first axis fastest. */
int b2n_nufft_fwd(b2n_plan *plan, const void *image_dev, void *samples_dev, int nbatch,
                  void *stream);
int b2n_nufft_adj(b2n_plan *plan, const void *samples_dev, void *image_dev, int nbatch,
                  void *stream);

/* Coil-sensitivity encoding fused around the full transforms.
 * The reference's callers (mrrt.operators / mrrt.mri MRI_Operator, named at _nufft.py:3-5
 * and :200-201; not in its tree) multiply the image by the coil maps before NufftBase.fft
 * and conjugate-multiply-and-sum the coil images after NufftBase.adj; here those steps are
 * part of the scale/zero-pad kernel and of the crop/scale kernel, so the coil images never
 * travel through HBM.
 *   forward: samples[:, c] = NUFFT(image * smaps[:, c])                    c = 0..ncoil-1
 *   adjoint: image = sum_c conj(smaps[:, c]) * NUFFT^H(samples[:, c])
 * image_dev complex[prod(Nd)]; smaps_dev complex[prod(Nd) * ncoil] (coil slowest);
 * samples_dev complex[M * ncoil] (coil slowest); all in the precision dtype. */
int b2n_sense_fwd(b2n_plan *plan, const void *image_dev, const void *smaps_dev,
                  void *samples_dev, int ncoil, void *stream);
int b2n_sense_adj(b2n_plan *plan, const void *samples_dev, const void *smaps_dev,
                  void *image_dev, int ncoil, void *stream);

/* Sparse mode.  coef[d]: DEVICE [Jd[d], M] (tap fastest) per-axis coefficient, double
 * (real table) or interleaved complex double; kidx[d]: DEVICE int32 [Jd[d], M] wrapped
 * grid index per tap.  The ELL matrix (prod(Jd) entries per row) is formed on the device
 * as the reference forms it: products in double in axis order, conjugated, times the
 * optional per-row phasor row_phase (DEVICE complex double[M] or NULL), then cast to the
 * precision dtype (_nufft.py:812-858). */
int b2n_plan_set_sparse(b2n_plan *plan, const void *const *coef, const int32_t *const *kidx,
                        int64_t M, const void *row_phase, void *stream);
int64_t b2n_plan_sparse_nnz(b2n_plan *plan);
/* copy-out of the ELL arrays: vals (real or complex, precision dtype) [nnz], cols int32
 * [nnz], row-major (row m holds entries m*prod(Jd) ..). */
int b2n_plan_get_sparse(b2n_plan *plan, void *vals_dev, int32_t *cols_dev, void *stream);
int b2n_spmv_fwd(b2n_plan *plan, const void *grid_dev, void *samples_dev, int nbatch,
                 int apply_phase, void *stream);
int b2n_spmv_adj(b2n_plan *plan, const void *samples_dev, void *grid_dev, int nbatch,
                 int apply_phase, void *stream);

/* Staged 3-D transforms for slab-distributed operation (no
 * reference counterpart: the reference is single-device, _nufft.py:1333-1369, :1526-1559).
 * The oversampled FFT of b2n_grid_fwd / b2n_grid_adj is split so that an all-to-all can sit
 * between its in-plane part and its axis-3 part:
 *   b2n_planes_fwd   image planes [z0, z0+nz) (complex[Nd[0]*Nd[1]*nz]) -> x*sn, zero-pad to
 *                    Kd[0] x Kd[1] per plane, batched 2-D FFT -> planes_dev complex[Kd[0]*Kd[1]*nz]
 *   b2n_axis3_fwd    grid_dev complex[prod(Kd)] whose planes >= Nd[2] are zero: FFT along axis 3
 *                    (all Kd[0]*Kd[1] columns), then phase_before
 *   b2n_axis3_adj    conj(phase_before), unnormalised inverse FFT along axis 3
 *   b2n_planes_adj   planes_dev (overwritten): batched inverse 2-D FFT, crop, adj_scale * conj(sn)
 *                    -> image planes [z0, z0+nz)
 * The axis-3 calls run on a SLAB plan: a plan created with the slab's local Kd[1] and the
 * options "slab_kglobal2" (global Kd[1]) / "slab_origin2" (global index of local row 0) set
 * before b2n_plan_set_points; its coordinates, tables and window origins stay global (weights
 * are bit-identical to the single-device plan's), only grid rows are addressed locally, and
 * its pb_angle[1] holds the angles of the rows it owns.  b2n_plan_set_points rejects samples
 * whose window leaves those rows. */
int b2n_planes_fwd(b2n_plan *plan, const void *image_planes_dev, int z0, int nz,
                   void *planes_dev, void *stream);
int b2n_planes_adj(b2n_plan *plan, void *planes_dev, int z0, int nz, void *image_planes_dev,
                   void *stream);
int b2n_axis3_fwd(b2n_plan *plan, void *grid_dev, void *stream);
int b2n_axis3_adj(b2n_plan *plan, void *grid_dev, void *stream);

/* Exchange of grid rows between the plane stage and the axis-3 stage over PEER memory (one
 * process per GPU; every rank's slab grid mapped into every rank's address space, e.g. with
 * torch.distributed._symmetric_memory): the all-to-all, its pack / unpack passes and the halo
 * summation in one kernel each (csrc/slab_exchange.cuh).  `plan` is the plan with the GLOBAL
 * geometry; slab s holds rows (row0[s] + arange(nrows[s])) mod Kd[1] as complex
 * [Kd[2]][nrows[s]][Kd[0]] at peer_grids[s] (pointers valid on THIS device).
 *   b2n_slab_scatter  forward: rows of this rank's planes [z0, z0+nz) (planes_dev, complex
 *                     [nz][Kd[1]][Kd[0]]) stored into planes z0.. of every slab that holds them
 *   b2n_slab_gather   adjoint: planes_dev[z][k2][:] = sum over the slabs holding row k2 of their
 *                     plane z0+z (the halo rows of neighbouring slabs add up), fixed order
 * The caller orders the phases across ranks (barriers).  Needs Kd[0]*sizeof(complex) % 16 == 0
 * and at most 16 ranks. */
int b2n_slab_scatter(b2n_plan *plan, const void *planes_dev, int nz, int z0, int world,
                     void *const *peer_grids, const int *row0, const int *nrows, void *stream);
int b2n_slab_gather(b2n_plan *plan, void *planes_dev, int nz, int z0, int world,
                    void *const *peer_grids, const int *row0, const int *nrows, void *stream);

/* bytes of device memory owned by the plan */
int64_t b2n_plan_device_bytes(b2n_plan *plan);
/* number of OUR kernel launches issued so far (cuFFT executions and memsets are counted
 * separately: option "lib_calls") */
int64_t b2n_plan_launch_count(b2n_plan *plan);
/* With option "profile"=1 every interpolation kernel launch is bracketed by CUDA events
 * on the launching stream.  out[0..3] = {forward kernel total ms, launches, adjoint
 * kernel total ms, launches} since the previous call; synchronises those events. */
int b2n_plan_get_timing(b2n_plan *plan, double *out);

#ifdef __cplusplus
}
#endif
#endif /* B200NUFFT_H */
