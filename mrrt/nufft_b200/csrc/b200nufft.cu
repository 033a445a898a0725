// libb200nufft.so -- C ABI (include/b200nufft.h) over hand-written sm_100a kernels.
// Build: see mrrt/nufft_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
#include <cuda_runtime.h>
#include <cufft.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../../include/b200nufft.h"
#include "aux_kernels.cuh"
#include "fft_axis3.cuh"
#include "slab_exchange.cuh"
#include "common.cuh"
#include "dispatch.h"

using namespace b2n;

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                     \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess)                                                       \
            return fail(B2N_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define FFT(call)                                                                    \
    do {                                                                             \
        cufftResult r_ = (call);                                                     \
        if (r_ != CUFFT_SUCCESS)                                                     \
            return fail(B2N_ECUDA, std::string(#call) + ": cufft error " + std::to_string((int)r_)); \
    } while (0)

// Every entry point runs on the plan's device and restores the caller's current device on
// return (a plan on cuda:1 must not switch a single-process multi-GPU program to cuda:1).
struct DevGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DevGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;      // nothing to restore
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};
#define ON_DEVICE(p)                                                                  \
    DevGuard dev_guard_((p)->device);                                                 \
    if (dev_guard_.err != cudaSuccess)                                                \
        return fail(B2N_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(dev_guard_.err))

// ---------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------
struct b2n_plan {
    Geom g{};
    int precision = 0;
    int cplx_table = 0;
    int device = 0;
    int sm_count = 148;
    int max_smem = 0;            // opt-in shared memory per block (queried on first use)
    // options
    long opt_chunk = 4096;
    long opt_force_generic = 0;
    long opt_use_tma = 1;
    long opt_sparse_mode = 0;
    long opt_slide_pts = 0;      // samples per warp of the register-window adjoint kernels (0 = automatic:
                                 // 256, column-group kernel 512)
    long opt_order_b = 1;        // build the adjoint sort order (3-D register-window adjoint)
    long opt_fwd_pitch = 0;      // shared-memory row pitch of the forward tile (0 = automatic)
    long opt_fwd_pair = 1;       // tiled forward: same-cell sample pairs share one window pass (1 = auto)
    long opt_fwd_interleave = 1; // ... and the slots of a bin are ordered column-interleaved
    long opt_win_facew = -1;     // window adjoint: face-weight staging.  -1 = automatic: J <= 6 (measured at
                                 // J = 4 and 6): float 2 (5 CTAs/SM), double 1; 0 = off
    long opt_win_maxslide = 0;   // longest window slide in cells before a new window is started (0 = J-1)
    long opt_adj_column = 1;     // 3-D real-table adjoint: column sort order + column-group window kernel
    bool tile_user_set = false;
    bool tile_b_user_set = false;
    // tables
    void* d_tab[3] = {nullptr, nullptr, nullptr};
    bool tables_set = false;
    bool tables_equal = false;   // all axes share one table
    // scaling
    double* d_sn[3] = {nullptr, nullptr, nullptr};
    void* d_pb[3] = {nullptr, nullptr, nullptr};
    bool scaling_set = false;
    bool have_pb = false;
    double fwd_scale = 1.0, adj_scale = 1.0;
    // points
    bool points_set = false;
    void* d_tm = nullptr;        // [ndim][M] acquisition order
    void* d_tm_s = nullptr;      // [ndim][M] sorted
    uint64_t* d_keys = nullptr;  // acquisition order
    int32_t* d_bin_ids = nullptr;
    int32_t* d_perm = nullptr;
    int32_t* d_pt_ko = nullptr;  // [ndim][M] sorted: unwrapped window origins
    int32_t* d_pt_kw = nullptr;  // [ndim][M] sorted: wrapped window origins
    // second ("adjoint") sort order: same bins, cells ordered last-axis-fastest
    bool have_b = false;
    void* d_tm_sb = nullptr;
    int32_t* d_perm_b = nullptr;
    int32_t* d_pt_ko_b = nullptr;
    int32_t* d_pt_kw_b = nullptr;
    void* d_phase_sb = nullptr;
    // plan-time interpolation weights [sum(J)][M] for each sort order (real tables)
    void* d_wts = nullptr;
    void* d_wts_b = nullptr;
    void* d_col_rec = nullptr;   // column order: blocked plan-time records of the column-group adjoint kernel
    long opt_precomp = 1;
    void* d_phase_s = nullptr;   // sorted sample phase or null
    int64_t nbins = 0;
    // work items of the tiled forward kernel: (bin, start, count, pad)
    int4* d_items = nullptr;
    int64_t n_items = 0;
    // forward slots (pairs of same-cell samples) and the work items over them
    uint32_t* d_slots = nullptr;
    int32_t* d_slot_kw = nullptr;     // slot-ordered copies read by the packed forward kernel
    int32_t* d_slot_perm = nullptr;
    void* d_wts_f = nullptr;
    void* d_phase_f = nullptr;
    int4* d_items_f = nullptr;
    int64_t n_items_f = 0;
    int64_t n_slots = 0;
    // sparse
    void* d_ell_vals = nullptr;
    int32_t* d_ell_cols = nullptr;
    int nnzr = 0;
    bool sparse_set = false;
    // fft
    std::map<int, cufftHandle> fft_plans;
    bool fft_pruned_ready = false;
    cufftHandle fft_2d = 0, fft_1d = 0;
    std::map<int, cufftHandle> fft_planes;   // staged transforms: batched 2-D plans keyed by plane count
    bool fft_ax3_ready = false;              // ... and the strided 1-D plan along axis 3
    cufftHandle fft_ax3 = 0;
    long opt_pruned_fft = 1;
    long opt_own_fft3 = 1;       // pruned FFT: own axis-3 pass fused with the zero-padding, phase_before and the crop
    Axis3Plan ax3{};             // its radix schedule, and the K3-entry twiddle table (precision dtype)
    bool ax3_general = false;    // the run-time-schedule kernel can serve K3 (else only the fixed-schedule one)
    int ax3_state = 0;           // 0 = not prepared, 1 = ready, -1 = K3 not supported
    void* d_tw3 = nullptr;
    long opt_own_fft12 = 1;      // own in-plane passes with the scale/pad and crop/scale fused (needs own_fft3 = 1)
    int inplane_state = 0;       // 0 = not prepared, 1 = ready, -1 = not available for this grid
    void* d_tw12[2] = {nullptr, nullptr};   // twiddle tables of K1, K2 (may alias d_tw3 / each other)
    bool tw12_owned[2] = {false, false};
    void* d_work = nullptr;
    void* d_acc64 = nullptr;     // complex128 scratch grid of the float one-RED-per-tap adjoint
    size_t acc64_bytes = 0;
    size_t work_bytes = 0;
    int64_t dev_bytes = 0;
    int64_t launches = 0;        // our own kernels
    int64_t lib_calls = 0;       // cuFFT executions and memsets
    long opt_profile = 0;        // record CUDA events around the interpolation kernels
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_fwd, ev_adj;
    // window width the tiled forward / register-window adjoint kernels run at: J itself when
    // all axes share a width the kernels are compiled for, else the next such width (the extra
    // taps get zero plan-time weights); 0 = no such kernel (1-D, complex table, J > 8)
    int jk_fwd = 0, jk_adj = 0;
    int last_fwd_kernel = -1;   // 0 generic, 1 tiled
    int last_adj_kernel = -1;   // 0 generic, 3 3-D register window, 4 2-D register window
    std::map<void*, size_t> alloc_bytes;   // what dev_alloc handed out (device_bytes accounting)

    size_t real_size() const { return precision == B2N_SINGLE ? 4 : 8; }
    size_t cplx_size() const { return 2 * real_size(); }
};

static int dev_alloc(b2n_plan* p, void** ptr, size_t bytes) {
    if (bytes == 0) bytes = 16;
    CU(cudaMalloc(ptr, bytes));
    p->dev_bytes += (int64_t)bytes;
    p->alloc_bytes[*ptr] = bytes;
    return B2N_OK;
}
static void dev_free(b2n_plan* p, void* ptr) {
    if (!ptr) return;
    auto it = p->alloc_bytes.find(ptr);
    if (it != p->alloc_bytes.end()) {
        p->dev_bytes -= (int64_t)it->second;
        p->alloc_bytes.erase(it);
    }
    cudaFree(ptr);
}

// plan-time scratch buffers: freed on every exit path
struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() {
        for (void* q : ptrs) cudaFree(q);
    }
    template <typename P> cudaError_t alloc(P** out, size_t bytes) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
        if (e == cudaSuccess) ptrs.push_back(q);
        *out = (P*)q;
        return e;
    }
};

static int grid_for(int64_t n, int block, int sm_count, int per_sm = 16) {
    int64_t b = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int b2n_version(void) { return 100; }
extern "C" const char* b2n_last_error(void) { return g_err.c_str(); }

// window width the tiled forward / register-window adjoint kernels run at (see b2n_plan)
static void choose_widths(b2n_plan* p) {
    const Geom& g = p->g;
    p->jk_fwd = p->jk_adj = 0;
    if (g.ndim < 2) return;
    int jmax = 0;
    bool equal = true;
    for (int d = 0; d < g.ndim; d++) {
        jmax = g.J[d] > jmax ? g.J[d] : jmax;
        equal = equal && g.J[d] == g.J[0];
    }
    const int even = jmax <= 4 ? 4 : (jmax + 1) / 2 * 2;       // forward, 2-D adjoint: 4, 6, 8
    p->jk_fwd = jmax <= 8 ? even : 0;
    // 3-D adjoint windows exist for every width 4..8
    p->jk_adj = g.ndim == 3 ? (jmax <= 8 ? (jmax < 4 ? 4 : (equal ? jmax : even)) : 0) : p->jk_fwd;
    for (int d = 0; d < g.ndim; d++) {
        if (g.K[d] < p->jk_fwd) p->jk_fwd = 0;
        if (g.K[d] < p->jk_adj) p->jk_adj = 0;
    }
}

// whether the 3-D adjoint runs the column-group kernel (and the plan builds the column order)
static bool column_mode(const b2n_plan* p, int* GB, int* GC) {
    const Geom& g = p->g;
    int FB = 0, FC = 0;
    if (g.ndim != 3 || p->cplx_table || !p->opt_precomp || !p->opt_order_b || !p->opt_adj_column) return false;
    if (!column_shape(p->jk_adj, &FB, &FC, GB, GC)) return false;
    // (the records pack the origin along axis 3 into 16 bits)
    return g.K[0] >= FB && g.K[1] >= FC && g.K[2] >= p->jk_adj && g.K[2] < 32768;
}

static void default_tiles(b2n_plan* p) {
    Geom& g = p->g;
    if (!p->tile_user_set) {
        if (g.ndim == 1) { g.tile[0] = 1024; }
        if (g.ndim == 2) { g.tile[0] = 32; g.tile[1] = 32; }
        if (g.ndim == 3) { g.tile[0] = 16; g.tile[1] = 8; g.tile[2] = 8; }
    }
    int GB = 1, GC = 1;
    g.colmode = column_mode(p, &GB, &GC) ? 1 : 0;
    if (g.colmode) {
        // column order: one bin per group of GB x GC grid columns, whole last axis
        g.tile_b[0] = GB; g.tile_b[1] = GC; g.tile_b[2] = g.K[2];
    } else if (!p->tile_b_user_set) {
        // adjoint order: long bins along the LAST axis (the one the register windows slide along)
        g.tile_b[0] = 16; g.tile_b[1] = g.ndim == 2 ? 64 : 16; g.tile_b[2] = 64;
    }
    for (int d = 0; d < 3; d++) {
        if (d >= g.ndim) g.tile_b[d] = 1;
        if (g.tile_b[d] > g.K[d]) g.tile_b[d] = g.K[d];
        if (g.tile_b[d] < 1) g.tile_b[d] = 1;
        g.nbin_b[d] = (g.K[d] + g.tile_b[d] - 1) / g.tile_b[d];
    }
    g.cells_per_tile = 1;
    for (int d = 0; d < 3; d++) {
        if (d >= g.ndim) g.tile[d] = 1;
        if (g.tile[d] > g.K[d]) g.tile[d] = g.K[d];
        if (g.tile[d] < 1) g.tile[d] = 1;
        g.nbin[d] = (g.K[d] + g.tile[d] - 1) / g.tile[d];
        g.cells_per_tile *= g.tile[d];
    }
}

extern "C" int b2n_plan_create(int ndim, const int* Nd, const int* Kd, const int* Jd, int L,
                               int precision, int table_is_complex, int device,
                               b2n_plan** out) {
    if (out == nullptr) return fail(B2N_EINVAL, "out is NULL");
    if (ndim < 1 || ndim > 3) return fail(B2N_EINVAL, "dimensions > 3 not implemented");
    if (precision != B2N_SINGLE && precision != B2N_DOUBLE)
        return fail(B2N_EINVAL, "precision must be B2N_SINGLE or B2N_DOUBLE");
    if (L < 1) return fail(B2N_EINVAL, "L must be >= 1");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(B2N_EINVAL, "bad device ordinal");
    DevGuard dev_guard_(device);
    if (dev_guard_.err != cudaSuccess)
        return fail(B2N_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(dev_guard_.err));
    b2n_plan* p = new b2n_plan();
    p->precision = precision;
    p->cplx_table = table_is_complex ? 1 : 0;
    p->device = device;
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    Geom& g = p->g;
    g.ndim = ndim;
    g.L = L;
    g.order = 1;
    g.PK = 1;
    g.PN = 1;
    for (int d = 0; d < 3; d++) {
        g.N[d] = d < ndim ? Nd[d] : 1;
        g.K[d] = d < ndim ? Kd[d] : 1;
        g.J[d] = d < ndim ? Jd[d] : 1;
        g.Kg[d] = g.K[d];
        g.korg[d] = 0;
        if (g.N[d] < 1 || g.K[d] < g.N[d] || g.J[d] < 1 || g.J[d] > kMaxJ) {
            delete p;
            return fail(B2N_EINVAL, "need 1 <= N <= K and 1 <= J <= 16 on every axis");
        }
        g.ncenter[d] = (g.J[d] * L) / 2;
        g.tlen[d] = g.J[d] * L + 1;
        g.PK *= g.K[d];
        g.PN *= g.N[d];
    }
    if (g.PK >= ((int64_t)1 << 31)) {
        delete p;
        return fail(B2N_EINVAL, "prod(Kd) must be < 2^31");
    }
    choose_widths(p);
    default_tiles(p);
    *out = p;
    return B2N_OK;
}

static void free_points(b2n_plan* p) {
    dev_free(p, p->d_tm); dev_free(p, p->d_tm_s); dev_free(p, p->d_keys); dev_free(p, p->d_bin_ids);
    dev_free(p, p->d_perm); dev_free(p, p->d_phase_s); dev_free(p, p->d_items);
    dev_free(p, p->d_slots); dev_free(p, p->d_items_f);
    dev_free(p, p->d_slot_kw); dev_free(p, p->d_slot_perm); dev_free(p, p->d_phase_f);
    p->d_slots = nullptr; p->d_items_f = nullptr; p->n_items_f = 0; p->n_slots = 0;
    p->d_slot_kw = p->d_slot_perm = nullptr; p->d_wts_f = p->d_phase_f = nullptr;
    dev_free(p, p->d_pt_ko); dev_free(p, p->d_pt_kw);
    p->d_pt_ko = p->d_pt_kw = nullptr;
    dev_free(p, p->d_tm_sb); dev_free(p, p->d_perm_b); dev_free(p, p->d_pt_ko_b); dev_free(p, p->d_pt_kw_b);
    dev_free(p, p->d_phase_sb);
    dev_free(p, p->d_wts); dev_free(p, p->d_wts_b); dev_free(p, p->d_wts_f); dev_free(p, p->d_col_rec);
    p->d_wts = p->d_wts_b = p->d_wts_f = p->d_col_rec = nullptr;
    p->d_tm_sb = p->d_phase_sb = nullptr;
    p->d_perm_b = p->d_pt_ko_b = p->d_pt_kw_b = nullptr;
    p->have_b = false;
    p->d_tm = p->d_tm_s = p->d_phase_s = nullptr;
    p->d_keys = nullptr; p->d_bin_ids = nullptr; p->d_perm = nullptr; p->d_items = nullptr;
    p->points_set = false;
}

extern "C" int b2n_plan_destroy(b2n_plan* p) {
    if (p == nullptr) return B2N_OK;
    DevGuard dev_guard_(p->device);
    for (auto& kv : p->fft_plans) cufftDestroy(kv.second);
    if (p->fft_pruned_ready) { cufftDestroy(p->fft_2d); cufftDestroy(p->fft_1d); }
    for (auto& kv : p->fft_planes) cufftDestroy(kv.second);
    if (p->fft_ax3_ready) cufftDestroy(p->fft_ax3);
    dev_free(p, p->d_tw3);
    for (int d = 0; d < 2; d++) if (p->tw12_owned[d]) dev_free(p, p->d_tw12[d]);
    free_points(p);
    for (int d = 0; d < 3; d++) {
        bool dup = false;
        for (int e = 0; e < d; e++) dup = dup || (p->d_tab[e] == p->d_tab[d]);
        if (!dup) dev_free(p, p->d_tab[d]);
        dev_free(p, p->d_sn[d]);
        dev_free(p, p->d_pb[d]);
    }
    dev_free(p, p->d_ell_vals);
    dev_free(p, p->d_ell_cols);
    dev_free(p, p->d_work);
    dev_free(p, p->d_acc64);
    delete p;
    return B2N_OK;
}

extern "C" int b2n_plan_set_option(b2n_plan* p, const char* name, long value) {
    if (p == nullptr || name == nullptr) return fail(B2N_EINVAL, "NULL argument");
    std::string n(name);
    if (n == "tile1" || n == "tile2" || n == "tile3") {
        if (p->points_set) return fail(B2N_ESTATE, "tile options must precede set_points");
        if (value < 1) return fail(B2N_EINVAL, "tile must be >= 1");
        p->g.tile[n[4] - '1'] = (int)value;
        p->tile_user_set = true;
        default_tiles(p);
    } else if (n == "tileb1" || n == "tileb2" || n == "tileb3") {
        if (p->points_set) return fail(B2N_ESTATE, "tile options must precede set_points");
        if (value < 1) return fail(B2N_EINVAL, "tile must be >= 1");
        p->g.tile_b[n[5] - '1'] = (int)value;
        p->tile_b_user_set = true;
        default_tiles(p);
    } else if (n == "slab_kglobal2" || n == "slab_origin2") {
        // slab plan (SlabShardedNufft): this plan holds rows [origin, origin + Kd[1]) (mod the
        // global size) of axis 2 of a larger periodic grid; coordinates stay global
        if (p->points_set) return fail(B2N_ESTATE, "slab options must precede set_points");
        if (p->g.ndim != 3) return fail(B2N_EINVAL, "slab plans are 3-D");
        if (n == "slab_kglobal2") {
            if (value < p->g.K[1]) return fail(B2N_EINVAL, "slab_kglobal2 must be >= Kd[1]");
            p->g.Kg[1] = (int)value;
        } else {
            if (value < 0 || value >= p->g.Kg[1]) return fail(B2N_EINVAL, "slab_origin2 must be in [0, slab_kglobal2)");
            p->g.korg[1] = (int)value;
        }
    } else if (n == "table_order") {
        // 1 = linear interpolation of the table (default), 0 = the entry at floor((t-k)*L)
        if (value != 0 && value != 1) return fail(B2N_EINVAL, "table_order must be 0 or 1");
        if (p->points_set) return fail(B2N_ESTATE, "table_order must precede set_points");
        p->g.order = (int)value;
    } else if (n == "chunk") {
        if (value < 32) return fail(B2N_EINVAL, "chunk must be >= 32");
        p->opt_chunk = value;
    } else if (n == "force_generic") {
        p->opt_force_generic = value;
    } else if (n == "use_tma") {
        p->opt_use_tma = value;
    } else if (n == "sparse_mode") {
        p->opt_sparse_mode = value;
    } else if (n == "precomp_weights") {
        if (p->points_set) return fail(B2N_ESTATE, "precomp_weights must precede set_points");
        p->opt_precomp = value;
        default_tiles(p);
    } else if (n == "fwd_pitch") {
        if (value < 0 || value > 127) return fail(B2N_EINVAL, "fwd_pitch must be in 0..127");
        p->opt_fwd_pitch = value;
    } else if (n == "pruned_fft") {
        p->opt_pruned_fft = value;
    } else if (n == "own_fft12") {
        p->opt_own_fft12 = value;
    } else if (n == "own_fft3") {
        p->opt_own_fft3 = value;
    } else if (n == "fwd_pair") {
        if (value < 0 || value > 2) return fail(B2N_EINVAL, "fwd_pair must be 0 (off), 1 (auto) or 2 (on)");
        p->opt_fwd_pair = value;
    } else if (n == "fwd_interleave") {
        p->opt_fwd_interleave = value;
    } else if (n == "win_maxslide") {
        if (value < 0 || value > 15) return fail(B2N_EINVAL, "win_maxslide must be in 0..15");
        if (p->d_col_rec != nullptr)
            return fail(B2N_ESTATE, "win_maxslide is baked into the column records: set it before the points and tables");
        p->opt_win_maxslide = value;
    } else if (n == "win_facew") {
        if (value < -1 || value > 2) return fail(B2N_EINVAL, "win_facew must be -1 (auto) or 0..2");
        p->opt_win_facew = value;
    } else if (n == "order_b") {
        if (p->points_set) return fail(B2N_ESTATE, "order_b must precede set_points");
        p->opt_order_b = value;
        default_tiles(p);
    } else if (n == "adj_column") {
        if (p->points_set) return fail(B2N_ESTATE, "adj_column must precede set_points");
        p->opt_adj_column = value;
        default_tiles(p);
    } else if (n == "profile") {
        p->opt_profile = value;
    } else if (n == "slide_pts") {
        if (value != 0 && value < 32) return fail(B2N_EINVAL, "slide_pts must be 0 (automatic) or >= 32");
        p->opt_slide_pts = value;
    } else {
        return fail(B2N_EINVAL, "unknown option " + n);
    }
    return B2N_OK;
}

template <typename T> static bool axis3_fused(b2n_plan* p, int nbatch);

extern "C" long b2n_plan_get_option(b2n_plan* p, const char* name) {
    if (p == nullptr || name == nullptr) return -1;
    std::string n(name);
    if (n == "axis3_fused") {       // 1 when the fused axis-3 kernel (not cuFFT) serves this plan
        DevGuard dev_guard_(p->device);
        return p->precision == B2N_SINGLE ? (axis3_fused<float>(p, 1) ? 1 : 0)
                                          : (axis3_fused<double>(p, 1) ? 1 : 0);
    }
    if (n == "tile1") return p->g.tile[0];
    if (n == "tile2") return p->g.tile[1];
    if (n == "tile3") return p->g.tile[2];
    if (n == "chunk") return p->opt_chunk;
    if (n == "table_order") return p->g.order;
    if (n == "slab_kglobal2") return p->g.Kg[1];
    if (n == "slab_origin2") return p->g.korg[1];
    if (n == "force_generic") return p->opt_force_generic;
    if (n == "use_tma") return p->opt_use_tma;
    if (n == "sparse_mode") return p->opt_sparse_mode;
    if (n == "slide_pts") return p->opt_slide_pts;
    if (n == "profile") return p->opt_profile;
    if (n == "precomp_weights") return (p->d_wts != nullptr || p->d_wts_f != nullptr) ? 1 : 0;
    if (n == "lib_calls") return (long)p->lib_calls;
    if (n == "n_items") return (long)p->n_items;
    if (n == "n_slots") return (long)p->n_slots;
    if (n == "own_fft3") return (long)p->opt_own_fft3;
    if (n == "own_fft12") return (long)p->opt_own_fft12;
    if (n == "inplane_own") return p->inplane_state == 1 ? 1 : 0;
    if (n == "pruned_fft") return (long)p->opt_pruned_fft;
    if (n == "win_facew") return (long)p->opt_win_facew;
    if (n == "win_maxslide") return (long)p->opt_win_maxslide;
    if (n == "fwd_pair") return (long)p->opt_fwd_pair;
    if (n == "fwd_interleave") return (long)p->opt_fwd_interleave;
    if (n == "last_fwd_kernel") return p->last_fwd_kernel;
    if (n == "last_adj_kernel") return p->last_adj_kernel;
    if (n == "adj_column") return p->g.colmode;
    return -1;
}

static int ensure_weights(b2n_plan* p, cudaStream_t st);

extern "C" int b2n_plan_set_tables(b2n_plan* p, const void* const* h_host) {
    if (p == nullptr || h_host == nullptr) return fail(B2N_EINVAL, "NULL argument");
    ON_DEVICE(p);
    const Geom& g = p->g;
    const size_t esz = p->cplx_table ? p->cplx_size() : p->real_size();
    for (int d = 0; d < g.ndim; d++)
        if (h_host[d] == nullptr) return fail(B2N_EINVAL, "h1 size problem");
    // identical tables across axes are stored once (lets kernels keep a single copy
    // in shared memory)
    p->tables_equal = true;
    for (int d = 1; d < g.ndim; d++)
        if (g.tlen[d] != g.tlen[0] || memcmp(h_host[d], h_host[0], esz * g.tlen[0]) != 0)
            p->tables_equal = false;
    for (int d = 0; d < g.ndim; d++) {
        if (d > 0 && p->tables_equal) {
            p->d_tab[d] = p->d_tab[0];
            continue;
        }
        if (p->d_tab[d] == nullptr) {
            int rc = dev_alloc(p, &p->d_tab[d], esz * g.tlen[d]);
            if (rc) return rc;
        }
        CU(cudaMemcpy(p->d_tab[d], h_host[d], esz * g.tlen[d], cudaMemcpyHostToDevice));
    }
    p->tables_set = true;
    dev_free(p, p->d_wts); dev_free(p, p->d_wts_b); dev_free(p, p->d_wts_f); dev_free(p, p->d_col_rec);
    p->d_wts = p->d_wts_b = p->d_wts_f = p->d_col_rec = nullptr;
    int rc = ensure_weights(p, (cudaStream_t)0);
    if (rc) return rc;
    CU(cudaStreamSynchronize((cudaStream_t)0));
    return B2N_OK;
}

extern "C" int b2n_plan_set_scaling(b2n_plan* p, const double* const* sn1d,
                                    const void* const* pb_angle, double fwd_scale,
                                    double adj_scale) {
    if (p == nullptr || sn1d == nullptr) return fail(B2N_EINVAL, "NULL argument");
    ON_DEVICE(p);
    const Geom& g = p->g;
    for (int d = 0; d < g.ndim; d++) {
        if (sn1d[d] == nullptr) return fail(B2N_EINVAL, "sn1d axis missing");
        if (p->d_sn[d] == nullptr) {
            int rc = dev_alloc(p, (void**)&p->d_sn[d], sizeof(double) * g.N[d]);
            if (rc) return rc;
        }
        CU(cudaMemcpy(p->d_sn[d], sn1d[d], sizeof(double) * g.N[d], cudaMemcpyHostToDevice));
    }
    p->have_pb = pb_angle != nullptr;
    if (p->have_pb) {
        for (int d = 0; d < g.ndim; d++) {
            if (pb_angle[d] == nullptr) return fail(B2N_EINVAL, "pb_angle axis missing");
            if (p->d_pb[d] == nullptr) {
                int rc = dev_alloc(p, &p->d_pb[d], p->real_size() * g.K[d]);
                if (rc) return rc;
            }
            CU(cudaMemcpy(p->d_pb[d], pb_angle[d], p->real_size() * g.K[d], cudaMemcpyHostToDevice));
        }
    }
    p->fwd_scale = fwd_scale;
    p->adj_scale = adj_scale;
    p->scaling_set = true;
    return B2N_OK;
}

// ---------------------------------------------------------------------------------
// points: tm, keys, stable sort, sorted copies, work items
// ---------------------------------------------------------------------------------
template <typename T>
static int set_points_t(b2n_plan* p, const void* coords, int64_t M, int kind, cudaStream_t st) {
    Geom& g = p->g;
    g.M = M;
    int rc;
    const size_t rs = sizeof(T);
    if ((rc = dev_alloc(p, &p->d_tm, rs * M * g.ndim))) return rc;
    if ((rc = dev_alloc(p, &p->d_tm_s, rs * M * g.ndim))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_keys, sizeof(uint64_t) * M))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_bin_ids, sizeof(int32_t) * M))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_perm, sizeof(int32_t) * M))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_pt_ko, sizeof(int32_t) * M * g.ndim))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_pt_kw, sizeof(int32_t) * M * g.ndim))) return rc;
    p->nbins = (int64_t)g.nbin[0] * g.nbin[1] * g.nbin[2];
    if (M == 0) {
        p->n_items = 0;
        p->points_set = true;
        return B2N_OK;
    }
    Scratch scratch;
    // the adjoint order is used by the 2-D and 3-D register-window kernels
    const bool want_b = g.ndim >= 2 && p->opt_order_b && (!p->cplx_table || p->opt_precomp);
    uint64_t* keys_b = nullptr;
    if (want_b) {
        if ((rc = dev_alloc(p, (void**)&p->d_perm_b, sizeof(int32_t) * M))) return rc;
        if (!g.colmode) {
            // (column order: the sorted coordinates and window origins are only needed while the
            // column records are built; build_weights_t makes them in scratch memory)
            if ((rc = dev_alloc(p, &p->d_tm_sb, rs * M * g.ndim))) return rc;
            if ((rc = dev_alloc(p, (void**)&p->d_pt_ko_b, sizeof(int32_t) * M * g.ndim))) return rc;
            if ((rc = dev_alloc(p, (void**)&p->d_pt_kw_b, sizeof(int32_t) * M * g.ndim))) return rc;
        }
        CU(scratch.alloc(&keys_b, sizeof(uint64_t) * M));
    }
    uint64_t* keys_s = nullptr;
    int32_t* iota = nullptr;
    int* flag = nullptr;
    int32_t* bin_start = nullptr;
    void* tmp = nullptr;
    CU(scratch.alloc(&keys_s, sizeof(uint64_t) * M));
    CU(scratch.alloc(&iota, sizeof(int32_t) * M));
    CU(scratch.alloc(&flag, sizeof(int)));
    CU(cudaMemsetAsync(flag, 0, sizeof(int), st));
    Gam<T> gam;
    for (int d = 0; d < 3; d++) gam.g[d] = (T)(2.0 * M_PI / (double)g.Kg[d]);
    int jmax = 1;          // slab plans: widest window any kernel will run over these samples
    for (int d = 0; d < g.ndim; d++) jmax = g.J[d] > jmax ? g.J[d] : jmax;
    if (p->opt_precomp) jmax = std::max(jmax, std::max(p->jk_fwd, p->jk_adj));
    prep_points_kernel<T><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
        g, gam, kind, jmax, (const T*)coords, (T*)p->d_tm, p->d_keys, keys_b, p->d_bin_ids, iota, flag);
    CU(cudaGetLastError());
    // stable LSD radix sort over just the significant key bits
    uint64_t maxkey = (uint64_t)p->nbins * (uint64_t)g.cells_per_tile;
    {
        uint64_t mb = 1;
        for (int d = 0; d < 3; d++) mb *= (uint64_t)g.nbin_b[d] * (uint64_t)g.tile_b[d];
        if (mb > maxkey) maxkey = mb;
    }
    int bits = 1;
    while (bits < 64 && (maxkey >> bits) != 0) bits++;
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, p->d_keys, keys_s, iota, p->d_perm,
                                       M, 0, bits, st));
    CU(scratch.alloc(&tmp, tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, p->d_keys, keys_s, iota, p->d_perm, M,
                                       0, bits, st));
    gather_points_kernel<T><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
        g.ndim, M, p->d_perm, (const T*)p->d_tm, (T*)p->d_tm_s);
    CU(cudaGetLastError());
    point_windows_kernel<T><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
        g, (const T*)p->d_tm_s, p->d_pt_ko, p->d_pt_kw);
    CU(cudaGetLastError());
    if (want_b) {
        uint64_t* keys_bs = nullptr;
        CU(scratch.alloc(&keys_bs, sizeof(uint64_t) * M));
        CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_b, keys_bs, iota, p->d_perm_b, M, 0,
                                           bits, st));
        if (!g.colmode) {
            gather_points_kernel<T><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
                g.ndim, M, p->d_perm_b, (const T*)p->d_tm, (T*)p->d_tm_sb);
            CU(cudaGetLastError());
            point_windows_kernel<T><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
                g, (const T*)p->d_tm_sb, p->d_pt_ko_b, p->d_pt_kw_b);
            CU(cudaGetLastError());
        }
        CU(cudaStreamSynchronize(st));
        p->have_b = true;
    }
    // bin boundaries -> host -> work items of at most opt_chunk samples
    CU(scratch.alloc(&bin_start, sizeof(int32_t) * p->nbins));
    CU(cudaMemsetAsync(bin_start, 0xff, sizeof(int32_t) * p->nbins, st));
    bin_start_kernel<<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(M, g.cells_per_tile, keys_s,
                                                                    bin_start);
    CU(cudaGetLastError());
    std::vector<int32_t> hstart(p->nbins);
    int hflag = 0;
    CU(cudaMemcpyAsync(hstart.data(), bin_start, sizeof(int32_t) * p->nbins,
                       cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (hflag) {
        free_points(p);
        if (hflag == 2) return fail(B2N_EINVAL, "a sample's window leaves the rows held by this slab plan");
        return fail(B2N_ENONFINITE, "omega contains NaN or Inf");
    }
    std::vector<int4> items;
    int64_t next = M;
    std::vector<int64_t> bend(p->nbins, 0);
    for (int64_t b = p->nbins - 1; b >= 0; b--) {
        if (hstart[b] >= 0) {
            bend[b] = next;
            next = hstart[b];
        }
    }
    for (int64_t b = 0; b < p->nbins; b++) {
        if (hstart[b] < 0) continue;
        int64_t s = hstart[b], e = bend[b];
        while (s < e) {
            int64_t c = e - s < p->opt_chunk ? e - s : p->opt_chunk;
            items.push_back(make_int4((int)b, (int)s, (int)c, 0));
            s += c;
        }
    }
    p->n_items = (int64_t)items.size();
    if ((rc = dev_alloc(p, (void**)&p->d_items, sizeof(int4) * (items.size() + 1)))) return rc;
    CU(cudaMemcpy(p->d_items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice));
    // forward slots: same-cell sample pairs share one pass over the window (tiled forward)
    // fwd_pair: 0 off, 2 always, 1 (default) automatic = 3-D single precision with J >= 6, and only when
    // at least 10 % of the samples are partners.  Measured: the paired kernel doubles the
    // FMAs per shared-memory read, which pays where the forward kernel is shared-memory
    // bound (3-D float: 3.17 -> 2.30 ms) and costs where it is not (double: FP64 pipe,
    // 6.3 -> 8.0 ms; 2-D: per-slot overhead, configs[3] 0.157 -> 0.197 ms; J = 4: 64 taps
    // per sample are too few, configs[2] 0.176 -> 0.182 ms).
    const bool want_pairs = g.ndim >= 2 && !p->cplx_table &&
                            (p->opt_fwd_pair == 2 ||
                             (p->opt_fwd_pair == 1 && g.ndim == 3 && p->precision == B2N_SINGLE &&
                              p->jk_fwd >= 6));
    if (want_pairs) {
        int32_t *head = nullptr, *isslot = nullptr, *slotidx = nullptr, *bss = nullptr;
        CU(scratch.alloc(&head, sizeof(int32_t) * M));
        CU(scratch.alloc(&isslot, sizeof(int32_t) * M));
        CU(scratch.alloc(&slotidx, sizeof(int32_t) * M));
        CU(scratch.alloc(&bss, sizeof(int32_t) * p->nbins));
        slot_heads_kernel<<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(M, keys_s, head);
        CU(cudaGetLastError());
        size_t sb1 = 0, sb2 = 0;
        CU(cub::DeviceScan::InclusiveScan(nullptr, sb1, head, head, cub::Max(), (int)M, st));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, sb2, isslot, slotidx, (int)M, st));
        void* stmp = nullptr;
        CU(scratch.alloc(&stmp, sb1 > sb2 ? sb1 : sb2));
        CU(cub::DeviceScan::InclusiveScan(stmp, sb1, head, head, cub::Max(), (int)M, st));
        slot_flags_kernel<<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(M, head, isslot);
        CU(cudaGetLastError());
        CU(cub::DeviceScan::ExclusiveSum(stmp, sb2, isslot, slotidx, (int)M, st));
        int32_t last_idx = 0, last_flag = 0;
        CU(cudaMemcpyAsync(&last_idx, slotidx + (M - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&last_flag, isslot + (M - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        bin_slot_start_kernel<<<grid_for(p->nbins, 256, p->sm_count), 256, 0, st>>>(
            p->nbins, bin_start, slotidx, bss);
        CU(cudaGetLastError());
        std::vector<int32_t> hss(p->nbins);
        CU(cudaMemcpyAsync(hss.data(), bss, sizeof(int32_t) * p->nbins, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        p->n_slots = (int64_t)last_idx + last_flag;
        if (p->opt_fwd_pair == 1 && (double)p->n_slots > 0.9 * (double)M) {
            p->n_slots = 0;          // too few pairs to pay for the doubled FMAs
            p->launches += 4;
            p->points_set = true;
            return B2N_OK;
        }
        if ((rc = dev_alloc(p, (void**)&p->d_slots, sizeof(uint32_t) * (p->n_slots + 1)))) return rc;
        slot_write_kernel<<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(M, keys_s, isslot, slotidx,
                                                                         p->d_slots);
        CU(cudaGetLastError());
        if (p->opt_fwd_interleave && g.tile[0] <= 256 && p->nbins < ((int64_t)1 << 23)) {
            // reorder the slots of every bin by (rank within column, column): two stable sorts
            const int64_t ns = p->n_slots;
            uint64_t *k1 = nullptr, *k1s = nullptr;
            uint32_t* sl2 = nullptr;
            CU(scratch.alloc(&k1, sizeof(uint64_t) * ns));
            CU(scratch.alloc(&k1s, sizeof(uint64_t) * ns));
            CU(scratch.alloc(&sl2, sizeof(uint32_t) * ns));
            slot_colkey_kernel<<<grid_for(ns, 256, p->sm_count), 256, 0, st>>>(
                ns, g.cells_per_tile, g.tile[0], keys_s, p->d_slots, k1);
            CU(cudaGetLastError());
            int b1 = 1;
            while (b1 < 64 && (((uint64_t)p->nbins * (uint64_t)g.tile[0]) >> b1) != 0) b1++;
            size_t tb = 0;
            CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, k1, k1s, p->d_slots, sl2, ns, 0, 64, st));
            void* t2 = nullptr;
            CU(scratch.alloc(&t2, tb));
            CU(cub::DeviceRadixSort::SortPairs(t2, tb, k1, k1s, p->d_slots, sl2, ns, 0, b1, st));
            // rank inside the (bin, column) group (head / isslot buffers are free again)
            slot_grouphead_kernel<<<grid_for(ns, 256, p->sm_count), 256, 0, st>>>(ns, k1s, head);
            CU(cudaGetLastError());
            CU(cub::DeviceScan::InclusiveScan(stmp, sb1, head, head, cub::Max(), (int)ns, st));
            slot_rankkey_kernel<<<grid_for(ns, 256, p->sm_count), 256, 0, st>>>(ns, g.tile[0], k1s, head, k1);
            CU(cudaGetLastError());
            int b2 = 40;
            while (b2 < 64 && (((uint64_t)p->nbins << 40) >> b2) != 0) b2++;
            CU(cub::DeviceRadixSort::SortPairs(t2, tb, k1, k1s, sl2, p->d_slots, ns, 0, b2, st));
            p->launches += 6;
        }
        if ((rc = dev_alloc(p, (void**)&p->d_slot_kw, sizeof(int32_t) * p->n_slots * g.ndim))) return rc;
        if ((rc = dev_alloc(p, (void**)&p->d_slot_perm, sizeof(int32_t) * p->n_slots * 2))) return rc;
        slot_pack_kernel<<<grid_for(p->n_slots, 256, p->sm_count), 256, 0, st>>>(
            p->n_slots, g.ndim, M, p->d_slots, p->d_pt_kw, p->d_perm, p->d_slot_kw, p->d_slot_perm);
        CU(cudaGetLastError());
        std::vector<int4> fitems;
        int64_t nexts = p->n_slots;
        std::vector<int64_t> send(p->nbins, 0);
        for (int64_t b = p->nbins - 1; b >= 0; b--) {
            if (hss[b] >= 0) {
                send[b] = nexts;
                nexts = hss[b];
            }
        }
        for (int64_t b = 0; b < p->nbins; b++) {
            if (hss[b] < 0) continue;
            int64_t s0 = hss[b], e0 = send[b];
            while (s0 < e0) {
                int64_t c = e0 - s0 < p->opt_chunk ? e0 - s0 : p->opt_chunk;
                fitems.push_back(make_int4((int)b, (int)s0, (int)c, 0));
                s0 += c;
            }
        }
        p->n_items_f = (int64_t)fitems.size();
        if ((rc = dev_alloc(p, (void**)&p->d_items_f, sizeof(int4) * (fitems.size() + 1)))) return rc;
        CU(cudaMemcpy(p->d_items_f, fitems.data(), sizeof(int4) * fitems.size(), cudaMemcpyHostToDevice));
        CU(cudaStreamSynchronize(st));
        p->launches += 5;
    }
    p->launches += 4;
    p->points_set = true;
    return B2N_OK;
}

extern "C" int b2n_plan_set_points(b2n_plan* p, const void* coords_dev, int64_t M, int kind,
                                   void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (M < 0 || M >= ((int64_t)1 << 31)) return fail(B2N_EINVAL, "need 0 <= M < 2^31");
    if (M > 0 && coords_dev == nullptr) return fail(B2N_EINVAL, "coords is NULL");
    if (kind != B2N_COORD_TM && kind != B2N_COORD_OMEGA) return fail(B2N_EINVAL, "bad coordinate kind");
    ON_DEVICE(p);
    free_points(p);
    p->sparse_set = false;
    int rc = p->precision == B2N_SINGLE ? set_points_t<float>(p, coords_dev, M, kind, (cudaStream_t)stream)
                                        : set_points_t<double>(p, coords_dev, M, kind, (cudaStream_t)stream);
    if (rc) return rc;
    // plan-time weights are built here (or in set_tables, whichever comes last), not lazily by
    // the first transform: later calls on other streams then only ever READ them
    if ((rc = ensure_weights(p, (cudaStream_t)stream))) return rc;
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return B2N_OK;
}

extern "C" int b2n_plan_set_sample_phase(b2n_plan* p, const void* phase_dev, void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (!p->points_set) return fail(B2N_ESTATE, "set_points must precede set_sample_phase");
    ON_DEVICE(p);
    cudaStream_t st = (cudaStream_t)stream;
    if (phase_dev == nullptr) {
        dev_free(p, p->d_phase_s);
        dev_free(p, p->d_phase_sb);
        dev_free(p, p->d_phase_f);
        p->d_phase_s = p->d_phase_sb = p->d_phase_f = nullptr;
        return B2N_OK;
    }
    const int64_t M = p->g.M;
    if (p->d_phase_s == nullptr) {
        int rc = dev_alloc(p, &p->d_phase_s, p->cplx_size() * M);
        if (rc) return rc;
    }
    if (M > 0) {
        if (p->precision == B2N_SINGLE)
            gather_c_kernel<float2><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
                M, p->d_perm, (const float2*)phase_dev, (float2*)p->d_phase_s);
        else
            gather_c_kernel<double2><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
                M, p->d_perm, (const double2*)phase_dev, (double2*)p->d_phase_s);
        CU(cudaGetLastError());
        p->launches++;
        if (p->d_slots != nullptr && p->n_slots > 0) {
            if (p->d_phase_f == nullptr) {
                int rc = dev_alloc(p, &p->d_phase_f, 2 * p->cplx_size() * p->n_slots);
                if (rc) return rc;
            }
            if (p->precision == B2N_SINGLE)
                slot_phase_kernel<float2><<<grid_for(p->n_slots, 256, p->sm_count), 256, 0, st>>>(
                    p->n_slots, p->d_slots, p->d_perm, (const float2*)phase_dev, (float2*)p->d_phase_f);
            else
                slot_phase_kernel<double2><<<grid_for(p->n_slots, 256, p->sm_count), 256, 0, st>>>(
                    p->n_slots, p->d_slots, p->d_perm, (const double2*)phase_dev, (double2*)p->d_phase_f);
            CU(cudaGetLastError());
            p->launches++;
        }
        if (p->have_b) {
            if (p->d_phase_sb == nullptr) {
                int rc = dev_alloc(p, &p->d_phase_sb, p->cplx_size() * M);
                if (rc) return rc;
            }
            if (p->precision == B2N_SINGLE)
                gather_c_kernel<float2><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
                    M, p->d_perm_b, (const float2*)phase_dev, (float2*)p->d_phase_sb);
            else
                gather_c_kernel<double2><<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(
                    M, p->d_perm_b, (const double2*)phase_dev, (double2*)p->d_phase_sb);
            CU(cudaGetLastError());
        }
    }
    return B2N_OK;
}

extern "C" int64_t b2n_plan_num_points(b2n_plan* p) { return p && p->points_set ? p->g.M : -1; }
extern "C" int64_t b2n_plan_num_bins(b2n_plan* p) { return p ? (int64_t)p->g.nbin[0] * p->g.nbin[1] * p->g.nbin[2] : -1; }
extern "C" int64_t b2n_plan_num_slots(b2n_plan* p) { return p && p->points_set ? p->n_slots : -1; }
extern "C" int b2n_plan_get_slots(b2n_plan* p, uint32_t* slots_dev, void* stream) {
    if (p == nullptr || slots_dev == nullptr) return fail(B2N_EINVAL, "NULL argument");
    if (!p->points_set) return fail(B2N_ESTATE, "points not set");
    ON_DEVICE(p);
    if (p->n_slots > 0)
        CU(cudaMemcpyAsync(slots_dev, p->d_slots, sizeof(uint32_t) * p->n_slots,
                           cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B2N_OK;
}
extern "C" int64_t b2n_plan_device_bytes(b2n_plan* p) { return p ? p->dev_bytes : -1; }
extern "C" int64_t b2n_plan_launch_count(b2n_plan* p) { return p ? p->launches : -1; }

// out[0..3] = {forward interpolation kernel: total ms, launches; adjoint: total ms, launches}
// measured with CUDA events on the launching stream since the last call (option "profile")
extern "C" int b2n_plan_get_timing(b2n_plan* p, double* out) {
    if (p == nullptr || out == nullptr) return fail(B2N_EINVAL, "NULL argument");
    ON_DEVICE(p);
    for (int k = 0; k < 2; k++) {
        auto& v = k == 0 ? p->ev_fwd : p->ev_adj;
        double tot = 0;
        for (auto& e : v) {
            float ms = 0;
            CU(cudaEventSynchronize(e.second));
            CU(cudaEventElapsedTime(&ms, e.first, e.second));
            tot += ms;
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        out[2 * k] = tot;
        out[2 * k + 1] = (double)v.size();
        v.clear();
    }
    return B2N_OK;
}

extern "C" int b2n_plan_get_points(b2n_plan* p, void* tm_dev, int32_t* bin_ids_dev,
                                   int64_t* keys_dev, int32_t* perm_dev, void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (!p->points_set) return fail(B2N_ESTATE, "points not set");
    ON_DEVICE(p);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = p->g.M;
    if (M == 0) return B2N_OK;
    if (tm_dev) CU(cudaMemcpyAsync(tm_dev, p->d_tm, p->real_size() * M * p->g.ndim, cudaMemcpyDeviceToDevice, st));
    if (bin_ids_dev) CU(cudaMemcpyAsync(bin_ids_dev, p->d_bin_ids, sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, st));
    if (perm_dev) CU(cudaMemcpyAsync(perm_dev, p->d_perm, sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, st));
    if (keys_dev) {
        keys_to_i64_kernel<<<grid_for(M, 256, p->sm_count), 256, 0, st>>>(M, p->d_keys, keys_dev);
        CU(cudaGetLastError());
    }
    return B2N_OK;
}

// ---------------------------------------------------------------------------------
// interpolation dispatch
// ---------------------------------------------------------------------------------
static TablePtrs table_ptrs(const b2n_plan* p) {
    return TablePtrs{{p->d_tab[0], p->d_tab[1], p->d_tab[2]}};
}

static int run_generic(b2n_plan* p, bool fwd, const void* in, void* out, int nbatch, bool phase,
                       cudaStream_t st) {
    const TablePtrs tabs = table_ptrs(p);
    const void* ph = phase ? p->d_phase_s : nullptr;
    if (p->precision == B2N_SINGLE) {
        // float adjoint: per-cell sums in a double scratch grid (kept by the plan, at most 1 GiB;
        // above that the float atomics go straight to the output grid)
        void* acc64 = nullptr;
        const size_t need = sizeof(double2) * (size_t)p->g.PK * nbatch;
        if (!fwd && need <= ((size_t)1 << 30)) {
            if (p->acc64_bytes < need) {
                dev_free(p, p->d_acc64);
                p->d_acc64 = nullptr;
                p->acc64_bytes = 0;
                if (dev_alloc(p, &p->d_acc64, need) == B2N_OK) p->acc64_bytes = need;
            }
            if (p->d_acc64 != nullptr) {
                if (cudaMemsetAsync(p->d_acc64, 0, need, st) != cudaSuccess) return (int)cudaGetLastError();
                acc64 = p->d_acc64;
                p->lib_calls++;
                p->launches++;       // the add-back kernel
            }
        }
        return generic_launch_f32(p->g, p->cplx_table, tabs, p->d_tm_s, p->d_perm, fwd, in, out, ph,
                                  nbatch, p->sm_count, acc64, st);
    }
    return generic_launch_f64(p->g, p->cplx_table, tabs, p->d_tm_s, p->d_perm, fwd, in, out, ph,
                              nbatch, p->sm_count, nullptr, st);
}

// weights need both the points and the tables; built on first use
template <typename T>
static int build_weights_t(b2n_plan* p, cudaStream_t st) {
    const Geom& g = p->g;
    TabArgs tabs{{p->d_tab[0], p->d_tab[1], p->d_tab[2]}};
    int rc;
    // forward arrays at the forward kernels' width, the adjoint order at the adjoint kernels';
    // without an adjoint order the adjoint reads the forward arrays (at the forward width)
    const int jf = p->jk_fwd, ja = p->jk_adj;
    const bool packed = p->opt_fwd_pair && p->d_slots != nullptr && jf > 0;
    if (packed) {
        // forward weights in slot order: (weight, partner's weight) pairs
        if ((rc = dev_alloc(p, &p->d_wts_f, 2 * sizeof(T) * (size_t)g.ndim * jf * p->n_slots))) return rc;
        slot_weights_kernel<T><<<grid_for(p->n_slots, 256, p->sm_count), 256, 0, st>>>(
            g, tabs, jf, p->n_slots, p->d_slots, (const T*)p->d_tm_s, p->d_pt_ko,
            (typename Cplx<T>::type*)p->d_wts_f);
        CU(cudaGetLastError());
        p->launches++;
    }
    // the sample-ordered weights of sort order A are read by the unpaired forward and by the
    // adjoint when it has no order of its own
    const size_t wsz = p->cplx_table ? 2 * sizeof(T) : sizeof(T);      // complex tables: complex weights
    using CW = typename Cplx<T>::type;
    if (jf > 0 && (!packed || !p->have_b)) {
        if ((rc = dev_alloc(p, &p->d_wts, wsz * (size_t)g.ndim * jf * g.M))) return rc;
        if (p->cplx_table)
            point_weights_kernel<T, true><<<grid_for(g.M, 256, p->sm_count), 256, 0, st>>>(
                g, tabs, jf, (const T*)p->d_tm_s, p->d_pt_ko, (CW*)p->d_wts);
        else
            point_weights_kernel<T><<<grid_for(g.M, 256, p->sm_count), 256, 0, st>>>(
                g, tabs, jf, (const T*)p->d_tm_s, p->d_pt_ko, (T*)p->d_wts);
        CU(cudaGetLastError());
        p->launches++;
    }
    if (p->have_b && ja > 0 && g.colmode) {
        // column order: blocked records (weights, origins, acquisition index) of the column-group kernel
        Scratch scratch;
        T* tm_sb = nullptr;
        int32_t *ko_b = nullptr, *kw_b = nullptr;
        CU(scratch.alloc(&tm_sb, sizeof(T) * g.M * 3));
        CU(scratch.alloc(&ko_b, sizeof(int32_t) * g.M * 3));
        CU(scratch.alloc(&kw_b, sizeof(int32_t) * g.M * 3));
        const int nb = grid_for(g.M, 256, p->sm_count);
        gather_points_kernel<T><<<nb, 256, 0, st>>>(3, g.M, p->d_perm_b, (const T*)p->d_tm, tm_sb);
        CU(cudaGetLastError());
        point_windows_kernel<T><<<nb, 256, 0, st>>>(g, tm_sb, ko_b, kw_b);
        CU(cudaGetLastError());
        const bool f32 = sizeof(T) == 4;
        const size_t bytes = f32 ? column_record_bytes_f32(ja, g.M) : column_record_bytes_f64(ja, g.M);
        if ((rc = dev_alloc(p, &p->d_col_rec, bytes))) return rc;
        CU(cudaMemsetAsync(p->d_col_rec, 0, bytes, st));
        const TablePtrs tp{{p->d_tab[0], p->d_tab[1], p->d_tab[2]}};
        const int e = f32 ? column_build_f32(g, ja, tp, tm_sb, ko_b, kw_b, p->d_perm_b, (int)p->opt_win_maxslide, p->d_col_rec, nb, st)
                          : column_build_f64(g, ja, tp, tm_sb, ko_b, kw_b, p->d_perm_b, (int)p->opt_win_maxslide, p->d_col_rec, nb, st);
        if (e != 0) return fail(B2N_ECUDA, std::string("column records: ") + cudaGetErrorString((cudaError_t)e));
        CU(cudaStreamSynchronize(st));      // the scratch arrays are freed on return
        p->launches += 3;
    } else if (p->have_b && ja > 0) {
        if ((rc = dev_alloc(p, &p->d_wts_b, wsz * (size_t)g.ndim * ja * g.M))) return rc;
        if (p->cplx_table)
            point_weights_kernel<T, true><<<grid_for(g.M, 256, p->sm_count), 256, 0, st>>>(
                g, tabs, ja, (const T*)p->d_tm_sb, p->d_pt_ko_b, (CW*)p->d_wts_b);
        else
            point_weights_kernel<T><<<grid_for(g.M, 256, p->sm_count), 256, 0, st>>>(
                g, tabs, ja, (const T*)p->d_tm_sb, p->d_pt_ko_b, (T*)p->d_wts_b);
        CU(cudaGetLastError());
        p->launches++;
    }
    return B2N_OK;
}

static int ensure_weights(b2n_plan* p, cudaStream_t st) {
    if (!p->opt_precomp || p->g.ndim < 2 || p->d_wts != nullptr ||
        p->d_wts_f != nullptr || p->d_wts_b != nullptr || p->d_col_rec != nullptr || p->g.M == 0 ||
        (p->jk_fwd == 0 && p->jk_adj == 0))
        return B2N_OK;
    if (!p->tables_set || !p->points_set) return B2N_OK;
    return p->precision == B2N_SINGLE ? build_weights_t<float>(p, st) : build_weights_t<double>(p, st);
}

static void prof_begin(b2n_plan* p, bool fwd, cudaStream_t st) {
    if (!p->opt_profile) return;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    (fwd ? p->ev_fwd : p->ev_adj).push_back({e0, e1});
}
static void prof_end(b2n_plan* p, bool fwd, cudaStream_t st) {
    if (!p->opt_profile) return;
    cudaEventRecord((fwd ? p->ev_fwd : p->ev_adj).back().second, st);
}

static int check_ready(b2n_plan* p, const void* a, const void* b, int nbatch) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (!p->points_set) return fail(B2N_ESTATE, "points not set");
    if (nbatch < 1) return fail(B2N_EINVAL, "nbatch must be >= 1");
    if ((a == nullptr || b == nullptr) && p->g.M > 0) return fail(B2N_EINVAL, "NULL array");
    return B2N_OK;
}

static int interp_fwd_impl(b2n_plan* p, const void* grid, void* samples, int nbatch, bool phase,
                           cudaStream_t st) {
    if (!p->tables_set) return fail(B2N_ESTATE, "tables not set");
    if (p->g.M == 0) return B2N_OK;
    {
        int rc = ensure_weights(p, st);
        if (rc) return rc;
    }
    bool done = false;
    prof_begin(p, true, st);
    if (!p->opt_force_generic && p->jk_fwd > 0 && (!p->cplx_table || p->d_wts != nullptr)) {
        const void* ph = phase ? p->d_phase_s : nullptr;
        FwdOpts fo;
        fo.use_tma = p->opt_use_tma ? 1 : 0;
        fo.pitch = (int)p->opt_fwd_pitch;
        const bool pairs = p->opt_fwd_pair && p->d_slots != nullptr;
        const int4* fit = pairs ? p->d_items_f : p->d_items;
        const int64_t nfit = pairs ? p->n_items_f : p->n_items;
        SlotArgs sa;
        if (pairs) {
            sa.slots = p->d_slots;
            sa.ns = p->n_slots;
            sa.packed = p->d_wts_f != nullptr ? 1 : 0;
            sa.wts2 = p->d_wts_f;
            sa.kw = p->d_slot_kw;
            sa.perm = p->d_slot_perm;
            sa.phase2 = phase ? p->d_phase_f : nullptr;
        }
        int rc = p->precision == B2N_SINGLE
                     ? tiled_fwd_f32(p->g, p->jk_fwd, p->cplx_table != 0, p->tables_equal, table_ptrs(p), p->d_tm_s, p->d_wts, p->d_pt_ko, p->d_pt_kw, p->d_perm, fit,
                                     nfit, sa, grid, samples, ph, nbatch, fo, st, &done)
                     : tiled_fwd_f64(p->g, p->jk_fwd, p->cplx_table != 0, p->tables_equal, table_ptrs(p), p->d_tm_s, p->d_wts, p->d_pt_ko, p->d_pt_kw, p->d_perm, fit,
                                     nfit, sa, grid, samples, ph, nbatch, fo, st, &done);
        if (rc != 0) return fail(B2N_ECUDA, "tiled forward launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
    }
    if (!done) {
        int rc = run_generic(p, true, grid, samples, nbatch, phase, st);
        if (rc != 0) return fail(B2N_ECUDA, "generic forward launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
    }
    prof_end(p, true, st);
    p->last_fwd_kernel = done ? 1 : 0;
    p->launches++;
    return B2N_OK;
}

static int interp_adj_impl(b2n_plan* p, const void* samples, void* grid, int nbatch, bool phase,
                           cudaStream_t st, bool accumulate = false) {
    if (!p->tables_set) return fail(B2N_ESTATE, "tables not set");
    // the profiled interval covers the zero-fill too: SURVEY 8(d) counts it in the adjoint's bytes
    const bool prof = p->g.M > 0;
    if (prof) prof_begin(p, false, st);
    if (!accumulate) {
        CU(cudaMemsetAsync(grid, 0, p->cplx_size() * p->g.PK * nbatch, st));
        p->lib_calls++;
    }
    if (p->g.M == 0) return B2N_OK;
    {
        int rc = ensure_weights(p, st);
        if (rc) return rc;
    }
    bool done = false;
    if (!p->opt_force_generic && p->g.ndim == 2 && p->have_b && p->jk_adj > 0 &&
        (!p->cplx_table || p->d_wts_b != nullptr)) {
        // 2-D: register windows sliding along axis 2 (adjoint sort order), lanes <-> (j1, coil)
        const void* ph = phase ? p->d_phase_sb : nullptr;
        WindowOpts wo;
        wo.pts_per_warp = p->opt_slide_pts ? (int)p->opt_slide_pts : 256;
        wo.max_slide = (int)p->opt_win_maxslide;
        int rc = p->precision == B2N_SINGLE
                     ? window2d_adj_f32(p->g, p->jk_adj, p->cplx_table != 0, table_ptrs(p), p->d_tm_sb, p->d_wts_b, p->d_pt_ko_b, p->d_pt_kw_b,
                                        p->d_perm_b, samples, grid, ph, nbatch, wo, st, &done)
                     : window2d_adj_f64(p->g, p->jk_adj, p->cplx_table != 0, table_ptrs(p), p->d_tm_sb, p->d_wts_b, p->d_pt_ko_b, p->d_pt_kw_b,
                                        p->d_perm_b, samples, grid, ph, nbatch, wo, st, &done);
        if (rc != 0) return fail(B2N_ECUDA, "2-D window adjoint launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
        if (done) p->last_adj_kernel = 4;
    }
    if (!done && !p->opt_force_generic && p->g.colmode && p->d_col_rec != nullptr) {
        // 3-D, real table: column-group register window (one warp per group of grid columns)
        WindowOpts wo;
        wo.pts_per_warp = p->opt_slide_pts ? (int)p->opt_slide_pts : 512;
        wo.max_slide = (int)p->opt_win_maxslide;
        const void* ph = phase ? p->d_phase_sb : nullptr;
        int rc = p->precision == B2N_SINGLE
                     ? column_adj_f32(p->g, p->jk_adj, wo, p->d_col_rec, samples, grid, ph, nbatch, st, &done)
                     : column_adj_f64(p->g, p->jk_adj, wo, p->d_col_rec, samples, grid, ph, nbatch, st, &done);
        if (rc != 0) return fail(B2N_ECUDA, "column adjoint launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
        if (done) p->last_adj_kernel = 5;
    }
    if (!done && !p->opt_force_generic && p->g.ndim == 3 && !p->g.colmode && (p->have_b ? p->jk_adj : p->jk_fwd) > 0 &&
        (!p->cplx_table || (p->have_b ? p->d_wts_b : p->d_wts) != nullptr)) {
        // register window, lane-parallel batch weights; adjoint sort order when built
        const bool ob = p->have_b;
        const int jk = ob ? p->jk_adj : p->jk_fwd;
        const void* ph = phase ? (ob ? p->d_phase_sb : p->d_phase_s) : nullptr;
        const void* tms = ob ? p->d_tm_sb : p->d_tm_s;
        const void* wts = ob ? p->d_wts_b : p->d_wts;
        const int32_t* ko = ob ? p->d_pt_ko_b : p->d_pt_ko;
        const int32_t* kw = ob ? p->d_pt_kw_b : p->d_pt_kw;
        const int32_t* pm = ob ? p->d_perm_b : p->d_perm;
        WindowOpts wo;
        wo.slide_axis = ob ? 2 : 0;
        wo.pts_per_warp = p->opt_slide_pts ? (int)p->opt_slide_pts : 256;
        wo.max_slide = (int)p->opt_win_maxslide;
        wo.facew = (int)p->opt_win_facew;
        if (wo.facew < 0) wo.facew = jk <= 6 ? (p->precision == B2N_SINGLE ? 2 : 1) : 0;
        int rc = p->precision == B2N_SINGLE
                     ? window_adj_f32(p->g, jk, p->cplx_table != 0, table_ptrs(p), wo, tms, wts, ko, kw, pm, samples, grid, ph, nbatch, st, &done)
                     : window_adj_f64(p->g, jk, p->cplx_table != 0, table_ptrs(p), wo, tms, wts, ko, kw, pm, samples, grid, ph, nbatch, st, &done);
        if (rc != 0) return fail(B2N_ECUDA, "window adjoint launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
        if (done) p->last_adj_kernel = 3;
    }
    if (!done) {
        int rc = run_generic(p, false, samples, grid, nbatch, phase, st);
        if (rc != 0) return fail(B2N_ECUDA, "generic adjoint launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
    }
    prof_end(p, false, st);
    if (!done) p->last_adj_kernel = 0;
    p->launches++;
    return B2N_OK;
}

extern "C" int b2n_interp_fwd(b2n_plan* p, const void* grid_dev, void* samples_dev, int nbatch,
                              int apply_phase, void* stream) {
    int rc = check_ready(p, grid_dev, samples_dev, nbatch);
    if (rc) return rc;
    ON_DEVICE(p);
    return interp_fwd_impl(p, grid_dev, samples_dev, nbatch, apply_phase && p->d_phase_s, (cudaStream_t)stream);
}

extern "C" int b2n_interp_adj(b2n_plan* p, const void* samples_dev, void* grid_dev, int nbatch,
                              int apply_phase, void* stream) {
    int rc = check_ready(p, samples_dev, grid_dev, nbatch);
    if (rc) return rc;
    if (grid_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    ON_DEVICE(p);
    return interp_adj_impl(p, samples_dev, grid_dev, nbatch, (apply_phase & 1) && p->d_phase_s,
                           (cudaStream_t)stream, (apply_phase & 2) != 0);
}

// ---------------------------------------------------------------------------------
// sparse mode
// ---------------------------------------------------------------------------------
extern "C" int b2n_plan_set_sparse(b2n_plan* p, const void* const* coef,
                                   const int32_t* const* kidx, int64_t M,
                                   const void* row_phase, void* stream) {
    if (p == nullptr || coef == nullptr || kidx == nullptr) return fail(B2N_EINVAL, "NULL argument");
    if (!p->points_set) return fail(B2N_ESTATE, "set_points must precede set_sparse");
    if (M != p->g.M) return fail(B2N_EINVAL, "M does not match set_points");
    if (p->g.Kg[1] != p->g.K[1]) return fail(B2N_EINVAL, "sparse mode is not available on slab plans");
    ON_DEVICE(p);
    cudaStream_t st = (cudaStream_t)stream;
    const Geom& g = p->g;
    int nnzr = 1;
    for (int d = 0; d < g.ndim; d++) nnzr *= g.J[d];
    const int64_t nnz = M * nnzr;
    const size_t vsz = p->cplx_table ? p->cplx_size() : p->real_size();
    dev_free(p, p->d_ell_vals); dev_free(p, p->d_ell_cols);
    p->d_ell_vals = nullptr; p->d_ell_cols = nullptr;
    int rc;
    if ((rc = dev_alloc(p, &p->d_ell_vals, vsz * nnz))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_ell_cols, sizeof(int32_t) * nnz))) return rc;
    p->nnzr = nnzr;
    if (nnz > 0) {
        SparseSrc src{};
        for (int d = 0; d < g.ndim; d++) { src.coef[d] = coef[d]; src.kidx[d] = kidx[d]; }
        const int grid = grid_for(nnz, 256, p->sm_count);
        const double2* rp = (const double2*)row_phase;
        if (p->precision == B2N_SINGLE) {
            if (p->cplx_table) build_ell_kernel<float, true><<<grid, 256, 0, st>>>(g, src, p->d_perm, rp, nnzr, (float2*)p->d_ell_vals, p->d_ell_cols);
            else build_ell_kernel<float, false><<<grid, 256, 0, st>>>(g, src, p->d_perm, rp, nnzr, (float*)p->d_ell_vals, p->d_ell_cols);
        } else {
            if (p->cplx_table) build_ell_kernel<double, true><<<grid, 256, 0, st>>>(g, src, p->d_perm, rp, nnzr, (double2*)p->d_ell_vals, p->d_ell_cols);
            else build_ell_kernel<double, false><<<grid, 256, 0, st>>>(g, src, p->d_perm, rp, nnzr, (double*)p->d_ell_vals, p->d_ell_cols);
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));
        p->launches++;
    }
    p->sparse_set = true;
    return B2N_OK;
}

extern "C" int64_t b2n_plan_sparse_nnz(b2n_plan* p) {
    return p && p->sparse_set ? p->g.M * p->nnzr : -1;
}

extern "C" int b2n_plan_get_sparse(b2n_plan* p, void* vals_dev, int32_t* cols_dev, void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (!p->sparse_set) return fail(B2N_ESTATE, "sparse matrix not set");
    ON_DEVICE(p);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nnz = p->g.M * p->nnzr;
    if (nnz == 0) return B2N_OK;
    const int grid = grid_for(nnz, 256, p->sm_count);
    if (p->precision == B2N_SINGLE) {
        if (p->cplx_table) ell_unpermute_kernel<float2><<<grid, 256, 0, st>>>(p->g.M, p->nnzr, p->d_perm, (const float2*)p->d_ell_vals, p->d_ell_cols, (float2*)vals_dev, cols_dev);
        else ell_unpermute_kernel<float><<<grid, 256, 0, st>>>(p->g.M, p->nnzr, p->d_perm, (const float*)p->d_ell_vals, p->d_ell_cols, (float*)vals_dev, cols_dev);
    } else {
        if (p->cplx_table) ell_unpermute_kernel<double2><<<grid, 256, 0, st>>>(p->g.M, p->nnzr, p->d_perm, (const double2*)p->d_ell_vals, p->d_ell_cols, (double2*)vals_dev, cols_dev);
        else ell_unpermute_kernel<double><<<grid, 256, 0, st>>>(p->g.M, p->nnzr, p->d_perm, (const double*)p->d_ell_vals, p->d_ell_cols, (double*)vals_dev, cols_dev);
    }
    CU(cudaGetLastError());
    return B2N_OK;
}

template <typename T, bool CT>
static void launch_spmv(b2n_plan* p, bool fwd, const void* in, void* out, int nbatch, bool phase,
                        cudaStream_t st) {
    using C = cplx_t<T>;
    using W = typename WeightT<T, CT>::type;
    const C* ph = phase ? (const C*)p->d_phase_s : nullptr;
    const int grid = grid_for(p->g.M * 32, 256, p->sm_count, 16);
    if (fwd)
        spmv_fwd_kernel<T, CT><<<grid, 256, 0, st>>>(p->g, p->nnzr, (const W*)p->d_ell_vals, p->d_ell_cols, p->d_perm, (const C*)in, (C*)out, ph, nbatch);
    else
        spmv_adj_kernel<T, CT><<<grid, 256, 0, st>>>(p->g, p->nnzr, (const W*)p->d_ell_vals, p->d_ell_cols, p->d_perm, (const C*)in, (C*)out, ph, nbatch);
}

static int spmv_impl(b2n_plan* p, bool fwd, const void* in, void* out, int nbatch, bool phase,
                     cudaStream_t st) {
    if (!p->sparse_set) return fail(B2N_ESTATE, "sparse matrix not set");
    if (!fwd) {
        CU(cudaMemsetAsync(out, 0, p->cplx_size() * p->g.PK * nbatch, st));
        p->lib_calls++;
    }
    if (p->g.M == 0) return B2N_OK;
    if (p->precision == B2N_SINGLE) {
        if (p->cplx_table) launch_spmv<float, true>(p, fwd, in, out, nbatch, phase, st);
        else launch_spmv<float, false>(p, fwd, in, out, nbatch, phase, st);
    } else {
        if (p->cplx_table) launch_spmv<double, true>(p, fwd, in, out, nbatch, phase, st);
        else launch_spmv<double, false>(p, fwd, in, out, nbatch, phase, st);
    }
    CU(cudaGetLastError());
    p->launches++;
    return B2N_OK;
}

extern "C" int b2n_spmv_fwd(b2n_plan* p, const void* grid_dev, void* samples_dev, int nbatch,
                            int apply_phase, void* stream) {
    int rc = check_ready(p, grid_dev, samples_dev, nbatch);
    if (rc) return rc;
    ON_DEVICE(p);
    return spmv_impl(p, true, grid_dev, samples_dev, nbatch, apply_phase && p->d_phase_s,
                     (cudaStream_t)stream);
}

extern "C" int b2n_spmv_adj(b2n_plan* p, const void* samples_dev, void* grid_dev, int nbatch,
                            int apply_phase, void* stream) {
    int rc = check_ready(p, samples_dev, grid_dev, nbatch);
    if (rc) return rc;
    if (grid_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    ON_DEVICE(p);
    return spmv_impl(p, false, samples_dev, grid_dev, nbatch, apply_phase && p->d_phase_s,
                     (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------
// full transforms: scale/pad -> cuFFT -> phase -> interpolate, and the mirror image
// ---------------------------------------------------------------------------------
static int get_fft(b2n_plan* p, int nbatch, cufftHandle* out) {
    auto it = p->fft_plans.find(nbatch);
    if (it != p->fft_plans.end()) {
        *out = it->second;
        return B2N_OK;
    }
    const Geom& g = p->g;
    long long n[3];
    for (int d = 0; d < g.ndim; d++) n[d] = g.K[g.ndim - 1 - d];   // slowest axis first
    cufftHandle h;
    FFT(cufftCreate(&h));
    size_t ws = 0;
    // 64-bit plan: prod(Kd) * nbatch may exceed 2^31 elements (384^3 with 38 coils)
    FFT(cufftMakePlanMany64(h, g.ndim, n, nullptr, 1, (long long)g.PK, nullptr, 1, (long long)g.PK,
                            p->precision == B2N_SINGLE ? CUFFT_C2C : CUFFT_Z2Z, (long long)nbatch, &ws));
    p->dev_bytes += (int64_t)ws;
    p->fft_plans[nbatch] = h;
    *out = h;
    return B2N_OK;
}

// Pruned oversampled FFT (3-D, one volume): the zero-padded input is non-zero only in the
// planes k3 < N3 and the adjoint output is cropped to them, so the two in-plane passes run
// on N3 of the K3 planes only (a batched 2-D plan over contiguous planes) and the pass
// along axis 3 is a strided batched 1-D plan.  Same transform, fewer bytes moved.
static bool pruned_ok(const b2n_plan* p, int nbatch) {
    return p->opt_pruned_fft && p->g.ndim == 3 && nbatch == 1 && p->g.N[2] < p->g.K[2];
}

static int get_pruned_fft(b2n_plan* p, cufftHandle* plan2d, cufftHandle* plan1d) {
    if (p->fft_pruned_ready) {
        *plan2d = p->fft_2d;
        *plan1d = p->fft_1d;
        return B2N_OK;
    }
    const Geom& g = p->g;
    const cufftType type = p->precision == B2N_SINGLE ? CUFFT_C2C : CUFFT_Z2Z;
    size_t ws = 0;
    int n2[2] = {g.K[1], g.K[0]};
    FFT(cufftCreate(&p->fft_2d));
    FFT(cufftMakePlanMany(p->fft_2d, 2, n2, nullptr, 1, g.K[0] * g.K[1], nullptr, 1, g.K[0] * g.K[1],
                          type, g.N[2], &ws));
    p->dev_bytes += (int64_t)ws;
    int n1[1] = {g.K[2]};
    int embed[1] = {g.K[2]};
    FFT(cufftCreate(&p->fft_1d));
    FFT(cufftMakePlanMany(p->fft_1d, 1, n1, embed, g.K[0] * g.K[1], 1, embed, g.K[0] * g.K[1], 1,
                          type, g.K[0] * g.K[1], &ws));
    p->dev_bytes += (int64_t)ws;
    p->fft_pruned_ready = true;
    *plan2d = p->fft_2d;
    *plan1d = p->fft_1d;
    return B2N_OK;
}

// own axis-3 pass (fft_axis3.cuh): radix schedule + twiddle table, prepared on first use
template <typename T>
static int prepare_axis3(b2n_plan* p) {
    if (p->ax3_state != 0) return B2N_OK;
    const int L = p->g.K[2];
    int max_smem = 0;
    CU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
    const bool fixed = fixed_npass(L) > 0 && fft_lines_smem<T, 0>(L) <= (size_t)max_smem && p->opt_own_fft3 != 2;
    const bool general = axis3_factor(L, &p->ax3) && Axis3Cfg<T>::smem(L) <= (size_t)max_smem && Axis3Cfg<T>::fits(L);
    if (!fixed && !general) {
        p->ax3_state = -1;
        return B2N_OK;
    }
    p->ax3_general = general;
    std::vector<T> tw(2 * (size_t)L);
    for (int t = 0; t < L; t++) {
        const double a = -2.0 * M_PI * (double)t / (double)L;
        tw[2 * t] = (T)std::cos(a);
        tw[2 * t + 1] = (T)std::sin(a);
    }
    int rc = dev_alloc(p, &p->d_tw3, sizeof(T) * 2 * (size_t)L);
    if (rc) return rc;
    CU(cudaMemcpy(p->d_tw3, tw.data(), sizeof(T) * 2 * (size_t)L, cudaMemcpyHostToDevice));
    p->ax3_state = 1;
    return B2N_OK;
}

// true when run_fft will apply phase_before / conj(phase_before) itself

template <typename T>
static int exec_fft(cufftHandle h, void* data, int dir) {
    if (sizeof(T) == 4) FFT(cufftExecC2C(h, (cufftComplex*)data, (cufftComplex*)data, dir));
    else FFT(cufftExecZ2Z(h, (cufftDoubleComplex*)data, (cufftDoubleComplex*)data, dir));
    return B2N_OK;
}

// the fused axis-3 pass: compile-time schedule when K3 has one (own_fft3 = 1), else / with
// own_fft3 = 2 the run-time radix schedule.  Returns 0 or a cudaError_t.
template <typename T>
static int run_axis3(b2n_plan* p, bool inverse, const void* a1, void* data, cudaStream_t st) {
    if (p->opt_own_fft3 != 2) {
        if (p->max_smem == 0) cudaDeviceGetAttribute(&p->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device);
        const Geom& g = p->g;
        LineArgs<T> la{};
        la.data = (cplx_t<T>*)data;
        la.tw = (const cplx_t<T>*)p->d_tw3;
        la.nz = g.N[2];
        la.ntiles = 1;                               // one outer block: the whole grid
        la.row_stride = (int64_t)g.K[0] * g.K[1];
        la.outer_stride = 0;
        la.inner_extent = la.row_stride;
        la.a1 = (const T*)a1;
        la.a2 = (const T*)p->d_pb[1];
        la.a3 = (const T*)p->d_pb[2];
        la.K1 = g.K[0];
        bool done = false;
        const int rc = fft_lines_launch<T, 0>(g.K[2], inverse, la, p->sm_count, p->max_smem, st, &done);
        if (rc != 0 || done) return rc;
    }
    if (!p->ax3_general) return (int)cudaErrorNotSupported;
    return fft_axis3_launch<T>(p->ax3, p->g, inverse, p->d_tw3, a1, p->d_pb[1], p->d_pb[2], data, p->sm_count, st);
}

// forward: in-plane passes on the non-zero planes, then axis 3; inverse: the reverse
template <typename T>
static bool axis3_fused(b2n_plan* p, int nbatch) {
    if (!p->opt_own_fft3 || !pruned_ok(p, nbatch)) return false;
    if (prepare_axis3<T>(p) != B2N_OK) return false;
    return p->ax3_state == 1;
}

// Own in-plane passes (option own_fft12, default on): axis 1 with the scale / zero-pad (forward)
// or crop / scale (adjoint) fused, axis 2 on the N2 non-zero rows only -- instead of a scale/pad
// sweep + cuFFT's two passes over the zero-padded planes.  Needs the fixed-schedule kernel for
// K1 and K2 and the fused axis-3 pass; one volume, no coil maps.
template <typename T>
static bool inplane_own(b2n_plan* p, int nbatch) {
    if (!p->opt_own_fft12 || p->opt_own_fft3 == 2 || !axis3_fused<T>(p, nbatch)) return false;
    if (p->inplane_state != 0) return p->inplane_state == 1;
    p->inplane_state = -1;
    const Geom& g = p->g;
    if (p->max_smem == 0) cudaDeviceGetAttribute(&p->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device);
    const size_t s1 = fft_lines_smem<T, 1>(g.K[0]), s2 = fft_lines_smem<T, 0>(g.K[1]);
    if (s1 == 0 || s2 == 0 || s1 > (size_t)p->max_smem || s2 > (size_t)p->max_smem) return false;
    if (fft_lines_smem<T, 0>(g.K[2]) == 0) return false;      // axis 3 on the run-time schedule: keep cuFFT in-plane
    for (int d = 0; d < 2; d++) {
        const int L = g.K[d];
        if (L == g.K[2]) { p->d_tw12[d] = p->d_tw3; continue; }
        if (d == 1 && L == g.K[0]) { p->d_tw12[1] = p->d_tw12[0]; continue; }
        std::vector<T> tw(2 * (size_t)L);
        for (int t = 0; t < L; t++) {
            const double a = -2.0 * M_PI * (double)t / (double)L;
            tw[2 * t] = (T)std::cos(a);
            tw[2 * t + 1] = (T)std::sin(a);
        }
        void* q = nullptr;
        if (dev_alloc(p, &q, sizeof(T) * 2 * (size_t)L) != B2N_OK) return false;
        if (cudaMemcpy(q, tw.data(), sizeof(T) * 2 * (size_t)L, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        p->d_tw12[d] = q;
        p->tw12_owned[d] = true;
    }
    p->inplane_state = 1;
    return true;
}

// the two in-plane passes; returns 0 or a cudaError_t
template <typename T>
static int run_inplane(b2n_plan* p, bool inverse, const void* image_in, void* image_out, void* grid,
                       cudaStream_t st, int z0 = 0, int nz = -1) {
    // (z0, nz): the image planes [z0, z0 + nz) only, image and grid pointers at plane z0 (slab
    // plane stage); default: the whole volume
    const Geom& g = p->g;
    if (nz < 0) nz = g.N[2];
    LineArgs<T> rows{};
    rows.data = (cplx_t<T>*)grid;
    rows.tw = (const cplx_t<T>*)p->d_tw12[0];
    rows.image = inverse ? (cplx_t<T>*)image_out : (cplx_t<T>*)const_cast<void*>(image_in);
    rows.NL2 = g.N[1];
    rows.K2 = g.K[1];
    rows.N1 = g.N[0];
    rows.nlines = (int64_t)g.N[1] * nz;
    rows.sn1 = p->d_sn[0];
    rows.sn2 = p->d_sn[1];
    rows.sn3 = p->d_sn[2] + z0;
    const double sc = inverse ? p->adj_scale : p->fwd_scale;
    rows.scale = (T)sc;
    rows.apply_scale = sc != 1.0;
    LineArgs<T> cols{};
    cols.data = (cplx_t<T>*)grid;
    cols.tw = (const cplx_t<T>*)p->d_tw12[1];
    cols.nz = g.N[1];
    cols.ntiles = nz;                                // outer blocks: the non-zero planes
    cols.row_stride = g.K[0];
    cols.outer_stride = (int64_t)g.K[0] * g.K[1];
    cols.inner_extent = g.K[0];
    bool done = false;
    int rc;
    if (!inverse) {
        rc = fft_lines_launch<T, 1>(g.K[0], false, rows, p->sm_count, p->max_smem, st, &done);
        if (rc != 0 || !done) return rc ? rc : (int)cudaErrorNotSupported;
        rc = fft_lines_launch<T, 0>(g.K[1], false, cols, p->sm_count, p->max_smem, st, &done);
        if (rc != 0 || !done) return rc ? rc : (int)cudaErrorNotSupported;
    } else {
        rc = fft_lines_launch<T, 0>(g.K[1], true, cols, p->sm_count, p->max_smem, st, &done);
        if (rc != 0 || !done) return rc ? rc : (int)cudaErrorNotSupported;
        rc = fft_lines_launch<T, 1>(g.K[0], true, rows, p->sm_count, p->max_smem, st, &done);
        if (rc != 0 || !done) return rc ? rc : (int)cudaErrorNotSupported;
    }
    p->launches += 2;
    return 0;
}

template <typename T>
static int run_fft(b2n_plan* p, void* data, int nbatch, int dir, cudaStream_t st) {
    int rc;
    if (pruned_ok(p, nbatch)) {
        cufftHandle h2, h1;
        if ((rc = get_pruned_fft(p, &h2, &h1))) return rc;
        FFT(cufftSetStream(h2, st));
        FFT(cufftSetStream(h1, st));
        if (axis3_fused<T>(p, nbatch)) {
            // in-plane passes by cuFFT on the N3 planes; axis 3 by the fused kernel, which also
            // applies phase_before (forward, on store) / conj(phase_before) (inverse, on load)
            const void* a1 = p->have_pb ? p->d_pb[0] : nullptr;
            if (dir == CUFFT_FORWARD) {
                if ((rc = exec_fft<T>(h2, data, dir))) return rc;
                rc = run_axis3<T>(p, false, a1, data, st);
                if (rc != 0) return fail(B2N_ECUDA, "axis-3 FFT launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
            } else {
                rc = run_axis3<T>(p, true, a1, data, st);
                if (rc != 0) return fail(B2N_ECUDA, "axis-3 FFT launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
                if ((rc = exec_fft<T>(h2, data, dir))) return rc;
            }
            p->lib_calls += 1;
            p->launches += 1;
            return B2N_OK;
        }
        if (dir == CUFFT_FORWARD) {
            if ((rc = exec_fft<T>(h2, data, dir))) return rc;
            if ((rc = exec_fft<T>(h1, data, dir))) return rc;
        } else {
            if ((rc = exec_fft<T>(h1, data, dir))) return rc;
            if ((rc = exec_fft<T>(h2, data, dir))) return rc;
        }
        p->lib_calls += 2;
        return B2N_OK;
    }
    cufftHandle fft;
    if ((rc = get_fft(p, nbatch, &fft))) return rc;
    FFT(cufftSetStream(fft, st));
    if ((rc = exec_fft<T>(fft, data, dir))) return rc;
    p->lib_calls += 1;
    return B2N_OK;
}

static int ensure_work(b2n_plan* p, int nbatch) {
    const size_t need = p->cplx_size() * (size_t)p->g.PK * nbatch;
    if (need <= p->work_bytes) return B2N_OK;
    if (p->d_work) {
        dev_free(p, p->d_work);
        p->d_work = nullptr;
        p->work_bytes = 0;
    }
    int rc = dev_alloc(p, &p->d_work, need);
    if (rc) return rc;
    p->work_bytes = need;
    return B2N_OK;
}

static AxisPtrs axis_ptrs(b2n_plan* p) {
    AxisPtrs ax{};
    for (int d = 0; d < 3; d++) { ax.sn[d] = p->d_sn[d]; ax.pb[d] = p->d_pb[d]; }
    return ax;
}

// image -> oversampled spectrum on the Kd grid (scale, zero-pad, FFT, phase_before)
// (smaps != nullptr: one image times nbatch coil maps, see pre_scale_pad_kernel)
template <typename T>
static int grid_fwd_t(b2n_plan* p, const void* image, void* grid, int nbatch, cudaStream_t st,
                      const void* smaps = nullptr) {
    using C = cplx_t<T>;
    const Geom& g = p->g;
    int rc;
    AxisPtrs ax = axis_ptrs(p);
    C* work = (C*)grid;
    constexpr int VEC = 32 / (int)sizeof(C);   // 32 bytes of grid per thread
    if (smaps == nullptr && inplane_own<T>(p, nbatch)) {
        // scale + zero-pad + axis 1, axis 2 on the non-zero rows, axis 3 + phase_before: three own passes
        rc = run_inplane<T>(p, false, image, nullptr, work, st);
        if (rc == 0) rc = run_axis3<T>(p, false, p->have_pb ? p->d_pb[0] : nullptr, work, st);
        if (rc != 0) return fail(B2N_ECUDA, "own FFT pass failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
        p->launches += 1;
        return B2N_OK;
    }
    if (smaps != nullptr)
        pre_scale_pad_kernel<T, VEC, true><<<grid_for(g.PK * nbatch / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, (T)p->fwd_scale, p->fwd_scale != 1.0, (const C*)image, work, nbatch, (const C*)smaps);
    else
        pre_scale_pad_kernel<T, VEC><<<grid_for(g.PK * nbatch / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, (T)p->fwd_scale, p->fwd_scale != 1.0, (const C*)image, work, nbatch);
    CU(cudaGetLastError());
    if ((rc = run_fft<T>(p, work, nbatch, CUFFT_FORWARD, st))) return rc;
    p->launches += 1;
    if (p->have_pb && !axis3_fused<T>(p, nbatch)) {
        phase_before_kernel<T, VEC><<<grid_for(g.PK * nbatch / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, 0, work, nbatch);
        CU(cudaGetLastError());
        p->launches++;
    }
    return B2N_OK;
}

// gridded spectrum -> image (conj phase_before, inverse FFT, crop, scale); grid is overwritten
// (smaps != nullptr: the nbatch coil images are combined into one, see sense_crop_combine_kernel)
template <typename T>
static int grid_adj_t(b2n_plan* p, void* grid, void* image, int nbatch, cudaStream_t st,
                      const void* smaps = nullptr) {
    using C = cplx_t<T>;
    const Geom& g = p->g;
    int rc;
    AxisPtrs ax = axis_ptrs(p);
    C* work = (C*)grid;
    if (smaps == nullptr && inplane_own<T>(p, nbatch)) {
        rc = run_axis3<T>(p, true, p->have_pb ? p->d_pb[0] : nullptr, work, st);
        if (rc == 0) rc = run_inplane<T>(p, true, nullptr, image, work, st);
        if (rc != 0) return fail(B2N_ECUDA, "own FFT pass failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
        p->launches += 1;
        return B2N_OK;
    }
    if (p->have_pb && !axis3_fused<T>(p, nbatch)) {
        constexpr int VEC = 32 / (int)sizeof(C);
        phase_before_kernel<T, VEC><<<grid_for(g.PK * nbatch / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, 1, work, nbatch);
        CU(cudaGetLastError());
        p->launches++;
    }
    if ((rc = run_fft<T>(p, work, nbatch, CUFFT_INVERSE, st))) return rc;
    if (smaps != nullptr)
        sense_crop_combine_kernel<T><<<grid_for(g.PN, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, (T)p->adj_scale, p->adj_scale != 1.0, work, (const C*)smaps, (C*)image, nbatch);
    else
        post_crop_scale_kernel<T><<<grid_for(g.PN * nbatch, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, (T)p->adj_scale, p->adj_scale != 1.0, work, (C*)image, nbatch);
    CU(cudaGetLastError());
    p->launches += 1;
    return B2N_OK;
}

static int grid_fwd(b2n_plan* p, const void* image, void* grid, int nbatch, cudaStream_t st,
                    const void* smaps = nullptr) {
    return p->precision == B2N_SINGLE ? grid_fwd_t<float>(p, image, grid, nbatch, st, smaps)
                                      : grid_fwd_t<double>(p, image, grid, nbatch, st, smaps);
}
static int grid_adj(b2n_plan* p, void* grid, void* image, int nbatch, cudaStream_t st,
                    const void* smaps = nullptr) {
    return p->precision == B2N_SINGLE ? grid_adj_t<float>(p, grid, image, nbatch, st, smaps)
                                      : grid_adj_t<double>(p, grid, image, nbatch, st, smaps);
}

extern "C" int b2n_grid_fwd(b2n_plan* p, const void* image_dev, void* grid_dev, int nbatch,
                            void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (nbatch < 1) return fail(B2N_EINVAL, "nbatch must be >= 1");
    if (image_dev == nullptr || grid_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return grid_fwd(p, image_dev, grid_dev, nbatch, (cudaStream_t)stream);
}

extern "C" int b2n_grid_adj(b2n_plan* p, void* grid_dev, void* image_dev, int nbatch, void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (nbatch < 1) return fail(B2N_EINVAL, "nbatch must be >= 1");
    if (image_dev == nullptr || grid_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return grid_adj(p, grid_dev, image_dev, nbatch, (cudaStream_t)stream);
}

static int nufft_fwd_impl(b2n_plan* p, const void* image, void* samples, int nbatch, cudaStream_t st,
                          const void* smaps = nullptr) {
    int rc;
    if ((rc = ensure_work(p, nbatch))) return rc;
    if ((rc = grid_fwd(p, image, p->d_work, nbatch, st, smaps))) return rc;
    if (p->opt_sparse_mode) return spmv_impl(p, true, p->d_work, samples, nbatch, p->d_phase_s != nullptr, st);
    return interp_fwd_impl(p, p->d_work, samples, nbatch, p->d_phase_s != nullptr, st);
}

static int nufft_adj_impl(b2n_plan* p, const void* samples, void* image, int nbatch, cudaStream_t st,
                          const void* smaps = nullptr) {
    int rc;
    if ((rc = ensure_work(p, nbatch))) return rc;
    if (p->opt_sparse_mode) rc = spmv_impl(p, false, samples, p->d_work, nbatch, p->d_phase_s != nullptr, st);
    else rc = interp_adj_impl(p, samples, p->d_work, nbatch, p->d_phase_s != nullptr, st);
    if (rc) return rc;
    return grid_adj(p, p->d_work, image, nbatch, st, smaps);
}

extern "C" int b2n_nufft_fwd(b2n_plan* p, const void* image_dev, void* samples_dev, int nbatch,
                             void* stream) {
    int rc = check_ready(p, image_dev, samples_dev, nbatch);
    if (rc) return rc;
    if (image_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return nufft_fwd_impl(p, image_dev, samples_dev, nbatch, (cudaStream_t)stream);
}

extern "C" int b2n_nufft_adj(b2n_plan* p, const void* samples_dev, void* image_dev, int nbatch,
                             void* stream) {
    int rc = check_ready(p, samples_dev, image_dev, nbatch);
    if (rc) return rc;
    if (image_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return nufft_adj_impl(p, samples_dev, image_dev, nbatch, (cudaStream_t)stream);
}

// Coil-sensitivity encoding fused around the transforms.
extern "C" int b2n_sense_fwd(b2n_plan* p, const void* image_dev, const void* smaps_dev,
                             void* samples_dev, int ncoil, void* stream) {
    int rc = check_ready(p, image_dev, samples_dev, ncoil);
    if (rc) return rc;
    if (image_dev == nullptr || smaps_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return nufft_fwd_impl(p, image_dev, samples_dev, ncoil, (cudaStream_t)stream, smaps_dev);
}

extern "C" int b2n_sense_adj(b2n_plan* p, const void* samples_dev, const void* smaps_dev,
                             void* image_dev, int ncoil, void* stream) {
    int rc = check_ready(p, samples_dev, image_dev, ncoil);
    if (rc) return rc;
    if (image_dev == nullptr || smaps_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return nufft_adj_impl(p, samples_dev, image_dev, ncoil, (cudaStream_t)stream, smaps_dev);
}

// grid[b] *= kernel (pointwise, complex) for every batch entry: the middle step of the
// Toeplitz normal operator (pad + FFT by b2n_grid_fwd, this, inverse FFT + crop by
// b2n_grid_adj on a plan with Kd = 2 Nd).
extern "C" int b2n_grid_multiply(b2n_plan* p, void* grid_dev, const void* kernel_dev, int nbatch,
                                 void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (nbatch < 1) return fail(B2N_EINVAL, "nbatch must be >= 1");
    if (grid_dev == nullptr || kernel_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    ON_DEVICE(p);
    cudaStream_t st = (cudaStream_t)stream;
    const Geom& g = p->g;
    const int nb = grid_for(g.PK * nbatch, 256, p->sm_count, 32);
    if (p->precision == B2N_SINGLE)
        grid_multiply_kernel<float><<<nb, 256, 0, st>>>(g.PK, (const cplx_t<float>*)kernel_dev,
                                                        (cplx_t<float>*)grid_dev, nbatch);
    else
        grid_multiply_kernel<double><<<nb, 256, 0, st>>>(g.PK, (const cplx_t<double>*)kernel_dev,
                                                         (cplx_t<double>*)grid_dev, nbatch);
    CU(cudaGetLastError());
    p->launches += 1;
    return B2N_OK;
}

// ---------------------------------------------------------------------------------
// Staged 3-D transforms for slab-distributed operation: the oversampled
// FFT split into its in-plane part on a range of image planes and its axis-3 part on a slab of
// grid rows, with an all-to-all between them (done by the host: SlabShardedNufft).  The
// arithmetic per element is that of b2n_grid_fwd / b2n_grid_adj.
// ---------------------------------------------------------------------------------
static int get_planes_fft(b2n_plan* p, int nz, cufftHandle* out) {
    auto it = p->fft_planes.find(nz);
    if (it != p->fft_planes.end()) {
        *out = it->second;
        return B2N_OK;
    }
    const Geom& g = p->g;
    int n2[2] = {g.K[1], g.K[0]};
    cufftHandle h;
    FFT(cufftCreate(&h));
    size_t ws = 0;
    FFT(cufftMakePlanMany(h, 2, n2, nullptr, 1, g.K[0] * g.K[1], nullptr, 1, g.K[0] * g.K[1],
                          p->precision == B2N_SINGLE ? CUFFT_C2C : CUFFT_Z2Z, nz, &ws));
    p->dev_bytes += (int64_t)ws;
    p->fft_planes[nz] = h;
    *out = h;
    return B2N_OK;
}

static int get_axis3_fft(b2n_plan* p, cufftHandle* out) {
    if (!p->fft_ax3_ready) {
        const Geom& g = p->g;
        int n1[1] = {g.K[2]};
        int embed[1] = {g.K[2]};
        size_t ws = 0;
        FFT(cufftCreate(&p->fft_ax3));
        FFT(cufftMakePlanMany(p->fft_ax3, 1, n1, embed, g.K[0] * g.K[1], 1, embed, g.K[0] * g.K[1], 1,
                              p->precision == B2N_SINGLE ? CUFFT_C2C : CUFFT_Z2Z, g.K[0] * g.K[1], &ws));
        p->dev_bytes += (int64_t)ws;
        p->fft_ax3_ready = true;
    }
    *out = p->fft_ax3;
    return B2N_OK;
}

static int check_planes(b2n_plan* p, const void* a, const void* b, int z0, int nz) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (a == nullptr || b == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (p->g.ndim != 3) return fail(B2N_EINVAL, "staged transforms are 3-D");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    if (z0 < 0 || nz < 1 || z0 + nz > p->g.N[2]) return fail(B2N_EINVAL, "plane range outside [0, Nd[2])");
    return B2N_OK;
}

template <typename T>
static int planes_fwd_t(b2n_plan* p, const void* image, int z0, int nz, void* planes, cudaStream_t st) {
    using C = cplx_t<T>;
    Geom g2 = p->g;                       // the planes [z0, z0 + nz) as a volume of their own
    g2.K[2] = nz; g2.N[2] = nz;
    g2.PK = (int64_t)g2.K[0] * g2.K[1] * nz;
    g2.PN = (int64_t)g2.N[0] * g2.N[1] * nz;
    AxisPtrs ax = axis_ptrs(p);
    ax.sn[2] += z0;
    if (inplane_own<T>(p, 1)) {
        const int e = run_inplane<T>(p, false, image, nullptr, planes, st, z0, nz);
        if (e != 0) return fail(B2N_ECUDA, "own FFT pass failed: " + std::string(cudaGetErrorString((cudaError_t)e)));
        return B2N_OK;
    }
    constexpr int VEC = 32 / (int)sizeof(C);
    pre_scale_pad_kernel<T, VEC><<<grid_for(g2.PK / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
        g2, ax, (T)p->fwd_scale, p->fwd_scale != 1.0, (const C*)image, (C*)planes, 1);
    CU(cudaGetLastError());
    cufftHandle h;
    int rc = get_planes_fft(p, nz, &h);
    if (rc) return rc;
    FFT(cufftSetStream(h, st));
    if ((rc = exec_fft<T>(h, planes, CUFFT_FORWARD))) return rc;
    p->launches += 1;
    p->lib_calls += 1;
    return B2N_OK;
}

template <typename T>
static int planes_adj_t(b2n_plan* p, void* planes, int z0, int nz, void* image, cudaStream_t st) {
    using C = cplx_t<T>;
    Geom g2 = p->g;
    g2.K[2] = nz; g2.N[2] = nz;
    g2.PK = (int64_t)g2.K[0] * g2.K[1] * nz;
    g2.PN = (int64_t)g2.N[0] * g2.N[1] * nz;
    AxisPtrs ax = axis_ptrs(p);
    ax.sn[2] += z0;
    if (inplane_own<T>(p, 1)) {
        const int e = run_inplane<T>(p, true, nullptr, image, planes, st, z0, nz);
        if (e != 0) return fail(B2N_ECUDA, "own FFT pass failed: " + std::string(cudaGetErrorString((cudaError_t)e)));
        return B2N_OK;
    }
    cufftHandle h;
    int rc = get_planes_fft(p, nz, &h);
    if (rc) return rc;
    FFT(cufftSetStream(h, st));
    if ((rc = exec_fft<T>(h, planes, CUFFT_INVERSE))) return rc;
    post_crop_scale_kernel<T><<<grid_for(g2.PN, 256, p->sm_count, 32), 256, 0, st>>>(
        g2, ax, (T)p->adj_scale, p->adj_scale != 1.0, (const C*)planes, (C*)image, 1);
    CU(cudaGetLastError());
    p->launches += 1;
    p->lib_calls += 1;
    return B2N_OK;
}

extern "C" int b2n_planes_fwd(b2n_plan* p, const void* image_planes_dev, int z0, int nz,
                              void* planes_dev, void* stream) {
    int rc = check_planes(p, image_planes_dev, planes_dev, z0, nz);
    if (rc) return rc;
    ON_DEVICE(p);
    return p->precision == B2N_SINGLE
               ? planes_fwd_t<float>(p, image_planes_dev, z0, nz, planes_dev, (cudaStream_t)stream)
               : planes_fwd_t<double>(p, image_planes_dev, z0, nz, planes_dev, (cudaStream_t)stream);
}

extern "C" int b2n_planes_adj(b2n_plan* p, void* planes_dev, int z0, int nz, void* image_planes_dev,
                              void* stream) {
    int rc = check_planes(p, planes_dev, image_planes_dev, z0, nz);
    if (rc) return rc;
    ON_DEVICE(p);
    return p->precision == B2N_SINGLE
               ? planes_adj_t<float>(p, planes_dev, z0, nz, image_planes_dev, (cudaStream_t)stream)
               : planes_adj_t<double>(p, planes_dev, z0, nz, image_planes_dev, (cudaStream_t)stream);
}

template <typename T>
static int axis3_t(b2n_plan* p, void* grid, bool inverse, cudaStream_t st) {
    using C = cplx_t<T>;
    const Geom& g = p->g;
    AxisPtrs ax = axis_ptrs(p);
    constexpr int VEC = 32 / (int)sizeof(C);
    int rc;
    if (axis3_fused<T>(p, 1)) {
        // the fused pass: planes >= Nd[2] are treated as zero on input (forward) and not
        // written (adjoint); phase_before rides along
        const void* a1 = p->have_pb ? p->d_pb[0] : nullptr;
        rc = run_axis3<T>(p, inverse, a1, grid, st);
        if (rc != 0) return fail(B2N_ECUDA, "axis-3 FFT launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
        p->launches += 1;
        return B2N_OK;
    }
    cufftHandle h;
    rc = get_axis3_fft(p, &h);
    if (rc) return rc;
    FFT(cufftSetStream(h, st));
    if (inverse && p->have_pb) {
        phase_before_kernel<T, VEC><<<grid_for(g.PK / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, 1, (C*)grid, 1);
        CU(cudaGetLastError());
        p->launches++;
    }
    if ((rc = exec_fft<T>(h, grid, inverse ? CUFFT_INVERSE : CUFFT_FORWARD))) return rc;
    p->lib_calls += 1;
    if (!inverse && p->have_pb) {
        phase_before_kernel<T, VEC><<<grid_for(g.PK / VEC + 1, 256, p->sm_count, 32), 256, 0, st>>>(
            g, ax, 0, (C*)grid, 1);
        CU(cudaGetLastError());
        p->launches++;
    }
    return B2N_OK;
}

static int axis3_entry(b2n_plan* p, void* grid, bool inverse, void* stream) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (grid == nullptr) return fail(B2N_EINVAL, "NULL array");
    if (p->g.ndim != 3) return fail(B2N_EINVAL, "staged transforms are 3-D");
    if (!p->scaling_set) return fail(B2N_ESTATE, "scaling not set");
    ON_DEVICE(p);
    return p->precision == B2N_SINGLE ? axis3_t<float>(p, grid, inverse, (cudaStream_t)stream)
                                      : axis3_t<double>(p, grid, inverse, (cudaStream_t)stream);
}

extern "C" int b2n_axis3_fwd(b2n_plan* p, void* grid_dev, void* stream) {
    return axis3_entry(p, grid_dev, false, stream);
}
extern "C" int b2n_axis3_adj(b2n_plan* p, void* grid_dev, void* stream) {
    return axis3_entry(p, grid_dev, true, stream);
}

// ---- peer-memory exchange of the slab-distributed transforms (slab_exchange.cuh)
static int slab_peers(b2n_plan* p, int world, void* const* grids, const int* row0, const int* nrows,
                      int nz, int z0, SlabPeers* out) {
    if (p == nullptr) return fail(B2N_EINVAL, "NULL plan");
    if (p->g.ndim != 3) return fail(B2N_EINVAL, "staged transforms are 3-D");
    if (world < 1 || world > kMaxPeers) return fail(B2N_EINVAL, "1 <= world <= 16");
    if (grids == nullptr || row0 == nullptr || nrows == nullptr) return fail(B2N_EINVAL, "NULL argument");
    if (z0 < 0 || nz < 1 || z0 + nz > p->g.K[2]) return fail(B2N_EINVAL, "plane range outside the grid");
    if ((p->g.K[0] * p->cplx_size()) % 16 != 0) return fail(B2N_EINVAL, "Kd[0] * sizeof(complex) must be a multiple of 16");
    out->world = world;
    for (int s = 0; s < world; s++) {
        if (grids[s] == nullptr || nrows[s] < 1 || nrows[s] > p->g.K[1] || row0[s] < 0 || row0[s] >= p->g.K[1])
            return fail(B2N_EINVAL, "bad slab description");
        out->grid[s] = grids[s];
        out->row0[s] = row0[s];
        out->nrows[s] = nrows[s];
    }
    return B2N_OK;
}

extern "C" int b2n_slab_scatter(b2n_plan* p, const void* planes_dev, int nz, int z0, int world,
                                void* const* peer_grids, const int* row0, const int* nrows,
                                void* stream) {
    SlabPeers P;
    int rc = slab_peers(p, world, peer_grids, row0, nrows, nz, z0, &P);
    if (rc) return rc;
    if (planes_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    ON_DEVICE(p);
    const int vpr = (int)(p->g.K[0] * p->cplx_size() / 16);
    int total = 0;
    for (int s = 0; s < world; s++) total += nrows[s];
    const int nb = grid_for((int64_t)nz * total * 32, 256, p->sm_count, 16);
    if (p->precision == B2N_SINGLE)
        slab_scatter_kernel<float4><<<nb, 256, 0, (cudaStream_t)stream>>>(P, (const float4*)planes_dev, nz, z0, p->g.K[1], vpr);
    else
        slab_scatter_kernel<double2><<<nb, 256, 0, (cudaStream_t)stream>>>(P, (const double2*)planes_dev, nz, z0, p->g.K[1], vpr);
    CU(cudaGetLastError());
    p->launches += 1;
    return B2N_OK;
}

extern "C" int b2n_slab_gather(b2n_plan* p, void* planes_dev, int nz, int z0, int world,
                               void* const* peer_grids, const int* row0, const int* nrows,
                               void* stream) {
    SlabPeers P;
    int rc = slab_peers(p, world, peer_grids, row0, nrows, nz, z0, &P);
    if (rc) return rc;
    if (planes_dev == nullptr) return fail(B2N_EINVAL, "NULL array");
    ON_DEVICE(p);
    const int vpr = (int)(p->g.K[0] * p->cplx_size() / 16);
    const int nb = grid_for((int64_t)nz * p->g.K[1] * 32, 256, p->sm_count, 16);
    if (p->precision == B2N_SINGLE)
        slab_gather_kernel<float4><<<nb, 256, 0, (cudaStream_t)stream>>>(P, (float4*)planes_dev, nz, z0, p->g.K[1], vpr);
    else
        slab_gather_kernel<double2><<<nb, 256, 0, (cudaStream_t)stream>>>(P, (double2*)planes_dev, nz, z0, p->g.K[1], vpr);
    CU(cudaGetLastError());
    p->launches += 1;
    return B2N_OK;
}
