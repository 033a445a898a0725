// Own passes of the pruned oversampled 3-D FFT, fused with what surrounds them.
//
// Two kernels:
//   fft_lines_kernel  (second half of this file) -- ALL THREE passes of a single-volume 3-D
//       transform whose lengths have a compile-time radix schedule (128 .. 1024): axis 1 with
//       x * sn * scale and the zero padding fused (adjoint: crop * scale * sn), axis 2 on the
//       non-zero rows only, axis 3 with the zero padding, phase_before and the crop fused.
//   fft_axis3_kernel  (first half) -- the axis-3 pass alone with a RUN-TIME mixed-radix schedule,
//       for every other K3 = 2^a 3^b; the in-plane passes then stay with cuFFT (b200nufft.cu:
//       run_fft) on the N3 non-zero planes.
// What rides along, instead of the reference's separate full-grid sweeps:
//   forward  (_nufft.py:1325-1371): x * sn is applied as the image rows are loaded; padding rows,
//            columns and planes are created in shared memory / registers and never read;
//            phase_before is multiplied into the output of the axis-3 pass as it is stored;
//   adjoint  (_nufft.py:1519-1572): conj(phase_before) is applied as the axis-3 lines are loaded,
//            only the planes / rows / columns that survive the crop are stored, and the axis-1
//            pass writes image[n] = grid[n] * scale * sn[n] directly.
// A CTA owns COLS adjacent lines (every global access is a run of COLS complex values, or a
// contiguous row in the axis-1 mode), keeps them in shared memory and runs a Stockham autosort
// FFT on them: rows of COLS values per shared-memory access (conflict-free), twiddles from an
// L-entry table evaluated in double precision on the host.  Unnormalised in both directions,
// like cuFFT.
#pragma once
#include "aux_kernels.cuh"
#include "common.cuh"
#include "dispatch.h"

namespace b2n {

template <typename T>
__device__ __forceinline__ cplx_t<T> cmulc(cplx_t<T> a, cplx_t<T> b) {      // a * b
    return make_c<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <typename T> __device__ __forceinline__ cplx_t<T> cadd(cplx_t<T> a, cplx_t<T> b) { return make_c<T>(a.x + b.x, a.y + b.y); }
template <typename T> __device__ __forceinline__ cplx_t<T> csub(cplx_t<T> a, cplx_t<T> b) { return make_c<T>(a.x - b.x, a.y - b.y); }
// float: one packed FADD2 per complex add / subtract (each half an IEEE add: same bits)
template <> __device__ __forceinline__ float2 cadd<float>(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
template <> __device__ __forceinline__ float2 csub<float>(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
// (+i) * a for the inverse transform, (-i) * a for the forward one
template <typename T, bool INV> __device__ __forceinline__ cplx_t<T> rot90(cplx_t<T> a) {
    return INV ? make_c<T>(-a.y, a.x) : make_c<T>(a.y, -a.x);
}

// length-4 DFT of (a0, a1, a2, a3) in place
template <typename T, bool INV>
__device__ __forceinline__ void dft4(cplx_t<T>& a0, cplx_t<T>& a1, cplx_t<T>& a2, cplx_t<T>& a3) {
    using C = cplx_t<T>;
    const C s02 = cadd<T>(a0, a2), d02 = csub<T>(a0, a2), s13 = cadd<T>(a1, a3);
    const C id13 = rot90<T, INV>(csub<T>(a1, a3));
    a0 = cadd<T>(s02, s13);
    a1 = cadd<T>(d02, id13);
    a2 = csub<T>(s02, s13);
    a3 = csub<T>(d02, id13);
}

// length-R DFT of v[0..R-1] in place, R in {2, 3, 4, 6, 8}
template <typename T, bool INV, int R>
__device__ __forceinline__ void butterfly(cplx_t<T>* v) {
    using C = cplx_t<T>;
    if (R == 2) {
        const C a = v[0];
        v[0] = cadd<T>(a, v[1]);
        v[1] = csub<T>(a, v[1]);
    } else if (R == 3) {
        // w = exp(-+ 2 pi i / 3) = (-1/2, -+ sqrt(3)/2)
        const T hs = (T)0.86602540378443864676 * (INV ? (T)1 : (T)-1);
        const C s12 = cadd<T>(v[1], v[2]), d12 = csub<T>(v[1], v[2]);
        const C m = make_c<T>(v[0].x - (T)0.5 * s12.x, v[0].y - (T)0.5 * s12.y);
        const C rr = make_c<T>(-hs * d12.y, hs * d12.x);                // i * hs * d12
        v[0] = cadd<T>(v[0], s12);
        v[1] = cadd<T>(m, rr);
        v[2] = csub<T>(m, rr);
    } else if (R == 4) {
        dft4<T, INV>(v[0], v[1], v[2], v[3]);
    } else if (R == 6) {
        // two length-3 DFTs (even / odd inputs) + one radix-2 stage with w6^q
        const T hs = (T)0.86602540378443864676 * (INV ? (T)1 : (T)-1);
        C e[3], o[3];
        {
            const C s12 = cadd<T>(v[2], v[4]), d12 = csub<T>(v[2], v[4]);
            const C m = make_c<T>(v[0].x - (T)0.5 * s12.x, v[0].y - (T)0.5 * s12.y);
            const C rr = make_c<T>(-hs * d12.y, hs * d12.x);
            e[0] = cadd<T>(v[0], s12); e[1] = cadd<T>(m, rr); e[2] = csub<T>(m, rr);
        }
        {
            const C s12 = cadd<T>(v[3], v[5]), d12 = csub<T>(v[3], v[5]);
            const C m = make_c<T>(v[1].x - (T)0.5 * s12.x, v[1].y - (T)0.5 * s12.y);
            const C rr = make_c<T>(-hs * d12.y, hs * d12.x);
            o[0] = cadd<T>(v[1], s12); o[1] = cadd<T>(m, rr); o[2] = csub<T>(m, rr);
        }
        // w6 = exp(-+ 2 pi i / 6) = (1/2, -+ sqrt(3)/2), w6^2 = (-1/2, -+ sqrt(3)/2)
        const C t1 = make_c<T>((T)0.5 * o[1].x - hs * o[1].y, (T)0.5 * o[1].y + hs * o[1].x);
        const C t2 = make_c<T>((T)-0.5 * o[2].x - hs * o[2].y, (T)-0.5 * o[2].y + hs * o[2].x);
        v[0] = cadd<T>(e[0], o[0]); v[3] = csub<T>(e[0], o[0]);
        v[1] = cadd<T>(e[1], t1);   v[4] = csub<T>(e[1], t1);
        v[2] = cadd<T>(e[2], t2);   v[5] = csub<T>(e[2], t2);
    } else {                                     // R == 8: two length-4 DFTs + one radix-2 stage
        dft4<T, INV>(v[0], v[2], v[4], v[6]);
        dft4<T, INV>(v[1], v[3], v[5], v[7]);
        // odd half times w8^q, w8 = exp(-+ 2 pi i / 8)
        const T h = (T)0.70710678118654752440;
        const C b1 = v[3], b3 = v[7];
        // w8^1 = (h, -+h), w8^2 = -+i, w8^3 = (-h, -+h)
        const C t1 = INV ? make_c<T>(h * (b1.x - b1.y), h * (b1.x + b1.y))
                         : make_c<T>(h * (b1.x + b1.y), h * (b1.y - b1.x));
        const C t2 = rot90<T, INV>(v[5]);
        const C t3 = INV ? make_c<T>(-h * (b3.x + b3.y), h * (b3.x - b3.y))
                         : make_c<T>(h * (b3.y - b3.x), -h * (b3.x + b3.y));
        const C e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
        v[0] = cadd<T>(e0, o0); v[4] = csub<T>(e0, o0);
        v[1] = cadd<T>(e1, t1); v[5] = csub<T>(e1, t1);
        v[2] = cadd<T>(e2, t2); v[6] = csub<T>(e2, t2);
        v[3] = cadd<T>(e3, t3); v[7] = csub<T>(e3, t3);
    }
}

// One Stockham pass of radix R over the CTA's COLS columns: butterfly j reads rows
// j + r*L/R, multiplies by the twiddles W^(r*k*L/(Ns*R)), k = j mod Ns, and writes rows
// (j div Ns)*Ns*R + k + r*Ns.
template <typename T, bool INV, int COLS, int R, int NT>
__device__ __forceinline__ void axis3_pass(const cplx_t<T>* __restrict__ in, cplx_t<T>* __restrict__ out,
                                           const cplx_t<T>* __restrict__ twS, int L, int Ns, int sh,
                                           bool pow2, int tid) {
    using C = cplx_t<T>;
    const int c = tid % COLS;
    const int Tn = L / R;
    const int tstep = Tn / Ns;                       // = L / (Ns * R)
    for (int j = tid / COLS; j < Tn; j += NT / COLS) {
        const int k = pow2 ? (j & (Ns - 1)) : j % Ns;
        const int jq = pow2 ? (j >> sh) : j / Ns;
        const int j0 = jq * Ns * R + k;
        C v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = in[(j + r * Tn) * COLS + c];
        if (k != 0) {
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmulc<T>(v[r], twS[r * k * tstep]);
        }
        butterfly<T, INV, R>(v);
#pragma unroll
        for (int r = 0; r < R; r++) out[(j0 + r * Ns) * COLS + c] = v[r];
    }
}

// INV = false: exp(-i...) (cuFFT forward); INV = true: exp(+i...)
// Persistent CTAs: a CTA walks tiles of COLS columns; the global loads of the NEXT tile are
// issued into registers (PF values per thread) before the passes of the current one, so that
// their latency overlaps the shared-memory passes instead of idling the SM.
template <typename T, bool INV, int COLS, int NT, int PF>
__global__ void __launch_bounds__(NT, sizeof(T) == 4 ? 3 : 2)
fft_axis3_kernel(Axis3Plan ap, int64_t plane, int K1, int NZ, int64_t ntiles,
                 const cplx_t<T>* __restrict__ tw,
                 const T* __restrict__ a1, const T* __restrict__ a2, const T* __restrict__ a3,
                 cplx_t<T>* __restrict__ data) {
    using C = cplx_t<T>;
    extern __shared__ __align__(16) unsigned char fft3_smem[];
    const int L = ap.L;
    C* bufA = (C*)fft3_smem;
    C* bufB = bufA + (size_t)L * COLS;
    C* twS = bufB + (size_t)L * COLS;
    const int tid = threadIdx.x;
    const bool phase = a1 != nullptr;
    constexpr int RL = NT / COLS;                    // rows per load round

    for (int e = tid; e < L; e += NT) {
        C w = tw[e];
        if (INV) w.y = -w.y;
        twS[e] = w;
    }
    // this thread's column inside a tile (fixed: NT % COLS == 0) and its first row
    const int c = tid % COLS;
    const int r0 = tid / COLS;
    const int rows_in = INV ? L : NZ;                // forward: only the non-zero planes are read
    const int rows_out = INV ? NZ : L;               // adjoint: only the planes that survive the crop
    C pf[PF];
    int64_t tile = blockIdx.x;
    if (tile < ntiles) {
        const int64_t col = tile * COLS + c;
#pragma unroll
        for (int i = 0; i < PF; i++) {
            const int k3 = r0 + i * RL;
            pf[i] = make_c<T>(0, 0);
            if (k3 < rows_in && col < plane) pf[i] = data[(int64_t)k3 * plane + col];
        }
    }
    for (; tile < ntiles; tile += gridDim.x) {
        const int64_t col = tile * COLS + c;
        const bool col_ok = col < plane;
        T a12 = (T)0;
        if (phase && col_ok) {
            const int k1 = (int)(col % K1), k2 = (int)(col / K1);
            a12 = a1[k1] + a2[k2];                   // the reference's summation order
        }
        // ---- registers -> shared memory (the rest of a forward column is zero padding)
#pragma unroll
        for (int i = 0; i < PF; i++) {
            const int k3 = r0 + i * RL;
            if (k3 < L) {
                C v = pf[i];
                if (INV && phase && k3 < rows_in) {
                    T sn, co;
                    sincos_t(a12 + a3[k3], &sn, &co);
                    v = make_c<T>(v.x * co + v.y * sn, v.y * co - v.x * sn);      // * conj(phase)
                }
                bufA[k3 * COLS + c] = v;
            }
        }
        __syncthreads();
        // ---- next tile's loads go out now; they land while this tile is transformed
        {
            const int64_t nt = tile + gridDim.x;
            const int64_t ncol = nt * COLS + c;
            if (nt < ntiles) {
#pragma unroll
                for (int i = 0; i < PF; i++) {
                    const int k3 = r0 + i * RL;
                    pf[i] = make_c<T>(0, 0);
                    if (k3 < rows_in && ncol < plane) pf[i] = data[(int64_t)k3 * plane + ncol];
                }
            }
        }
        // ---- Stockham passes (radix dispatch is uniform; Ns is a power of two until the first
        // pass with a factor 3, so the index split is a shift and a mask there)
        C* in = bufA;
        C* out = bufB;
        int Ns = 1;
        for (int p = 0; p < ap.npass; p++) {
            const int R = ap.radix[p];
            const bool pow2 = (Ns & (Ns - 1)) == 0;
            const int sh = 31 - __clz(Ns);
            if (R == 8) axis3_pass<T, INV, COLS, 8, NT>(in, out, twS, L, Ns, sh, pow2, tid);
            else if (R == 6) axis3_pass<T, INV, COLS, 6, NT>(in, out, twS, L, Ns, sh, pow2, tid);
            else if (R == 4) axis3_pass<T, INV, COLS, 4, NT>(in, out, twS, L, Ns, sh, pow2, tid);
            else if (R == 2) axis3_pass<T, INV, COLS, 2, NT>(in, out, twS, L, Ns, sh, pow2, tid);
            else axis3_pass<T, INV, COLS, 3, NT>(in, out, twS, L, Ns, sh, pow2, tid);
            __syncthreads();
            C* t = in; in = out; out = t;
            Ns *= R;
        }
        // ---- store
        if (col_ok) {
            for (int k3 = r0; k3 < rows_out; k3 += RL) {
                C v = in[k3 * COLS + c];
                if (!INV && phase) {
                    T sn, co;
                    sincos_t(a12 + a3[k3], &sn, &co);
                    v = make_c<T>(v.x * co - v.y * sn, v.x * sn + v.y * co);      // * phase
                }
                data[(int64_t)k3 * plane + col] = v;
            }
        }
        __syncthreads();                             // the buffers are rewritten by the next tile
    }
}

// ---------------------------------------------------------------------------------------------
// Fixed-schedule variant for the transform lengths oversampled MRI grids actually have
// (K = 1.5 N or 2 N with N a power of two).  Same data flow as fft_axis3_kernel, but the length
// and the radix schedule are template parameters, so every index split, twiddle stride and
// loop bound of the passes is a compile-time constant (the run-time schedule spent 150 warp
// instructions per output value, two thirds of them index arithmetic and sincosf), and
// phase_before costs one accurate sincos per COLUMN and tile instead of one per value:
//     exp(i * fl(a12 + a3[k])) = exp(i a12) * exp(i a3[k]) * exp(-i err),
// with err the rounding error of the float sum, recovered exactly by TwoSum (|err| <= ulp/2 of
// an angle of a few thousand radians, so exp(-i err) = 1 - err^2/2 - i err to 1e-13), and
// exp(i a3[k]) from a table filled once per CTA.  The angle that is exponentiated is still the
// reference's float32 sum (_nufft.py:703-715); only the evaluation is factored.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fixed_npass(int L) {
    return (L == 128 || L == 192 || L == 256 || L == 384 || L == 512) ? 3 : ((L == 768 || L == 1024) ? 4 : 0);
}
__host__ __device__ constexpr int fixed_radix(int L, int p) {
    // 128 = 8 4 4, 192 = 8 8 3, 256 = 8 8 4, 384 = 8 8 6, 512 = 8 8 8, 768 = 8 8 4 3, 1024 = 8 8 4 4
    if (p < 2) return (L == 128 && p == 1) ? 4 : 8;
    if (p == 2) return L == 192 ? 3 : (L == 384 ? 6 : (L == 512 ? 8 : 4));
    return L == 768 ? 3 : 4;
}
__host__ __device__ constexpr int fixed_ns(int L, int p) {          // product of the radices before pass p
    int ns = 1;
    for (int q = 0; q < p; q++) ns *= fixed_radix(L, q);
    return ns;
}

// SRC: 0 = inputs from shared memory `in`, 1 = inputs already in registers (`regs`, butterfly
// order: value i * R + r is input r of butterfly slot + i * SLOTS).  DST: 0 = outputs to shared
// memory `out`; 1 = handed to `sink(row, value)` (the last pass of the strided mode writes global
// memory itself).
template <typename T, bool INV, int L, int PITCH, int SLOTS, int P, int SRC = 0, int DST = 0, typename Sink = int>
__device__ __forceinline__ void fixed_pass(const cplx_t<T>* __restrict__ in, cplx_t<T>* __restrict__ out,
                                           const cplx_t<T>* __restrict__ twS, int slot, int c,
                                           const cplx_t<T>* regs = nullptr, Sink sink = Sink()) {
    using C = cplx_t<T>;
    constexpr int R = fixed_radix(L, P);
    constexpr int Ns = fixed_ns(L, P);
    constexpr int Tn = L / R;
    constexpr int tstep = Tn / Ns;
    constexpr int ROUNDS = (Tn + SLOTS - 1) / SLOTS;
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
        const int j = slot + i * SLOTS;
        if (Tn % SLOTS != 0 && j >= Tn) break;
        const int k = j % Ns;
        const int j0 = j * R - k * (R - 1);          // (j div Ns) * Ns * R + k
        C v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if constexpr (SRC == 1) v[r] = regs[i * R + r];
            else v[r] = in[(j + r * Tn) * PITCH + c];
        }
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmulc<T>(v[r], twS[r * k * tstep]);
        }
        butterfly<T, INV, R>(v);
#pragma unroll
        for (int r = 0; r < R; r++) {
            if constexpr (DST == 1) sink(j0 + r * Ns, v[r]);
            else out[(j0 + r * Ns) * PITCH + c] = v[r];
        }
    }
}

// v * exp(+-i * fl(a12 + a3k)) given e12 = exp(i a12) and e3k = exp(i a3k) (see above)
template <typename T, bool CONJ>
__device__ __forceinline__ cplx_t<T> phase_mul(cplx_t<T> v, T a12, cplx_t<T> e12, T a3k, cplx_t<T> e3k) {
    const T s = a12 + a3k;
    const T bb = s - a12;
    const T err = (a12 - (s - bb)) + (a3k - bb);     // a12 + a3k = s + err exactly
    const T d = -err;
    const T h = (T)1 - (T)0.5 * d * d;
    const cplx_t<T> c0 = cmulc<T>(e12, e3k);
    const T cr = c0.x * h - d * c0.y;
    T ci = c0.y * h + d * c0.x;
    if (CONJ) ci = -ci;
    return make_c<T>(v.x * cr - v.y * ci, v.x * ci + v.y * cr);
}

// float: the same product with the angle reduced to [-pi, pi] by a two-term Cody-Waite step
// (exact to ~1e-7 rad for the few-thousand-radian angles of phase_before) and the hardware
// sine / cosine (2^-21.4 absolute on that interval): ~5e-7 in the phase, 1/20 of the parity
// budget, for 10 instructions instead of 22.  The angle itself is still the reference's float32 sum.
template <bool CONJ>
__device__ __forceinline__ float2 phase_mul(float2 v, float a12, float2, float a3k, float2) {
    const float s = a12 + a3k;
    const float n = rintf(s * 0.15915494309189535f);
    float r = fmaf(n, -6.2831854820251465f, s);
    r = fmaf(n, 1.7484555e-7f, r);                   // 2 pi = 6.2831854820251465 - 1.7484555e-7
    const float cr = __cosf(r);
    const float ci = CONJ ? -__sinf(r) : __sinf(r);
    return make_float2(v.x * cr - v.y * ci, v.x * ci + v.y * cr);
}
template <bool CONJ>
__device__ __forceinline__ double2 phase_mul(double2 v, double a12, double2 e12, double a3k, double2 e3k) {
    return phase_mul<double, CONJ>(v, a12, e12, a3k, e3k);
}

// One kernel for the three passes of the pruned oversampled FFT ("lines" = the 1-D transforms):
//   MODE 0  lines strided in memory, COLS adjacent lines per tile (axis 3: row stride K1*K2, one
//           outer block, phase_before fused; axis 2: row stride K1, one outer block per plane)
//   MODE 1  lines contiguous in memory (axis 1): a tile is COLS consecutive grid rows, loaded and
//           stored along the row (coalesced) through a transposed shared-memory tile of pitch
//           COLS + 1; the forward reads the IMAGE rows and applies sn * scale on the way in (the
//           zero padding is created in shared memory), the adjoint writes the cropped, scaled image
//           rows: the scale/zero-pad and crop/scale passes of _nufft.py:1325-1331 / :1560-1572 ride
//           along instead of sweeping the padded planes.
// nz: forward = leading non-zero entries of a line (the rest is zero padding and is not read),
// adjoint = leading entries that survive the crop (the rest is not stored).
template <typename T> struct LineArgs {
    cplx_t<T>* data;              // oversampled grid
    const cplx_t<T>* tw;          // L twiddles exp(-2 pi i t / L)
    int nz;
    int64_t ntiles;
    // MODE 0
    int64_t row_stride, outer_stride, inner_extent, tiles_per_outer;
    const T *a1, *a2, *a3;        // phase_before angles per axis (axis 3 only), or nullptr
    int K1;
    // MODE 1
    cplx_t<T>* image;             // [..][NL2][N1] image rows (read by the forward, written by the adjoint)
    int NL2, K2, N1;              // lines per plane (= N2), grid rows per plane, image row length
    int64_t nlines;
    const double *sn1, *sn2, *sn3;
    T scale;
    int apply_scale;
};

template <typename T, bool INV, int L, int COLS, int NT, int MODE>
__global__ void __launch_bounds__(NT)
fft_lines_kernel(const LineArgs<T> la) {
    using C = cplx_t<T>;
    constexpr int NP = fixed_npass(L);
    constexpr int SLOTS = NT / COLS;                 // butterflies in flight per pass round
    constexpr int PITCH = MODE == 1 ? COLS + 1 : COLS;
    constexpr int LPW = COLS / (NT / 32);            // MODE 1: lines per warp
    constexpr int VPL = L / 32;                      // MODE 1: values per line and lane
    // first pass (MODE 0: fed from registers): radix, butterflies per line, rounds per thread
    constexpr int R0 = fixed_radix(L, 0);
    constexpr int Tn0 = L / R0;
    constexpr int ROUNDS0 = (Tn0 + SLOTS - 1) / SLOTS;
    constexpr int PF = MODE == 1 ? LPW * VPL : ROUNDS0 * R0;      // values per thread and tile
    static_assert(MODE == 0 || (L % 32 == 0 && COLS % (NT / 32) == 0 && LPW >= 1), "row mode tiling");
    extern __shared__ __align__(16) unsigned char fft3_smem[];
    C* bufA = (C*)fft3_smem;
    C* bufB = bufA + L * PITCH;
    C* twS = bufB + L * PITCH;
    C* e3S = twS + L;
    T* a3S = (T*)(e3S + L);
    const int tid = threadIdx.x;
    const int lane = tid & 31, wib = tid >> 5;
    const bool phase = MODE == 0 && la.a1 != nullptr;
    for (int e = tid; e < L; e += NT) {
        C w = la.tw[e];
        if (INV) w.y = -w.y;
        twS[e] = w;
        if (phase) {
            const T a = la.a3[e];
            T sn, co;
            sincos_t(a, &sn, &co);
            a3S[e] = a;
            e3S[e] = make_c<T>(co, sn);
        }
    }
    __syncthreads();
    const int c = tid % COLS;                        // this thread's line in the tile (passes, MODE 0 I/O)
    const int r0 = tid / COLS;                       // ... and its butterfly slot
    const int rows_in = INV ? L : la.nz;
    const int rows_out = INV ? la.nz : L;

    // ---- loads of one tile into registers, issued one tile ahead.  MODE 0: exactly the inputs of
    // this thread's first-pass butterflies (rows j + r * L / R0 of its line: 128-byte runs across
    // the 16 lanes of a row), so the first pass runs out of registers and the tile never has to be
    // staged in shared memory first; rows >= rows_in are the zero padding and are not read.
    C pf[PF];
    auto prefetch = [&](int64_t tile) {
        if constexpr (MODE == 0) {
            const int64_t o = tile / la.tiles_per_outer;
            const int64_t col = (tile - o * la.tiles_per_outer) * COLS + c;
            const C* src = la.data + (o * la.outer_stride + col);
#pragma unroll
            for (int i = 0; i < ROUNDS0; i++) {
#pragma unroll
                for (int r = 0; r < R0; r++) {
                    const int k = r0 + i * SLOTS + r * Tn0;
                    pf[i * R0 + r] = make_c<T>(0, 0);
                    if ((Tn0 % SLOTS == 0 || r0 + i * SLOTS < Tn0) && k < rows_in && col < la.inner_extent)
                        pf[i * R0 + r] = src[(int64_t)k * la.row_stride];
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < LPW; q++) {
                const int64_t line = tile * COLS + wib * LPW + q;
                const int64_t z = line / la.NL2;
                const int k2 = (int)(line - z * la.NL2);
                const C* src = INV ? la.data + ((z * la.K2 + k2) * (int64_t)L)
                                   : la.image + ((z * la.NL2 + k2) * (int64_t)la.N1);
#pragma unroll
                for (int i = 0; i < VPL; i++) {
                    const int k = lane + 32 * i;
                    pf[q * VPL + i] = make_c<T>(0, 0);
                    if (k < (INV ? L : la.N1) && line < la.nlines) pf[q * VPL + i] = src[k];
                }
            }
        }
    };
    int64_t tile = blockIdx.x;
    if (tile < la.ntiles) prefetch(tile);
    for (; tile < la.ntiles; tile += gridDim.x) {
        if constexpr (MODE == 0) {
            const int64_t o = tile / la.tiles_per_outer;
            const int64_t col = (tile - o * la.tiles_per_outer) * COLS + c;
            const bool col_ok = col < la.inner_extent;
            T a12 = (T)0;
            C e12 = make_c<T>(1, 0);
            if (phase && col_ok) {
                const int k1 = (int)(col % la.K1), k2 = (int)(col / la.K1);
                a12 = la.a1[k1] + la.a2[k2];         // the reference's summation order
                T sn, co;
                sincos_t(a12, &sn, &co);
                e12 = make_c<T>(co, sn);
            }
            if (INV && phase) {
                // conj(phase_before) on the way in
#pragma unroll
                for (int i = 0; i < ROUNDS0; i++)
#pragma unroll
                    for (int r = 0; r < R0; r++) {
                        const int k = r0 + i * SLOTS + r * Tn0;
                        if (k < L) pf[i * R0 + r] = phase_mul<true>(pf[i * R0 + r], a12, e12, a3S[k], e3S[k]);
                    }
            }
            // ---- first pass out of the registers
            fixed_pass<T, INV, L, PITCH, SLOTS, 0, 1, 0>(nullptr, bufA, twS, r0, c, pf);
            __syncthreads();
            // ---- next tile's loads go out now; they land while this tile is transformed
            if (tile + gridDim.x < la.ntiles) prefetch(tile + gridDim.x);
            // ---- middle passes in shared memory, last pass straight to global memory (rows of the
            // 16 lines are 128-byte runs; phase_before on the way out; rows >= rows_out are cropped)
            C* dst = la.data + (o * la.outer_stride + col);
            auto sink = [&](int k, C v) {
                if (col_ok && k < rows_out) {
                    if (!INV && phase) v = phase_mul<false>(v, a12, e12, a3S[k], e3S[k]);
                    dst[(int64_t)k * la.row_stride] = v;
                }
            };
            if constexpr (NP == 3) {
                fixed_pass<T, INV, L, PITCH, SLOTS, 1>(bufA, bufB, twS, r0, c);
                __syncthreads();
                fixed_pass<T, INV, L, PITCH, SLOTS, 2, 0, 1>(bufB, nullptr, twS, r0, c, nullptr, sink);
            } else {
                fixed_pass<T, INV, L, PITCH, SLOTS, 1>(bufA, bufB, twS, r0, c);
                __syncthreads();
                fixed_pass<T, INV, L, PITCH, SLOTS, 2>(bufB, bufA, twS, r0, c);
                __syncthreads();
                fixed_pass<T, INV, L, PITCH, SLOTS, 3, 0, 1>(bufA, nullptr, twS, r0, c, nullptr, sink);
            }
            // (no barrier here: the next tile's first pass writes bufA, last read before the
            // barrier above; its second pass writes bufB after the next tile's first barrier,
            // which every thread reaches only after its last-pass reads of bufB / bufA)
            if constexpr (NP == 4) __syncthreads();  // 4 passes: the last one reads bufA
        } else {
            // ---- registers -> transposed shared-memory tile (the rest of a forward line is zero padding)
#pragma unroll
            for (int q = 0; q < LPW; q++) {
                const int lc = wib * LPW + q;
                const int64_t line = tile * COLS + lc;
                const int64_t z = line / la.NL2;
                const int k2 = (int)(line - z * la.NL2);
#pragma unroll
                for (int i = 0; i < VPL; i++) {
                    const int k = lane + 32 * i;
                    C v = pf[q * VPL + i];
                    if (!INV && k < la.N1 && line < la.nlines) {
                        // x * sn (sn = ((s1*s2)*s3) in double, then cast), then the transform scale
                        const T st = (T)((la.sn1[k] * la.sn2[k2]) * la.sn3[z]);
                        v = make_c<T>(v.x * st, v.y * st);
                        if (la.apply_scale) { v.x *= la.scale; v.y *= la.scale; }
                    }
                    bufA[k * PITCH + lc] = v;
                }
            }
            __syncthreads();
            if (tile + gridDim.x < la.ntiles) prefetch(tile + gridDim.x);
            fixed_pass<T, INV, L, PITCH, SLOTS, 0>(bufA, bufB, twS, r0, c);
            __syncthreads();
            fixed_pass<T, INV, L, PITCH, SLOTS, 1>(bufB, bufA, twS, r0, c);
            __syncthreads();
            fixed_pass<T, INV, L, PITCH, SLOTS, 2>(bufA, bufB, twS, r0, c);
            __syncthreads();
            if constexpr (NP == 4) {
                fixed_pass<T, INV, L, PITCH, SLOTS, 3>(bufB, bufA, twS, r0, c);
                __syncthreads();
            }
            const C* res = NP == 4 ? bufA : bufB;
#pragma unroll
            for (int q = 0; q < LPW; q++) {
                const int lc = wib * LPW + q;
                const int64_t line = tile * COLS + lc;
                if (line >= la.nlines) continue;
                const int64_t z = line / la.NL2;
                const int k2 = (int)(line - z * la.NL2);
                if (INV) {
                    // image[n] = grid[n] * adj_scale * sn[n]
                    C* dst = la.image + ((z * la.NL2 + k2) * (int64_t)la.N1);
                    for (int k = lane; k < la.N1; k += 32) {
                        C v = res[k * PITCH + lc];
                        const T st = (T)((la.sn1[k] * la.sn2[k2]) * la.sn3[z]);
                        if (la.apply_scale) { v.x *= la.scale; v.y *= la.scale; }
                        dst[k] = make_c<T>(v.x * st, v.y * st);
                    }
                } else {
                    C* dst = la.data + ((z * la.K2 + k2) * (int64_t)L);
#pragma unroll 4
                    for (int k = lane; k < L; k += 32) dst[k] = res[k * PITCH + lc];
                }
            }
            __syncthreads();                         // the buffers are rewritten by the next tile
        }
    }
}

// Tile width: 128-byte runs per global access (16 float / 8 double lines, 256 threads: measured
// 0.25 ms per axis-3 pass on the bench grid against 0.29 ms with 64-byte runs and 0.36 ms with
// the run-time schedule), 64-byte runs where two 128-byte buffers would not fit in shared memory
template <typename T, int L, int MODE> struct LineCfg {
    static constexpr size_t smem_for(int cols) {
        return (size_t)(2 * L * (cols + (MODE == 1 ? 1 : 0)) + 2 * L) * 2 * sizeof(T) + (size_t)L * sizeof(T);
    }
    static constexpr bool WIDE = smem_for(128 / (2 * (int)sizeof(T))) <= 200 * 1024;
    static constexpr int COLS = (WIDE ? 128 : 64) / (2 * (int)sizeof(T));
    // MODE 0: <= 24 values per thread.  MODE 1: a warp owns LPW whole lines (LPW a power of two,
    // about 24 values per thread)
    static constexpr int lpw() {
        int l = 24 / (L / 32), q = 1;
        while (2 * q <= l && 2 * q <= COLS) q *= 2;
        return q;
    }
    static constexpr int NT = MODE == 1 ? 32 * COLS / lpw() : (L >= 768 ? 256 : 128) * (WIDE ? 2 : 1);
    static constexpr size_t smem = smem_for(COLS);
};

template <typename T, bool INV, int L, int MODE>
static int fft_lines_launch_L(const LineArgs<T>& la_in, int sm_count, int max_smem, cudaStream_t st, bool* done) {
    constexpr int COLS = LineCfg<T, L, MODE>::COLS;
    constexpr int NT = LineCfg<T, L, MODE>::NT;
    const size_t smem = LineCfg<T, L, MODE>::smem;
    if (smem > (size_t)max_smem) return 0;
    LineArgs<T> la = la_in;
    if (MODE == 0) {
        la.tiles_per_outer = (la.inner_extent + COLS - 1) / COLS;
        la.ntiles *= la.tiles_per_outer;             // (ntiles came in as the number of outer blocks)
    } else {
        la.ntiles = (la.nlines + COLS - 1) / COLS;
    }
    auto k = fft_lines_kernel<T, INV, L, COLS, NT, MODE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, NT, smem);
    if (e != cudaSuccess) return (int)e;
    int64_t nb = (int64_t)sm_count * (per_sm < 1 ? 1 : per_sm);      // persistent: resident CTAs only
    if (nb > la.ntiles) nb = la.ntiles;
    if (nb < 1) { *done = true; return 0; }
    k<<<(unsigned)nb, NT, smem, st>>>(la);
    *done = true;
    return (int)cudaGetLastError();
}

// returns 0 or a cudaError_t; *done = false when L has no fixed schedule or the tile does not fit
template <typename T, int MODE>
static int fft_lines_launch(int L, bool inverse, const LineArgs<T>& la, int sm_count, int max_smem,
                            cudaStream_t st, bool* done) {
    *done = false;
#define B2N_FL(LL)                                                                                  \
    case LL:                                                                                        \
        return inverse ? fft_lines_launch_L<T, true, LL, MODE>(la, sm_count, max_smem, st, done)    \
                       : fft_lines_launch_L<T, false, LL, MODE>(la, sm_count, max_smem, st, done);
    switch (L) {
        B2N_FL(128) B2N_FL(192) B2N_FL(256) B2N_FL(384) B2N_FL(512) B2N_FL(768) B2N_FL(1024)
        default: return 0;
    }
#undef B2N_FL
}

// shared memory the fixed-schedule kernel needs for length L in mode MODE (0: no fixed schedule)
template <typename T, int MODE> static size_t fft_lines_smem(int L) {
    switch (L) {
        case 128: return LineCfg<T, 128, MODE>::smem;
        case 192: return LineCfg<T, 192, MODE>::smem;
        case 256: return LineCfg<T, 256, MODE>::smem;
        case 384: return LineCfg<T, 384, MODE>::smem;
        case 512: return LineCfg<T, 512, MODE>::smem;
        case 768: return LineCfg<T, 768, MODE>::smem;
        case 1024: return LineCfg<T, 1024, MODE>::smem;
        default: return 0;
    }
}

template <typename T> struct Axis3Cfg {
    // 8 (4) columns = 64-byte runs per global access, 256 threads, up to PF values per thread
    // prefetched in registers; at K3 = 384 the two buffers take 48 KB: 3 CTAs per SM.
    static constexpr int COLS = sizeof(T) == 4 ? 8 : 4;
    static constexpr int NT = 256;
    static constexpr int PF = sizeof(T) == 4 ? 16 : 8;
    static size_t smem(int L) { return (size_t)(2 * L * COLS + L) * 2 * sizeof(T); }
    static bool fits(int L) { return (L * COLS + NT - 1) / NT <= PF; }
};

// returns 0 or a cudaError_t
template <typename T>
static int fft_axis3_launch(const Axis3Plan& ap, const Geom& g, bool inverse, const void* tw,
                            const void* a1, const void* a2, const void* a3, void* data,
                            int sm_count, cudaStream_t st) {
    using C = cplx_t<T>;
    constexpr int COLS = Axis3Cfg<T>::COLS;
    constexpr int NT = Axis3Cfg<T>::NT;
    constexpr int PF = Axis3Cfg<T>::PF;
    const int64_t plane = (int64_t)g.K[0] * g.K[1];
    const size_t smem = Axis3Cfg<T>::smem(ap.L);
    const int64_t ntiles = (plane + COLS - 1) / COLS;
    int64_t nb = (int64_t)sm_count * (sizeof(T) == 4 ? 3 : 2);   // persistent: resident CTAs only
    if (nb > ntiles) nb = ntiles;
    cudaError_t e;
    if (inverse) {
        auto k = fft_axis3_kernel<T, true, COLS, NT, PF>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<(unsigned)nb, NT, smem, st>>>(ap, plane, g.K[0], g.N[2], ntiles, (const C*)tw, (const T*)a1,
                                          (const T*)a2, (const T*)a3, (C*)data);
    } else {
        auto k = fft_axis3_kernel<T, false, COLS, NT, PF>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<(unsigned)nb, NT, smem, st>>>(ap, plane, g.K[0], g.N[2], ntiles, (const C*)tw, (const T*)a1,
                                          (const T*)a2, (const T*)a3, (C*)data);
    }
    return (int)cudaGetLastError();
}

}  // namespace b2n
