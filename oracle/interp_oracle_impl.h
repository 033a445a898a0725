/*
 * TEST INFRASTRUCTURE ONLY -- never linked or imported by the product path.
 *
 * CPU restatement of the reference's table-lookup interpolators, generic in the
 * number of axes (1..3) and in the table kind (real / complex), operating on
 * interleaved complex arrays in column-major (first axis fastest) layout.
 *
 * Follows /root/reference/mrrt/nufft/_extensions/c/nufft_table.template.c:
 *   window origin   koff = 1 + floor(t - J/2.)  in double      (:865-867, :62, :118)
 *   table argument  p = (t - k) * L  in REAL, n = floor(p), alf = p - n (:870-872)
 *   coefficient     (1 - alf) * h[n] + alf * h[n+1], table centred by
 *                   ncenter = floor(J*L/2)                       (:843-854, :873)
 *   periodic wrap   kmod = k - K*floor(k/K)                     (:29, :874-875)
 *   forward         nested partial sums  j3{ j2{ j1{} } }       (:869-919)
 *   adjoint         nested partial products v3 -> v2 -> ck +=   (:1122-1163)
 *   complex tables  full complex multiply; adjoint conjugates   (:76-77, :189-190,
 *                                                                :1021-1022)
 * This file is included once per REAL type by interp_oracle.c.
 */

static inline int NAME(wrap)(int k, int K)
{
    int r = k % K;
    return r < 0 ? r + K : r;
}

/* Table lookup order: 1 = linear interpolation between entries (the only mode of the
 * reference's C code), 0 = the entry at floor(p) -- "order 0" of the reference's GPU templates
 * (cuda/jinja/table_1d_forward.jinja:40-47, table_2d_forward.jinja:61-67, :84-90 and the
 * adjoint / 3-D siblings; cuda/cupy.py:96-98).  Process-wide switch set from Python. */
static int NAME(g_order) = 1;
void NAME(set_table_order)(int order) { NAME(g_order) = order; }

/* coefficient for one tap: linear interpolation of the centred table.
 * The reference indexes h[n] and h[n+1] unguarded; n can be -1 (weight 1-alf ~ 1e-12,
 * when `t - J/2.` rounds up to an integer in the window-origin formula) and n+1 can be
 * one past the end (weight alf == 0): undefined behaviour there.  Both out-of-table
 * entries count as 0 here, so that the checker never returns heap garbage. */
static inline void NAME(coef)(const REAL *h, int h_cplx, int ncenter, int tlen, REAL t,
                              int k, int L, REAL *cr, REAL *ci)
{
    const REAL p = (t - (REAL)k) * (REAL)L;
    const int n = (int)floor((double)p);
    const REAL alf = p - (REAL)n;
    const long i0 = (long)ncenter + n;
    const int ok0 = i0 >= 0 && i0 < tlen, ok1 = i0 + 1 >= 0 && i0 + 1 < tlen;
    if (NAME(g_order) == 0) {
        *cr = ok0 ? (h_cplx ? h[2 * i0] : h[i0]) : 0;
        *ci = (ok0 && h_cplx) ? h[2 * i0 + 1] : 0;
        return;
    }
    if (h_cplx) {
        *cr = (1 - alf) * (ok0 ? h[2 * i0] : 0) + alf * (ok1 ? h[2 * (i0 + 1)] : 0);
        *ci = (1 - alf) * (ok0 ? h[2 * i0 + 1] : 0) + alf * (ok1 ? h[2 * (i0 + 1) + 1] : 0);
    } else {
        *cr = (1 - alf) * (ok0 ? h[i0] : 0) + alf * (ok1 ? h[i0 + 1] : 0);
        *ci = 0;
    }
}

/*
 * Forward gather.  K,J: per-axis sizes (axes >= ndim must hold 1).
 * h[d]: table for axis d, J[d]*L+1 entries (interleaved complex if h_cplx).
 * tm: [M, ndim] column-major.  ck: [K1*K2*K3] complex.  fm: [M] complex out.
 */
void NAME(interp_fwd)(int ndim, const int *K, const int *J, int L,
                      const REAL *h1, const REAL *h2, const REAL *h3,
                      int h_cplx, const REAL *tm, long M, const REAL *ck,
                      REAL *fm)
{
    const int J1 = J[0], J2 = ndim > 1 ? J[1] : 1, J3 = ndim > 2 ? J[2] : 1;
    const int K1 = K[0], K2 = ndim > 1 ? K[1] : 1, K3 = ndim > 2 ? K[2] : 1;
    const int nc1 = (J1 * L) / 2, nc2 = (J2 * L) / 2, nc3 = (J3 * L) / 2;
    long mm;
#pragma omp parallel for schedule(dynamic, 1000)
    for (mm = 0; mm < M; mm++) {
        const REAL t1 = tm[mm];
        const REAL t2 = ndim > 1 ? tm[M + mm] : 0;
        const REAL t3 = ndim > 2 ? tm[2 * M + mm] : 0;
        const int koff1 = 1 + (int)floor((double)t1 - J1 / 2.);
        const int koff2 = ndim > 1 ? 1 + (int)floor((double)t2 - J2 / 2.) : 0;
        const int koff3 = ndim > 2 ? 1 + (int)floor((double)t3 - J3 / 2.) : 0;
        REAL s3r = 0, s3i = 0;
        for (int j3 = 0; j3 < J3; j3++) {
            REAL c3r = 1, c3i = 0;
            if (ndim > 2)
                NAME(coef)(h3, h_cplx, nc3, J3 * L + 1, t3, koff3 + j3, L, &c3r, &c3i);
            const long k3 = NAME(wrap)(koff3 + j3, K3);
            REAL s2r = 0, s2i = 0;
            for (int j2 = 0; j2 < J2; j2++) {
                REAL c2r = 1, c2i = 0;
                if (ndim > 1)
                    NAME(coef)(h2, h_cplx, nc2, J2 * L + 1, t2, koff2 + j2, L, &c2r, &c2i);
                const long k2 = NAME(wrap)(koff2 + j2, K2);
                const long row = (k3 * K2 + k2) * K1;
                REAL s1r = 0, s1i = 0;
                for (int j1 = 0; j1 < J1; j1++) {
                    REAL c1r, c1i;
                    NAME(coef)(h1, h_cplx, nc1, J1 * L + 1, t1, koff1 + j1, L, &c1r, &c1i);
                    const long kk = row + NAME(wrap)(koff1 + j1, K1);
                    const REAL gr = ck[2 * kk], gi = ck[2 * kk + 1];
                    s1r += c1r * gr - c1i * gi;
                    s1i += c1r * gi + c1i * gr;
                }
                s2r += c2r * s1r - c2i * s1i;
                s2i += c2r * s1i + c2i * s1r;
            }
            s3r += c3r * s2r - c3i * s2i;
            s3i += c3r * s2i + c3i * s2r;
        }
        fm[2 * mm] = s3r;
        fm[2 * mm + 1] = s3i;
    }
}

/* Adjoint scatter-add: ck is zeroed first (template.c:965-966, :1107-1108). */
void NAME(interp_adj)(int ndim, const int *K, const int *J, int L,
                      const REAL *h1, const REAL *h2, const REAL *h3,
                      int h_cplx, const REAL *tm, long M, const REAL *fm,
                      REAL *ck)
{
    const int J1 = J[0], J2 = ndim > 1 ? J[1] : 1, J3 = ndim > 2 ? J[2] : 1;
    const int K1 = K[0], K2 = ndim > 1 ? K[1] : 1, K3 = ndim > 2 ? K[2] : 1;
    const int nc1 = (J1 * L) / 2, nc2 = (J2 * L) / 2, nc3 = (J3 * L) / 2;
    memset(ck, 0, sizeof(REAL) * 2 * (size_t)K1 * K2 * K3);
    for (long mm = 0; mm < M; mm++) {
        const REAL t1 = tm[mm];
        const REAL t2 = ndim > 1 ? tm[M + mm] : 0;
        const REAL t3 = ndim > 2 ? tm[2 * M + mm] : 0;
        const REAL fr = fm[2 * mm], fi = fm[2 * mm + 1];
        const int koff1 = 1 + (int)floor((double)t1 - J1 / 2.);
        const int koff2 = ndim > 1 ? 1 + (int)floor((double)t2 - J2 / 2.) : 0;
        const int koff3 = ndim > 2 ? 1 + (int)floor((double)t3 - J3 / 2.) : 0;
        for (int j3 = 0; j3 < J3; j3++) {
            REAL c3r = 1, c3i = 0;
            if (ndim > 2)
                NAME(coef)(h3, h_cplx, nc3, J3 * L + 1, t3, koff3 + j3, L, &c3r, &c3i);
            const long k3 = NAME(wrap)(koff3 + j3, K3);
            const REAL v3r = c3r * fr + c3i * fi;
            const REAL v3i = c3r * fi - c3i * fr;
            for (int j2 = 0; j2 < J2; j2++) {
                REAL c2r = 1, c2i = 0;
                if (ndim > 1)
                    NAME(coef)(h2, h_cplx, nc2, J2 * L + 1, t2, koff2 + j2, L, &c2r, &c2i);
                const long k2 = NAME(wrap)(koff2 + j2, K2);
                const long row = (k3 * K2 + k2) * K1;
                const REAL v2r = c2r * v3r + c2i * v3i;
                const REAL v2i = c2r * v3i - c2i * v3r;
                for (int j1 = 0; j1 < J1; j1++) {
                    REAL c1r, c1i;
                    NAME(coef)(h1, h_cplx, nc1, J1 * L + 1, t1, koff1 + j1, L, &c1r, &c1i);
                    const long kk = row + NAME(wrap)(koff1 + j1, K1);
                    ck[2 * kk] += c1r * v2r + c1i * v2i;
                    ck[2 * kk + 1] += c1r * v2i - c1i * v2r;
                }
            }
        }
    }
}
