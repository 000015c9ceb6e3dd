/*
 * wifi_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C99 restatement, in x87 `long double _Complex` like the reference, of
 * the 802.11 channel-estimation hot path of usmandroid/80211ParallelEstimation.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's library.  The product (the CUDA library
 * behind include/wifi_b200.h) never links, imports or calls it.
 *
 * Pinning (see tests/test_oracle.py, tests/golden/make_golden.py):
 *   - orc_lt_ls / orc_ps_linear / orc_ps_cubic / orc_ps_sinc, orc_multiply,
 *     orc_hermitian_as_written, orc_addition_as_written, orc_identity,
 *     orc_outer, orc_inverse_cofactor: pinned against the reference's own
 *     sequential C code compiled in place (oracle/_ref, built by oracle/Makefile
 *     from /root/reference/{main.c,utils.c}) and against its outputs on the
 *     inputs.h frame stored in tests/golden/.
 *   - orc_equalize and the *_matlab estimators: pinned against the reference's
 *     matlab.mat goldens (H_EST_*, eq_symbols).
 *   - orc_mmse_*: the reference's own PS_MMSE body (main.c:148-212) returns NaN on
 *     every input (utils.c:6 hermitian is not a conjugate, utils.c:117 addition
 *     ignores M2, utils.c:543-569 has no pivoting) and matlab.mat holds no MMSE
 *     output, so there is no estimator-level reference output.  The oracle
 *     restates the intended formula (WiFi_channel_estimation_PS_MMSE.m:16-33 ==
 *     the north-star H = R (R + s2 (X X^H)^-1)^-1 (rx/tx)) and is pinned
 *     ROUTINE BY ROUTINE: against PS_MMSE composed from the reference's own
 *     compiled multiply / multiplyVxVeqM / identity / inverse as the .m text
 *     prescribes (tests/golden/mmse_ref_composed.npz: 3e-13 on a full-rank
 *     covariance, 5e-9 on the inputs.h frame = the accuracy of the reference's
 *     un-pivoted cofactor inverse there), and against a 40-digit mpmath
 *     evaluation and the rank-1 closed form (tests/golden/make_golden.py).
 *
 * Data convention of every entry point: complex arrays are interleaved
 * (re, im) `double`; arithmetic inside is `long double _Complex`; matrices are
 * dense row-major n x m (the reference's row-pointer tables point into exactly
 * such a block, utils.c:817-835).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long double _Complex ldc;

#define NSC 53      /* SAMPUTIL  utils.h:13 */
#define NBLK 15     /* OFDMBLK   utils.h:15 */
#define PIL0 5      /* P0..P3    utils.h:16-19 */
#define PIL1 19
#define PIL2 33
#define PIL3 47
#define DCBIN 26

static inline ldc ld_get(const double *a, long i) { return (long double)a[2 * i] + (long double)a[2 * i + 1] * I; }
static inline void ld_put(double *a, long i, ldc v) { a[2 * i] = (double)creall(v); a[2 * i + 1] = (double)cimagl(v); }

/* ------------------------------------------------------------------ */
/* Estimators, C semantics (block vectors of 53 sub-carriers)          */
/* ------------------------------------------------------------------ */

/* main.c:66-75.  c = Re(tx) - Im(tx) is a REAL scalar (not a conjugate);
 * H = (c*rx)/(c*tx); NaN when Re(tx) == Im(tx); H[26] = 0. */
void orc_lt_ls(const double *tx_pre, const double *rx_pre, double *H, long n_frames)
{
    for (long f = 0; f < n_frames; ++f) {
        const double *tx = tx_pre + 2 * NSC * f, *rx = rx_pre + 2 * NSC * f;
        double *h = H + 2 * NSC * f;
        for (int i = 0; i < 26; ++i) {
            ldc t1 = ld_get(tx, i), t2 = ld_get(tx, i + 27);
            ldc c1 = creall(t1) - cimagl(t1);
            ldc c2 = creall(t2) - cimagl(t2);
            ld_put(h, i, (c1 * ld_get(rx, i)) / (c1 * t1));
            ld_put(h, i + 27, (c2 * ld_get(rx, i + 27)) / (c2 * t2));
        }
        ld_put(h, DCBIN, 0.0L);
    }
}

static void pilot_ls(const double *tx, const double *rx, ldc hp[4])
{
    static const int P[4] = {PIL0, PIL1, PIL2, PIL3};
    for (int i = 0; i < 4; ++i) hp[i] = ld_get(rx, P[i]) / ld_get(tx, P[i]);   /* main.c:82-84 */
}

/* main.c:77-101.  frame_stride = complex elements between consecutive frames
 * (53 for stacked block vectors, 795 for whole frames whose block 0 is used). */
void orc_ps_linear(const double *tx, const double *rx, long frame_stride, double *H, long n_frames)
{
    const long double delta = PIL1 - PIL0;
    for (long f = 0; f < n_frames; ++f) {
        ldc hp[4];
        pilot_ls(tx + 2 * frame_stride * f, rx + 2 * frame_stride * f, hp);
        double *h = H + 2 * NSC * f;
        for (int i = 0; i < NSC; ++i) {
            long double alpha;
            ldc v;
            if (i < PIL1)      { alpha = (i - PIL0) / delta; v = hp[0] + ((hp[1] - hp[0]) * alpha); }
            else if (i < PIL2) { alpha = (i - PIL1) / delta; v = hp[1] + ((hp[2] - hp[1]) * alpha); }
            else               { alpha = (i - PIL2) / delta; v = hp[2] + ((hp[3] - hp[2]) * alpha); }  /* :93-99, both last branches */
            ld_put(h, i, v);
        }
    }
}

/* main.c:103-122.  Newton form, EVERY divided difference / 14 (sic). */
void orc_ps_cubic(const double *tx, const double *rx, long frame_stride, double *H, long n_frames)
{
    const long double delta = PIL1 - PIL0;
    for (long f = 0; f < n_frames; ++f) {
        ldc hp[4];
        pilot_ls(tx + 2 * frame_stride * f, rx + 2 * frame_stride * f, hp);
        ldc f0 = hp[0];
        ldc f01 = (hp[1] - hp[0]) / delta, f12 = (hp[2] - hp[1]) / delta, f23 = (hp[3] - hp[2]) / delta;
        ldc f012 = (f12 - f01) / delta, f123 = (f23 - f12) / delta;
        ldc f0123 = (f123 - f012) / delta;
        double *h = H + 2 * NSC * f;
        for (int k = 0; k < NSC; ++k)
            ld_put(h, k, f0 + f01 * (k - PIL0) + f012 * (k - PIL0) * (k - PIL1) + f0123 * (k - PIL0) * (k - PIL1) * (k - PIL2));
    }
}

/* utils.c:727-733 */
static double ref_sinc(double x) { return x != 0 ? sin(M_PI * x) / (M_PI * x) : 1; }

/* main.c:124-146.  Pilot LS and partial sums held in `double complex` (sic). */
void orc_ps_sinc(const double *tx, const double *rx, long frame_stride, double *H, long n_frames)
{
    const long double delta = PIL1 - PIL0;
    for (long f = 0; f < n_frames; ++f) {
        ldc hpl[4];
        double _Complex hp[4];
        pilot_ls(tx + 2 * frame_stride * f, rx + 2 * frame_stride * f, hpl);
        for (int i = 0; i < 4; ++i) hp[i] = (double _Complex)hpl[i];
        double *h = H + 2 * NSC * f;
        for (int k = 0; k < NSC; ++k) {
            double a = (k - PIL0) / delta, b = (k - PIL1) / delta, c = (k - PIL2) / delta, d = (k - PIL3) / delta;
            double _Complex s1 = hp[0] * ref_sinc(a), s2 = hp[1] * ref_sinc(b), s3 = hp[2] * ref_sinc(c), s4 = hp[3] * ref_sinc(d);
            ldc v = s1 + s2 + s3 + s4;
            ld_put(h, k, v);
        }
    }
}

/* ------------------------------------------------------------------ */
/* MATLAB semantics (second parity mode; survey 8(f)-2) + equalizer    */
/* ------------------------------------------------------------------ */

/* WiFi_channel_estimation_PS_{Linear,Cubic,Sinc}.m: per-block estimate for
 * blocks 1..4 (0-based 0..3), averaged; Cubic with true spans 14/28/42.
 * tx/rx are whole frames [n][15][53].  which: 0 linear, 1 cubic, 2 sinc. */
void orc_ps_matlab(int which, const double *tx, const double *rx, double *H, long n_frames)
{
    for (long f = 0; f < n_frames; ++f) {
        ldc acc[NSC];
        for (int k = 0; k < NSC; ++k) acc[k] = 0;
        for (int b = 0; b < 4; ++b) {
            ldc hp[4];
            pilot_ls(tx + 2 * (NSC * NBLK * f + NSC * b), rx + 2 * (NSC * NBLK * f + NSC * b), hp);
            for (int k = 0; k < NSC; ++k) {
                ldc v;
                if (which == 0) {
                    int s = k < PIL1 ? 0 : (k < PIL2 ? 1 : 2);
                    static const int P[4] = {PIL0, PIL1, PIL2, PIL3};
                    long double alpha = (long double)(k - P[s]) / 14.0L;
                    v = hp[s] + ((hp[s + 1] - hp[s]) * alpha);
                } else if (which == 1) {
                    ldc f01 = (hp[1] - hp[0]) / 14.0L, f12 = (hp[2] - hp[1]) / 14.0L, f23 = (hp[3] - hp[2]) / 14.0L;
                    ldc f012 = (f12 - f01) / 28.0L, f123 = (f23 - f12) / 28.0L;
                    ldc f0123 = (f123 - f012) / 42.0L;
                    v = hp[0] + f01 * (k - PIL0) + f012 * (k - PIL0) * (k - PIL1) + f0123 * (k - PIL0) * (k - PIL1) * (k - PIL2);
                } else {
                    long double a = (k - PIL0) / 14.0L, bb = (k - PIL1) / 14.0L, c = (k - PIL2) / 14.0L, d = (k - PIL3) / 14.0L;
#define SINCL(x) ((x) != 0 ? sinl(M_PIl * (x)) / (M_PIl * (x)) : 1.0L)
                    v = hp[0] * SINCL(a) + hp[1] * SINCL(bb) + hp[2] * SINCL(c) + hp[3] * SINCL(d);
                }
                acc[k] += v;
            }
        }
        for (int k = 0; k < NSC; ++k) ld_put(H + 2 * NSC * f, k, acc[k] / 4.0L);
    }
}

/* WiFi_Equalization.m:1-9.  rx [n][15][53]; H_lt, H_ps [n][53]; eq [n][15][53].
 * Block i (1-based): Hu = ((15-i)/15) H_lt + (i/15) H_ps; eq = rx/Hu; DC bin 0. */
void orc_equalize(const double *rx, const double *H_lt, const double *H_ps, double *eq, long n_frames)
{
    for (long f = 0; f < n_frames; ++f)
        for (int b = 0; b < NBLK; ++b) {
            long double wl = (long double)(NBLK - (b + 1)) / NBLK, wp = (long double)(b + 1) / NBLK;
            for (int k = 0; k < NSC; ++k) {
                long e = NSC * NBLK * f + NSC * b + k;
                if (k == DCBIN) { ld_put(eq, e, 0.0L); continue; }
                ldc hu = wl * ld_get(H_lt, NSC * f + k) + wp * ld_get(H_ps, NSC * f + k);
                ld_put(eq, e, ld_get(rx, e) / hu);
            }
        }
}

/* ------------------------------------------------------------------ */
/* utils.c routines                                                    */
/* ------------------------------------------------------------------ */

/* utils.c:16-31: res[c][d] = sum_k M1[c][k]*M2[k][d], k ascending.  Returns -1
 * (and writes nothing) on a dimension mismatch, where the reference prints. */
int orc_multiply(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    if (c1 != r2) return -1;
    for (int c = 0; c < r1; ++c)
        for (int d = 0; d < c2; ++d) {
            ldc sum = 0;
            for (int k = 0; k < r2; ++k) sum = sum + ld_get(M1, (long)c * c1 + k) * ld_get(M2, (long)k * c2 + d);
            ld_put(res, (long)c * c2 + d, sum);
        }
    return 0;
}

/* utils.c:3-7 as written: res[c][r] = Re(M[r][c]) - Im(M[r][c])  (real-valued, sic) */
void orc_hermitian_as_written(const double *M, int row, int col, double *res)
{
    for (int r = 0; r < row; ++r)
        for (int c = 0; c < col; ++c) {
            ldc m = ld_get(M, (long)r * col + c);
            ld_put(res, (long)c * row + r, creall(m) - cimagl(m));
        }
}

/* the intended conjugate transpose */
void orc_conj_transpose(const double *M, int row, int col, double *res)
{
    for (int r = 0; r < row; ++r)
        for (int c = 0; c < col; ++c) ld_put(res, (long)c * row + r, conjl(ld_get(M, (long)r * col + c)));
}

/* utils.c:55-65: res[r][c] = M1[r][0]*M2[0][c] */
int orc_outer(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    if (c1 != r2) return -1;
    for (int r = 0; r < r1; ++r)
        for (int c = 0; c < c2; ++c) ld_put(res, (long)r * c2 + c, ld_get(M1, (long)r * c1) * ld_get(M2, c));
    return 0;
}

/* utils.c:84-93 */
void orc_identity(double *Id, int size, double scalar)
{
    for (int r = 0; r < size; ++r)
        for (int c = 0; c < size; ++c) ld_put(Id, (long)r * size + c, r == c ? (ldc)scalar : (ldc)0.0L);
}

/* utils.c:111-121 as written: res = M1 + M1 (M2 ignored, sic) */
int orc_addition_as_written(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    (void)M2;
    if (r1 != r2 || c1 != c2) return -1;
    for (long i = 0; i < (long)r1 * c1; ++i) ld_put(res, i, ld_get(M1, i) + ld_get(M1, i));
    return 0;
}

int orc_add(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    if (r1 != r2 || c1 != c2) return -1;
    for (long i = 0; i < (long)r1 * c1; ++i) ld_put(res, i, ld_get(M1, i) + ld_get(M2, i));
    return 0;
}

/* utils.c:543-569 determinant_impl_rec, restated without the per-level
 * allocation: eliminate on element [0][0] with NO pivoting, det = m00 * det(sub),
 * order 2 closes with the 2x2 formula.  Same operation order as the recursion
 * (the product unwinds from the innermost level outwards). */
static ldc det_nopivot(ldc *m, int order, int ld)
{
    /* m is overwritten */
    ldc piv[64];
    int lvl = 0, n = order;
    ldc *a = m;
    while (n > 2) {
        ldc *s = a + ld + 1; /* sub-matrix starts at [1][1] of current */
        for (int i = 1; i < n; ++i)
            for (int j = 1; j < n; ++j) a[i * ld + j] = a[i * ld + j] - (a[i * ld] * a[j] / a[0]);
        piv[lvl++] = a[0];
        a = s;
        --n;
    }
    ldc det = (n == 1) ? a[0] : a[0] * a[ld + 1] - a[1] * a[ld];
    while (lvl > 0) det = piv[--lvl] * det;
    return det;
}

/* utils.c:141-170 + GetMinor :440-459: Y[i][j] = (-1)^(i+j) det(minor_{j,i}) / det(A).
 * O(n^5); seconds for n = 53.  order <= 64. */
int orc_inverse_cofactor(const double *A, int order, double *Y)
{
    if (order < 1 || order > 64) return -1;
    int n = order;
    ldc *w = (ldc *)malloc(sizeof(ldc) * n * n);
    if (!w) return -2;
    for (long i = 0; i < (long)n * n; ++i) w[i] = ld_get(A, i);
    ldc det = 1.0L / det_nopivot(w, n, n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            int rc = 0;
            for (int r = 0; r < n; ++r) {
                if (r == j) continue;
                int cc = 0;
                for (int c = 0; c < n; ++c) {
                    if (c == i) continue;
                    w[rc * (n - 1) + cc] = ld_get(A, (long)r * n + c);
                    ++cc;
                }
                ++rc;
            }
            ldc y = det * det_nopivot(w, n - 1, n - 1);
            if ((i + j) % 2 == 1) y = (-1) * y;
            ld_put(Y, (long)i * n + j, y);
        }
    free(w);
    return 0;
}

/* LU with partial pivoting on [A | B] in long double, then back-substitution: the numerically sound
 * inverse / solve the intended MMSE needs (the reference has none).  A n x n, B n x m, both overwritten;
 * on return B = A^-1 B.  Returns -1 if singular.  (Forward elimination + back-substitution, not Gauss-Jordan:
 * for the ill-conditioned R + D systems of the MMSE path the Jordan variant's forward error in the signal
 * directions is ~cond times larger once H = R z is formed.) */
static int gj_solve(ldc *A, ldc *B, int n, int m)
{
    for (int k = 0; k < n; ++k) {
        int p = k;
        long double best = cabsl(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            long double v = cabsl(A[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        if (best == 0.0L) return -1;
        if (p != k) {
            for (int j = 0; j < n; ++j) { ldc t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int j = 0; j < m; ++j) { ldc t = B[k * m + j]; B[k * m + j] = B[p * m + j]; B[p * m + j] = t; }
        }
        for (int i = k + 1; i < n; ++i) {
            ldc l = A[i * n + k] / A[k * n + k];
            if (l == 0) continue;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
            for (int j = 0; j < m; ++j) B[i * m + j] -= l * B[k * m + j];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        for (int j = 0; j < m; ++j) {
            ldc s = B[k * m + j];
            for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * B[c * m + j];
            B[k * m + j] = s / A[k * n + k];
        }
    }
    return 0;
}

int orc_inverse_gj(const double *A, int n, double *Y)
{
    ldc *a = (ldc *)malloc(sizeof(ldc) * n * n), *b = (ldc *)malloc(sizeof(ldc) * n * n);
    if (!a || !b) { free(a); free(b); return -2; }
    for (long i = 0; i < (long)n * n; ++i) { a[i] = ld_get(A, i); b[i] = 0; }
    for (int i = 0; i < n; ++i) b[i * n + i] = 1;
    int rc = gj_solve(a, b, n, n);
    if (rc == 0) for (long i = 0; i < (long)n * n; ++i) ld_put(Y, i, b[i]);
    free(a); free(b);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Intended MMSE                                                       */
/* ------------------------------------------------------------------ */

/* W = R (R + diag(d))^-1, R Hermitian 53x53, d real [53] (= s2/|x_k|^2).
 * Computed as W^H = (R + diag d)^-H R^H via one pivoted solve. */
int orc_mmse_filter(const double *R, const double *d, double *W)
{
    const int n = NSC;
    ldc A[NSC * NSC], B[NSC * NSC];
    /* W = R A^-1  <=>  W^T = A^-T R^T */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            ldc r = ld_get(R, i * n + j);
            A[j * n + i] = r + (i == j ? (ldc)(long double)d[i] : (ldc)0);
            B[j * n + i] = r;
        }
    int rc = gj_solve(A, B, n, n);
    if (rc) return rc;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) ld_put(W, i * n + j, B[j * n + i]);
    return 0;
}

/* H[f] = W H_ls[f]  (multiply utils.c:16-31 over all frames, k ascending) */
void orc_mmse_apply(const double *W, const double *H_ls, double *H, long n_frames)
{
    for (long f = 0; f < n_frames; ++f)
        for (int r = 0; r < NSC; ++r) {
            ldc sum = 0;
            for (int k = 0; k < NSC; ++k) sum = sum + ld_get(W, r * NSC + k) * ld_get(H_ls, NSC * f + k);
            ld_put(H, NSC * f + r, sum);
        }
}

/* Per-frame north-star form:  y = rx/tx, A = R + diag(s2_f/|tx_k|^2), solve A z = y, H = R z.
 * Rl: 53x53 long double complex (shared), or NULL with hls != NULL for the per-frame rank-1 R_f = h h^H. */
static int mmse_perframe_ld(const ldc *Rl, const double *hls, const double *tx, const double *rx, const double *sigma2, double *H, long n_frames)
{
    const int n = NSC;
    ldc *A = (ldc *)malloc(sizeof(ldc) * n * n), *Rf = (ldc *)malloc(sizeof(ldc) * n * n);
    if (!A || !Rf) { free(A); free(Rf); return -2; }
    int rc = 0;
    for (long f = 0; f < n_frames && !rc; ++f) {
        ldc z[NSC];
        const ldc *Ruse = Rl;
        if (hls) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) Rf[i * n + j] = ld_get(hls, n * f + i) * conjl(ld_get(hls, n * f + j));
            Ruse = Rf;
        }
        for (int i = 0; i < n; ++i) {
            ldc x = ld_get(tx, n * f + i);
            long double ax2 = creall(x) * creall(x) + cimagl(x) * cimagl(x);
            for (int j = 0; j < n; ++j) A[i * n + j] = Ruse[i * n + j];
            A[i * n + i] += (long double)sigma2[f] / ax2;
            z[i] = ld_get(rx, n * f + i) / x;
        }
        rc = gj_solve(A, z, n, 1);
        if (rc) break;
        for (int i = 0; i < n; ++i) {
            ldc sum = 0;
            for (int k = 0; k < n; ++k) sum = sum + Ruse[i * n + k] * z[k];
            ld_put(H, n * f + i, sum);
        }
    }
    free(A); free(Rf);
    return rc;
}

/* shared R; tx, rx [n][53] (block vectors), sigma2 [n] */
int orc_mmse_perframe(const double *R, const double *tx, const double *rx, const double *sigma2, double *H, long n_frames)
{
    ldc *Rl = (ldc *)malloc(sizeof(ldc) * NSC * NSC);
    if (!Rl) return -2;
    for (int i = 0; i < NSC * NSC; ++i) Rl[i] = ld_get(R, i);
    int rc = mmse_perframe_ld(Rl, NULL, tx, rx, sigma2, H, n_frames);
    free(Rl);
    return rc;
}

/* Frames in the C calling convention of main.c:148 (tx, rx block vectors, ow2, H_EST_LS = LT_LS estimate):
 * R = F (F^-1 H_ls)(F^-1 H_ls)^H F^H = H_ls H_ls^H (main.c:186-189 intent; F is the 53-point DFT of
 * main.c:22-26), then the per-frame north-star form above.  ow2 [n], H_ls [n][53]. */
int orc_mmse_cconv_batch(const double *tx, const double *rx, const double *ow2, const double *H_ls, double *H, long n_frames)
{
    return mmse_perframe_ld(NULL, H_ls, tx, rx, ow2, H, n_frames);
}
int orc_mmse_cconv(const double *tx, const double *rx, double ow2, const double *H_ls, double *H)
{
    return mmse_perframe_ld(NULL, H_ls, tx, rx, &ow2, H, 1);
}

/* The MATLAB text itself (WiFi_channel_estimation_PS_MMSE.m:16-33) for ONE block,
 * with the explicit 53-point F, Rhh_t = ifft(H)ifft(H)', Rhy = Rhh_t F' X (X not
 * conjugated, as written), Ryy = X F Rhh_t F' X' + ow2 I, H = F Rhy Ryy^-1 rx. */
int orc_mmse_matlab_block(const double *tx, const double *rx, double ow2, const double *H_ls, double *H)
{
    const int n = NSC;
    size_t sz = sizeof(ldc) * n * n;
    ldc *F = (ldc *)malloc(sz), *Rt = (ldc *)malloc(sz), *T1 = (ldc *)malloc(sz), *G = (ldc *)malloc(sz), *Ryy = (ldc *)malloc(sz);
    ldc h[NSC], rhs[NSC], out[NSC];
    if (!F || !Rt || !T1 || !G || !Ryy) return -2;
    for (int t = 0; t < n; ++t)
        for (int f = 0; f < n; ++f) F[t * n + f] = cexpl(-2.0L * I * M_PIl * (long double)t * (long double)f / (long double)n);
    for (int t = 0; t < n; ++t) { /* h = ifft(H_ls) = F^H H_ls / n */
        ldc s = 0;
        for (int f = 0; f < n; ++f) s += conjl(F[f * n + t]) * ld_get(H_ls, f);
        h[t] = s / (long double)n;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Rt[i * n + j] = h[i] * conjl(h[j]);
    /* T1 = Rhh_t F' */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            ldc s = 0;
            for (int k = 0; k < n; ++k) s += Rt[i * n + k] * conjl(F[j * n + k]);
            T1[i * n + j] = s;
        }
    /* G = F T1  (= F Rhh_t F') */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            ldc s = 0;
            for (int k = 0; k < n; ++k) s += F[i * n + k] * T1[k * n + j];
            G[i * n + j] = s;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            Ryy[i * n + j] = ld_get(tx, i) * G[i * n + j] * conjl(ld_get(tx, j)) + (i == j ? (ldc)(long double)ow2 : (ldc)0);
    for (int i = 0; i < n; ++i) rhs[i] = ld_get(rx, i);
    int rc = gj_solve(Ryy, rhs, n, 1);
    if (!rc) {
        for (int i = 0; i < n; ++i) { /* H = G X rhs  (F Rhy = G X, X unconjugated) */
            ldc s = 0;
            for (int k = 0; k < n; ++k) s += G[i * n + k] * ld_get(tx, k) * rhs[k];
            out[i] = s;
        }
        for (int i = 0; i < n; ++i) ld_put(H, i, out[i]);
    }
    free(F); free(Rt); free(T1); free(G); free(Ryy);
    return rc;
}

/* Rank-1 closed form (survey 8(c)-ii): with R = H_ls H_ls^H,
 * H = H_ls (v^H y)/(s2 + v^H v), v = x .* H_ls, y = rx.  A second, independent
 * evaluation of orc_mmse_cconv used as a known-answer identity. */
void orc_mmse_rank1(const double *tx, const double *rx, double ow2, const double *H_ls, double *H)
{
    ldc vy = 0;
    long double vv = 0;
    for (int k = 0; k < NSC; ++k) {
        ldc v = ld_get(tx, k) * ld_get(H_ls, k);
        vy += conjl(v) * ld_get(rx, k);
        vv += creall(v) * creall(v) + cimagl(v) * cimagl(v);
    }
    ldc g = vy / ((long double)ow2 + vv);
    for (int k = 0; k < NSC; ++k) ld_put(H, k, g * ld_get(H_ls, k));
}

/* ------------------------------------------------------------------ */
/* receiver front-end (the step before the estimators)  */
/* ------------------------------------------------------------------ */

/* y[i] = X[(i - 26) mod 64], i < 53, X = 64-point DFT of x (MATLAB fft: X[k] = sum_n x[n] exp(-2 pi i k n / 64)),
 * evaluated as the O(N^2) definition in long double: WiFi_blocks_extraction.m:7-9, WiFi_RX.m:22-23. */
static void dft64_shift_keep53(const ldc x[64], ldc y[NSC])
{
    for (int i = 0; i < NSC; ++i) {
        int k = (i - 26 + 64) % 64;
        ldc acc = 0;
        for (int n = 0; n < 64; ++n) {
            int m = (k * n) % 64;                       /* exact argument reduction */
            acc += x[n] * (cosl(-2.0L * M_PIl * m / 64.0L) + sinl(-2.0L * M_PIl * m / 64.0L) * I);
        }
        y[i] = acc;
    }
}

/* WiFi_blocks_extraction.m:5-10 + WiFi_RX.m:19-31.  packet [n][1200] (15 blocks of 16 CP + 64 samples), lptot [n][160]
 * (two 64-sample long-training symbols at the end) -> symb [n][15][53], pre_fft [n][53], ow2 [n] (may be NULL):
 *   symb[b]  = keep53(circshift(fft(block b without its cyclic prefix), 26))
 *   pre_fft  = keep53(circshift(fft((p1 + p2)/2), 26)),  p1 = lptot[96:160], p2 = lptot[32:96]
 *   ow2      = sum |p2 - p1|^2 / (2*64) */
void orc_frontend(const double *packet, const double *lptot, double *symb, double *pre_fft, double *ow2, long n_frames)
{
    for (long f = 0; f < n_frames; ++f) {
        ldc x[64], y[NSC];
        for (int b = 0; b < NBLK; ++b) {
            for (int n = 0; n < 64; ++n) x[n] = ld_get(packet, 1200 * f + 80 * b + 16 + n);
            dft64_shift_keep53(x, y);
            for (int k = 0; k < NSC; ++k) ld_put(symb, NSC * NBLK * f + NSC * b + k, y[k]);
        }
        long double nv = 0;
        for (int n = 0; n < 64; ++n) {
            ldc p1 = ld_get(lptot, 160 * f + 96 + n), p2 = ld_get(lptot, 160 * f + 32 + n), dlt = p2 - p1;
            x[n] = (p1 + p2) / 2.0L;
            nv += creall(dlt * conjl(dlt));
        }
        dft64_shift_keep53(x, y);
        for (int k = 0; k < NSC; ++k) ld_put(pre_fft, NSC * f + k, y[k]);
        if (ow2) ow2[f] = (double)(nv / 128.0L);
    }
}

int orc_sizeof_long_double(void) { return (int)sizeof(long double); }
