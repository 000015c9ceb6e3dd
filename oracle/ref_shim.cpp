/*
 * ref_shim.cpp -- builds oracle/_ref/libwifi_ref.so: the reference's OWN sequential
 * C(++) code, compiled where it lies (/root/reference/main.c, utils.c; nothing is
 * copied), behind a tiny extern "C" double-interleaved interface so ctypes can call
 * it.  Test infrastructure only: used to pin oracle/wifi_oracle.c, to produce
 * tests/golden/, and as bench.py's CPU reference arm.  Built by oracle/Makefile with
 *   g++ -std=gnu++98 -w -Ioracle/stub -I$(REF)
 * (gnu++98 is required: newer standards drop the C99 `complex`/`I` macros the
 * reference relies on, utils.h:24).
 */
#define main ref_main
#include "main.c"      /* estimators main.c:66-212 (+ inputs.h globals, utils.h prototypes) */
#undef main

typedef long double complex ldc_t;

static void to_ld(const double *a, ldc_t *o, long n) { for (long i = 0; i < n; ++i) o[i] = (long double)a[2*i] + (long double)a[2*i+1] * I; }
static void from_ld(const ldc_t *a, double *o, long n) { for (long i = 0; i < n; ++i) { o[2*i] = (double)creall(a[i]); o[2*i+1] = (double)cimagl(a[i]); } }

struct Mat {            /* reference-style row-pointer table over one block (utils.c:817-835) */
    ldc_t **rows; int n, m;
    Mat(int n_, int m_) : n(n_), m(m_) { malloc2dLongDoubleComplex(&rows, n, m); }
    ~Mat() { free2dLongDoubleComplex(&rows); }
    void load(const double *a) { to_ld(a, rows[0], (long)n * m); }
    void store(double *a) { from_ld(rows[0], a, (long)n * m); }
};

extern "C" {

/* which: 0 LT_LS, 1 PS_Linear, 2 PS_Cubic, 3 PS_Sinc.  a, b: [n][53]; H: [n][53] */
void ref_estimate(int which, const double *a, const double *b, double *H, long n_frames)
{
    ldc_t x[SAMPUTIL], y[SAMPUTIL], h[SAMPUTIL];
    for (long f = 0; f < n_frames; ++f) {
        to_ld(a + 2 * SAMPUTIL * f, x, SAMPUTIL);
        to_ld(b + 2 * SAMPUTIL * f, y, SAMPUTIL);
        switch (which) {
        case 0: WiFi_channel_estimation_LT_LS(x, y, h); break;
        case 1: WiFi_channel_estimation_PS_Linear(x, y, h); break;
        case 2: WiFi_channel_estimation_PS_Cubic(x, y, h); break;
        default: WiFi_channel_estimation_PS_Sinc(x, y, h); break;
        }
        from_ld(h, H + 2 * SAMPUTIL * f, SAMPUTIL);
    }
}

/* PS_MMSE exactly as written (main.c:148-212): minutes per frame, NaN output. */
void ref_mmse_as_written(const double *tx, const double *rx, double ow2, const double *H_ls, double *H)
{
    ldc_t x[SAMPUTIL], y[SAMPUTIL], hls[SAMPUTIL], h[SAMPUTIL];
    to_ld(tx, x, SAMPUTIL); to_ld(rx, y, SAMPUTIL); to_ld(H_ls, hls, SAMPUTIL);
    Mat F(SAMPUTIL, SAMPUTIL);
    for (int f = 0; f < SAMPUTIL; f++) for (int t = 0; t < SAMPUTIL; t++) F.rows[t][f] = cexp(-2*I*PI*t*f/SAMPUTIL);
    WiFi_channel_estimation_PS_MMSE(x, y, F.rows, ow2, hls, h);
    from_ld(h, H, SAMPUTIL);
}

void ref_multiply(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    Mat a(r1, c1), b(r2, c2), c(r1, c2);
    a.load(M1); b.load(M2); c.load(res);      /* copy-in/copy-out: a call that writes nothing leaves res as it was */
    multiply(a.rows, r1, c1, b.rows, r2, c2, c.rows);
    c.store(res);
}
void ref_hermitian(const double *M, int row, int col, double *res)
{
    Mat a(row, col), c(col, row);
    a.load(M); hermitian(a.rows, row, col, c.rows); c.store(res);
}
void ref_outer(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    Mat a(r1, c1), b(r2, c2), c(r1, c2);
    a.load(M1); b.load(M2);
    multiplyVxVeqM(a.rows, r1, c1, b.rows, r2, c2, c.rows);
    c.store(res);
}
void ref_identity(double *Id, int size, double scalar)
{
    Mat a(size, size); identity(a.rows, size, scalar); a.store(Id);
}
void ref_addition(const double *M1, int r1, int c1, const double *M2, int r2, int c2, double *res)
{
    Mat a(r1, c1), b(r2, c2), c(r1, c1);
    a.load(M1); b.load(M2);
    addition(a.rows, r1, c1, b.rows, r2, c2, c.rows);
    c.store(res);
}
void ref_inverse(const double *A, int order, double *Y)
{
    Mat a(order, order), y(order, order);
    a.load(A); inverse(a.rows, order, y.rows); y.store(Y);
}
double ref_sinc(double x) { return sinc(x); }

/* the inputs.h globals (inputs.h:18-1724), for the golden-vector script */
double ref_ow2(void) { return OW2; }
void ref_inputs(double *tx_pre, double *rx_pre, double *txs, double *rxs)
{
    from_ld(tx_preamble_fft, tx_pre, SAMPUTIL); from_ld(rx_preamble_fft, rx_pre, SAMPUTIL);
    from_ld(tx_symb, txs, SIZESYMBOL); from_ld(rx_symb, rxs, SIZESYMBOL);
}

/* OpenMP team size of the timing loops below (torchrun exports OMP_NUM_THREADS=1; the bench overrides it here) */
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_get_max_threads(void) { return omp_get_max_threads(); }

/* ---- timing loops for bench.py's CPU arm: the reference functions, frame-parallel
 * over all host threads (the fair multi-core number, survey 8(d)-iii) ---- */
void ref_estimate_omp(int which, const double *a, const double *b, double *H, long n_frames)
{
    #pragma omp parallel for schedule(static)
    for (long f = 0; f < n_frames; ++f) ref_estimate(which, a + 2 * SAMPUTIL * f, b + 2 * SAMPUTIL * f, H + 2 * SAMPUTIL * f, 1);
}

/* shared-filter MMSE on the CPU with the reference's routines: per frame the pilot-free
 * LS divide (rx/tx, main.c:83 arithmetic) and multiply(W 53x53, H_ls 53x1) utils.c:16-31 */
void ref_mmse_shared_omp(const double *W, const double *tx, const double *rx, double *H, long n_frames)
{
    #pragma omp parallel
    {
        Mat w(SAMPUTIL, SAMPUTIL), v(SAMPUTIL, 1), o(SAMPUTIL, 1);
        w.load(W);
        ldc_t x[SAMPUTIL], y[SAMPUTIL];
        #pragma omp for schedule(static)
        for (long f = 0; f < n_frames; ++f) {
            to_ld(tx + 2 * SAMPUTIL * f, x, SAMPUTIL); to_ld(rx + 2 * SAMPUTIL * f, y, SAMPUTIL);
            for (int k = 0; k < SAMPUTIL; ++k) v.rows[k][0] = y[k] / x[k];
            multiply(w.rows, SAMPUTIL, SAMPUTIL, v.rows, SAMPUTIL, 1, o.rows);
            from_ld(o.rows[0], H + 2 * SAMPUTIL * f, SAMPUTIL);
        }
    }
}

} /* extern "C" */
