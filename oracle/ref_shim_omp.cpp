/*
 * ref_shim_omp.cpp -- builds oracle/_ref/libwifi_ref_omp.so: the reference's OWN OpenMP
 * build (/root/reference/main_openmp.c + the _omp twins of utils.c), compiled where it
 * lies with -fopenmp, exactly as it is: every estimator call spawns its intra-frame team
 * (num_threads(53), one thread per sub-carrier: main_openmp.c:72-77, 92-109, 132-135,
 * 150-173).  Test infrastructure only: bench.py times it on a small N' next to the
 * sequential build and the frame-parallel harness (SURVEY 8(d) CPU baseline (ii)).
 * PS_MMSE of this build is not timed: inverse_omp is racy and the estimator returns NaN
 * after minutes per frame.
 */
#define main ref_main_openmp
#include "main_openmp.c"      /* estimators main_openmp.c:70-276 (+ inputs.h globals) */
#undef main

typedef long double complex ldc_t;

extern "C" {

/* which: 0 LT_LS, 1 PS_Linear, 2 PS_Cubic, 3 PS_Sinc.  a, b, H: [n][53] interleaved doubles.  Frames are taken one
 * after the other, as the reference's driver would: the parallelism is the reference's own, inside each call. */
void refomp_estimate(int which, const double *a, const double *b, double *H, long n_frames)
{
    ldc_t x[SAMPUTIL], y[SAMPUTIL], h[SAMPUTIL];
    for (long f = 0; f < n_frames; ++f) {
        for (int k = 0; k < SAMPUTIL; ++k) {
            x[k] = (long double)a[2 * (SAMPUTIL * f + k)] + (long double)a[2 * (SAMPUTIL * f + k) + 1] * I;
            y[k] = (long double)b[2 * (SAMPUTIL * f + k)] + (long double)b[2 * (SAMPUTIL * f + k) + 1] * I;
        }
        switch (which) {
        case 0: WiFi_channel_estimation_LT_LS(x, y, h); break;
        case 1: WiFi_channel_estimation_PS_Linear(x, y, h); break;
        case 2: WiFi_channel_estimation_PS_Cubic(x, y, h); break;
        default: WiFi_channel_estimation_PS_Sinc(x, y, h); break;
        }
        for (int k = 0; k < SAMPUTIL; ++k) { H[2 * (SAMPUTIL * f + k)] = (double)creall(h[k]); H[2 * (SAMPUTIL * f + k) + 1] = (double)cimagl(h[k]); }
    }
}

/* the build's thread settings: the reference's job script exports OMP_NUM_THREADS=60, OMP_NESTED=TRUE (main_openmp.bash:10-12);
 * the num_threads(53) clauses of the estimators override the team size anyway */
void refomp_set_threads(int n) { omp_set_num_threads(n); omp_set_dynamic(0); }

} /* extern "C" */
