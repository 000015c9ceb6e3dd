/* wifi_dropin.c -- C99 host shim: the reference's single-frame entry points (include/wifi_dropin.h) over the
 * `_host` C-ABI of libwifi_b200.so.  Only conversion (x87 long double <-> FP64) and row-table gather/scatter
 * happen here; all arithmetic runs in the sm_100a kernels. */
#include <stdio.h>
#include <stdlib.h>

#include "wifi_b200.h"
#include "wifi_dropin.h"

typedef WIFI_LDC ldc;
typedef double _Complex dc;

static int g_intended = 0;
void wifi_dropin_set_intended(int on) { g_intended = on ? 1 : 0; }

static void die(const char *what, int rc)
{
    fprintf(stderr, "wifi_dropin: %s failed (%d): %s\n", what, rc, wifi_last_error(wifi_default_ctx()));
    abort();
}

static dc *vec_in(const ldc *v, int n)
{
    dc *o = (dc *)malloc(sizeof(dc) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) o[i] = (dc)v[i];
    return o;
}
static void vec_out(const dc *v, ldc *o, int n) { for (int i = 0; i < n; ++i) o[i] = (ldc)v[i]; }
static dc *mat_in(ldc **m, int r, int c)
{
    dc *o = (dc *)malloc(sizeof(dc) * (size_t)(r * c > 0 ? r * c : 1));
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) o[i * c + j] = (dc)m[i][j];
    return o;
}
static void mat_out(const dc *v, ldc **m, int r, int c)
{
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m[i][j] = (ldc)v[i * c + j];
}

void wifi_dropin_LT_LS(ldc tx_pre[], ldc rx_pre[], ldc H_EST[])
{
    dc *a = vec_in(tx_pre, WIFI_NSC), *b = vec_in(rx_pre, WIFI_NSC), h[WIFI_NSC];
    int rc = wifi_lt_ls_host(wifi_default_ctx(), WIFI_F64, a, b, h, 1);
    if (rc) die("LT_LS", rc);
    vec_out(h, H_EST, WIFI_NSC);
    free(a); free(b);
}

static void ps_one(int which, ldc tx[], ldc rx[], ldc H_EST[])
{
    dc *a = vec_in(tx, WIFI_NSC), *b = vec_in(rx, WIFI_NSC), h[WIFI_NSC];
    int rc = wifi_ps_host(wifi_default_ctx(), WIFI_F64, which, a, b, WIFI_NSC, h, h, h, 1);
    if (rc) die("PS", rc);
    vec_out(h, H_EST, WIFI_NSC);
    free(a); free(b);
}
void wifi_dropin_PS_Linear(ldc tx[], ldc rx[], ldc H[]) { ps_one(WIFI_PS_LINEAR, tx, rx, H); }
void wifi_dropin_PS_Cubic(ldc tx[], ldc rx[], ldc H[]) { ps_one(WIFI_PS_CUBIC, tx, rx, H); }
void wifi_dropin_PS_Sinc(ldc tx[], ldc rx[], ldc H[]) { ps_one(WIFI_PS_SINC, tx, rx, H); }

void wifi_dropin_PS_MMSE(ldc tx[], ldc rx[], ldc **F, double ow2, ldc H_EST_LS[], ldc H_EST[])
{
    (void)F;   /* R = F (F^-1 H_ls)(F^-1 H_ls)^H F^H = H_ls H_ls^H for the 53-point DFT of main.c:22-26 */
    dc *a = vec_in(tx, WIFI_NSC), *b = vec_in(rx, WIFI_NSC), *l = vec_in(H_EST_LS, WIFI_NSC), h[WIFI_NSC];
    int rc = wifi_mmse_cconv_host(wifi_default_ctx(), WIFI_F64, a, b, &ow2, l, h, 1);
    if (rc) die("PS_MMSE", rc);
    vec_out(h, H_EST, WIFI_NSC);
    free(a); free(b); free(l);
}

void wifi_dropin_hermitian(ldc **M, int row, int col, ldc **res)
{
    dc *a = mat_in(M, row, col), *o = (dc *)malloc(sizeof(dc) * (size_t)(row * col > 0 ? row * col : 1));
    int rc = wifi_chermitian_host(wifi_default_ctx(), WIFI_F64, g_intended ? WIFI_INTENDED : WIFI_AS_WRITTEN, a, row, col, o, 1);
    if (rc) die("hermitian", rc);
    mat_out(o, res, col, row);
    free(a); free(o);
}

void wifi_dropin_multiply(ldc **M1, int row1, int col1, ldc **M2, int row2, int col2, ldc **res)
{
    if (col1 != row2) { printf("Matrices dimension missmatch\n"); return; }     /* utils.c:18-19 */
    dc *a = mat_in(M1, row1, col1), *b = mat_in(M2, row2, col2), *o = (dc *)malloc(sizeof(dc) * (size_t)(row1 * col2 > 0 ? row1 * col2 : 1));
    int rc = wifi_cmatmul_host(wifi_default_ctx(), WIFI_F64, a, row1, col1, b, row2, col2, o, 1);
    if (rc) die("multiply", rc);
    mat_out(o, res, row1, col2);
    free(a); free(b); free(o);
}

void wifi_dropin_multiplyVxVeqM(ldc **M1, int row1, int col1, ldc **M2, int row2, int col2, ldc **res)
{
    if (col1 != row2) { printf("Matrices dimension missmatch\n"); return; }     /* utils.c:56-57 */
    dc *a = mat_in(M1, row1, col1), *b = mat_in(M2, row2, col2), *o = (dc *)malloc(sizeof(dc) * (size_t)(row1 * col2 > 0 ? row1 * col2 : 1));
    int rc = wifi_couter_host(wifi_default_ctx(), WIFI_F64, a, row1, col1, b, row2, col2, o, 1);
    if (rc) die("multiplyVxVeqM", rc);
    mat_out(o, res, row1, col2);
    free(a); free(b); free(o);
}

void wifi_dropin_identity(ldc **Identity, int size, double scalar)
{
    dc *o = (dc *)malloc(sizeof(dc) * (size_t)(size * size > 0 ? size * size : 1));
    int rc = wifi_cidentity_host(wifi_default_ctx(), WIFI_F64, o, size, scalar, 1);
    if (rc) die("identity", rc);
    mat_out(o, Identity, size, size);
    free(o);
}

void wifi_dropin_addition(ldc **M1, int row1, int col1, ldc **M2, int row2, int col2, ldc **res)
{
    if (row1 != row2 || col1 != col2) { printf("Matrices dimension missmatch\n"); return; }   /* utils.c:112-113 */
    dc *a = mat_in(M1, row1, col1), *b = mat_in(M2, row2, col2), *o = (dc *)malloc(sizeof(dc) * (size_t)(row1 * col1 > 0 ? row1 * col1 : 1));
    int rc = wifi_cadd_host(wifi_default_ctx(), WIFI_F64, g_intended ? WIFI_INTENDED : WIFI_AS_WRITTEN, a, row1, col1, b, row2, col2, o, 1);
    if (rc) die("addition", rc);
    mat_out(o, res, row1, col1);
    free(a); free(b); free(o);
}

void wifi_dropin_inverse(ldc **A, int order, ldc **Y)
{
    dc *a = mat_in(A, order, order), *o = (dc *)malloc(sizeof(dc) * (size_t)order * order);
    int info = 0;
    int rc = wifi_cinverse_host(wifi_default_ctx(), WIFI_F64, a, order, o, 1, &info);
    if (rc && rc != WIFI_ERR_SINGULAR) die("inverse", rc);   /* singular: NaN/Inf out, like the reference's 0/0 */
    mat_out(o, Y, order, order);
    free(a); free(o);
}

/* the reference's names */
void WiFi_channel_estimation_LT_LS(ldc a[], ldc b[], ldc h[]) { wifi_dropin_LT_LS(a, b, h); }
void WiFi_channel_estimation_PS_Linear(ldc a[], ldc b[], ldc h[]) { wifi_dropin_PS_Linear(a, b, h); }
void WiFi_channel_estimation_PS_Cubic(ldc a[], ldc b[], ldc h[]) { wifi_dropin_PS_Cubic(a, b, h); }
void WiFi_channel_estimation_PS_Sinc(ldc a[], ldc b[], ldc h[]) { wifi_dropin_PS_Sinc(a, b, h); }
void WiFi_channel_estimation_PS_MMSE(ldc a[], ldc b[], ldc **F, double ow2, ldc l[], ldc h[]) { wifi_dropin_PS_MMSE(a, b, F, ow2, l, h); }
void hermitian(ldc **M, int row, int col, ldc **res) { wifi_dropin_hermitian(M, row, col, res); }
void multiply(ldc **M1, int r1, int c1, ldc **M2, int r2, int c2, ldc **res) { wifi_dropin_multiply(M1, r1, c1, M2, r2, c2, res); }
void multiplyVxVeqM(ldc **M1, int r1, int c1, ldc **M2, int r2, int c2, ldc **res) { wifi_dropin_multiplyVxVeqM(M1, r1, c1, M2, r2, c2, res); }
void identity(ldc **Id, int size, double scalar) { wifi_dropin_identity(Id, size, scalar); }
void addition(ldc **M1, int r1, int c1, ldc **M2, int r2, int c2, ldc **res) { wifi_dropin_addition(M1, r1, c1, M2, r2, c2, res); }
void inverse(ldc **A, int order, ldc **Y) { wifi_dropin_inverse(A, order, Y); }
