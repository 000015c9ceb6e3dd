/*
 * wifi_dropin.h -- single-frame drop-ins with the reference's own entry points and `long double complex` array
 * layouts, so one frame is a drop-in (libwifi_dropin.so, C99; libwifi_dropin_cxx.so carries the same functions
 * with the C++ linkage the reference's g++-built objects expect, compile.c:25-30).
 *
 *   main.c:4-8     WiFi_channel_estimation_{LT_LS,PS_Linear,PS_Cubic,PS_Sinc,PS_MMSE}
 *   utils.h:38-60  hermitian, multiply, multiplyVxVeqM, identity, addition, inverse
 *
 * Each call converts x87 long double -> FP64, runs the sm_100a kernels of libwifi_b200.so through its `_host`
 * C-ABI on the process-global context (wifi_default_ctx(): device $WIFI_B200_DEVICE, aborts without a GPU), and
 * converts back.  Semantics:
 *   - LT_LS, PS_Linear, PS_Cubic, PS_Sinc, multiply, multiplyVxVeqM, identity: the reference's, to FP64 rounding
 *     (LT_LS keeps the NaN-when-Re(tx)==Im(tx) behaviour of main.c:69-72).
 *   - hermitian / addition: AS WRITTEN by default (utils.c:6 Re-Im, utils.c:117 M1+M1) so results are identical to
 *     the reference's; wifi_dropin_set_intended(1) switches both to the conjugate transpose / M1+M2.
 *   - inverse: partial-pivoting Gauss-Jordan instead of the O(n^5) un-pivoted cofactor expansion (utils.c:141-170);
 *     same result wherever the reference's is finite, to its own accuracy.
 *   - PS_MMSE: the INTENDED formula H = R (R + ow2 (X X^H)^-1)^-1 (rx/tx), R = H_ls H_ls^H
 *     (WiFi_channel_estimation_PS_MMSE.m:16-33); the reference's C body returns NaN for every input.
 * A dimension mismatch prints "Matrices dimension missmatch" and writes nothing, like utils.c:18-19.
 */
#ifndef WIFI_DROPIN_H
#define WIFI_DROPIN_H
#include <complex.h>

#ifdef __cplusplus
#define WIFI_LDC long double _Complex
extern "C" {
#else
#define WIFI_LDC long double _Complex
#endif

void wifi_dropin_set_intended(int on);

void wifi_dropin_LT_LS(WIFI_LDC tx_pre[], WIFI_LDC rx_pre[], WIFI_LDC H_EST[]);
void wifi_dropin_PS_Linear(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC H_EST[]);
void wifi_dropin_PS_Cubic(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC H_EST[]);
void wifi_dropin_PS_Sinc(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC H_EST[]);
void wifi_dropin_PS_MMSE(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC **F, double ow2, WIFI_LDC H_EST_LS[], WIFI_LDC H_EST[]);
void wifi_dropin_hermitian(WIFI_LDC **M, int row, int col, WIFI_LDC **res);
void wifi_dropin_multiply(WIFI_LDC **M1, int row1, int col1, WIFI_LDC **M2, int row2, int col2, WIFI_LDC **res);
void wifi_dropin_multiplyVxVeqM(WIFI_LDC **M1, int row1, int col1, WIFI_LDC **M2, int row2, int col2, WIFI_LDC **res);
void wifi_dropin_identity(WIFI_LDC **Identity, int size, double scalar);
void wifi_dropin_addition(WIFI_LDC **M1, int row1, int col1, WIFI_LDC **M2, int row2, int col2, WIFI_LDC **res);
void wifi_dropin_inverse(WIFI_LDC **A, int order, WIFI_LDC **Y);

#ifndef WIFI_DROPIN_NO_REFERENCE_NAMES
/* the reference's names, C linkage (main.c:4-8, utils.h:38-60) */
void WiFi_channel_estimation_LT_LS(WIFI_LDC tx_pre[], WIFI_LDC rx_pre[], WIFI_LDC H_EST[]);
void WiFi_channel_estimation_PS_Linear(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC H_EST[]);
void WiFi_channel_estimation_PS_Cubic(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC H_EST[]);
void WiFi_channel_estimation_PS_Sinc(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC H_EST[]);
void WiFi_channel_estimation_PS_MMSE(WIFI_LDC tx_symbols[], WIFI_LDC rx_symbols[], WIFI_LDC **F, double ow2, WIFI_LDC H_EST_LS[], WIFI_LDC H_EST[]);
void hermitian(WIFI_LDC **M, int row, int col, WIFI_LDC **res);
void multiply(WIFI_LDC **M1, int row1, int col1, WIFI_LDC **M2, int row2, int col2, WIFI_LDC **res);
void multiplyVxVeqM(WIFI_LDC **M1, int row1, int col1, WIFI_LDC **M2, int row2, int col2, WIFI_LDC **res);
void identity(WIFI_LDC **Identity, int size, double scalar);
void addition(WIFI_LDC **M1, int row1, int col1, WIFI_LDC **M2, int row2, int col2, WIFI_LDC **res);
void inverse(WIFI_LDC **A, int order, WIFI_LDC **Y);
#endif

#ifdef __cplusplus
}
#endif
#endif
