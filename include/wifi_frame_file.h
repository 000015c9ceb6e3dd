/*
 * wifi_frame_file.h -- flat binary frame files.  The reference's only data formats are the C literals
 * of inputs.h and WiFi_inputs.m; this is the batched equivalent: a 64-byte header followed by whole arrays ("planes"),
 * each a dense [n_frames][...] block of interleaved (re, im) float2 / double2 values in the inputs.h layout.
 *
 *   kind WIFI_FILE_FREQ   tx_pre [n][53], rx_pre [n][53], tx_symb [n][15][53], rx_symb [n][15][53]
 *                         (inputs.h:20,75,130,928: element 53*b + k = sub-carrier k of OFDM block b)
 *   kind WIFI_FILE_TIME   tx_packet [n][1200], rx_packet [n][1200], tx_lptot [n][160], rx_lptot [n][160]
 *                         (WiFi_inputs.m; the front-end of wifi_frontend_* turns them into the FREQ arrays)
 *   kind WIFI_FILE_EST    H_lt_ls, H_ps_linear, H_ps_cubic, H_ps_sinc, H_ps_mmse [n][53] each, eq [n][15][53], ow2 [n] real
 *                         (what host/wifi_host_main --file writes)
 */
#ifndef WIFI_FRAME_FILE_H
#define WIFI_FRAME_FILE_H
#include <stdint.h>

#define WIFI_FILE_MAGIC "WIFIFRM1"
enum { WIFI_FILE_FREQ = 0, WIFI_FILE_TIME = 1, WIFI_FILE_EST = 2 };

typedef struct {
    char magic[8];       /* WIFI_FILE_MAGIC */
    uint32_t dtype;      /* wifi_dtype: 0 = complex64, 1 = complex128 */
    uint32_t kind;       /* WIFI_FILE_* */
    uint64_t n_frames;
    uint8_t pad[40];     /* zero; planes start at byte 64 */
} wifi_file_header;

#endif
