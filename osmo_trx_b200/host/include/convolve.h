/* convolve.h — the C FIR entry points of arch/common/convolve.h:4-26, served by the GPU library.
 * y[i] = sum_k x[i + start - (h_len-1) + k] * h[k], i < len; x, h, y interleaved complex float; real taps
 * are stored as complex with zero imaginary part, exactly as the reference's callers do.  x must have the
 * head-room the reference requires (samples before x[0] are read when start < h_len - 1).  Return value:
 * len, or -1 on a bounds_check failure (convolve_base.c:114-131). */
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void *convolve_h_alloc(size_t num);
int convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
int convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
int base_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
int base_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
void convolve_init(void);
#ifdef __cplusplus
}
#endif
