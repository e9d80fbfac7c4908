/*
 * oracle/oracle_convolve.c — scalar-C restatement of the reference FIR kernels.
 * TEST INFRASTRUCTURE ONLY (see oracle_trx.h).
 *
 * y[i] = sum_k x[i + start - (h_len-1) + k] * h[k],  i < len      (arch/common/convolve_base.c:57-82)
 *
 * The SSE3 entry points (arch/x86/convolve.c:94-152) use a fixed float32 summation
 * tree per tap count; it is restated here lane by lane so results are bit-identical
 * to the SSE build.  Lane j holds taps k == j (mod 4); the two _mm_hadd_ps at the end
 * of every kernel give (L0+L1)+(L2+L3)  (the pair-swap done by _MM_SHUFFLE(0,2,0,2)
 * is irrelevant because float addition is commutative).
 */
#include <string.h>
#include <stdio.h>
#include "oracle_trx.h"

/* arch/common/convolve_base.c:114-131 */
static int bounds_check(int x_len, int h_len, int y_len, int start, int len)
{
	if (x_len < 1 || h_len < 1 || y_len < 1 || len < 1)
		return -1;
	if (start + len > x_len || len > y_len || x_len < h_len)
		return -1;
	return 0;
}

/* arch/common/convolve_base.c:27-82: strictly sequential MAC, y pre-zeroed */
static void base_real(const float *x, const float *h, int h_len, float *y, int start, int len)
{
	for (int i = 0; i < len; i++) {
		const float *xx = &x[2 * (i - (h_len - 1) + start)];
		float yr = 0.0f, yi = 0.0f;
		for (int k = 0; k < h_len; k++) {
			yr += xx[2 * k] * h[2 * k];
			yi += xx[2 * k + 1] * h[2 * k];
		}
		y[2 * i] = yr;
		y[2 * i + 1] = yi;
	}
}

static void base_complex(const float *x, const float *h, int h_len, float *y, int start, int len)
{
	for (int i = 0; i < len; i++) {
		const float *xx = &x[2 * (i - (h_len - 1) + start)];
		float yr = 0.0f, yi = 0.0f;
		for (int k = 0; k < h_len; k++) {
			yr += xx[2 * k] * h[2 * k] - xx[2 * k + 1] * h[2 * k + 1];
			yi += xx[2 * k] * h[2 * k + 1] + xx[2 * k + 1] * h[2 * k];
		}
		y[2 * i] = yr;
		y[2 * i + 1] = yi;
	}
}

/* one output of the fixed-size real kernels; c selects re (0) / im (1) of x */
static inline float real_tree(const float *xx, const float *h, int h_len, int c)
{
	float L[4];
	for (int j = 0; j < 4; j++) {
#define P(k) (xx[2 * (k) + c] * h[2 * (k)])
		switch (h_len) {
		case 4: /* convolve_sse_3.c:30-68 */
			L[j] = P(j);
			break;
		case 8: /* :71-119 */
			L[j] = P(j) + P(4 + j);
			break;
		case 12: /* :122-185 */
			L[j] = (P(j) + P(4 + j)) + P(8 + j);
			break;
		case 16: /* :188-264 */
			L[j] = (P(j) + P(4 + j)) + (P(8 + j) + P(12 + j));
			break;
		case 20: /* :267-354 */
			L[j] = ((P(j) + P(4 + j)) + P(8 + j)) + (P(12 + j) + P(16 + j));
			break;
		default: { /* sse_conv_real4n :357-401: sequential over groups from 0 */
			float a = 0.0f;
			for (int n = 0; n < h_len / 4; n++)
				a = a + P(4 * n + j);
			L[j] = a;
		}
		}
#undef P
	}
	return (L[0] + L[1]) + (L[2] + L[3]);
}

/* arch/x86/convolve.c:94-132 */
int orc_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	(void)x_len; (void)y_len; /* bounds only checked in non-optimised reference builds (convolve.c:98-101) */
	memset(y, 0, (size_t)len * 2 * sizeof(float));
	if (h_len % 4) {
		base_real(x, h, h_len, y, start, len);
		return len;
	}
	for (int i = 0; i < len; i++) {
		const float *xx = &x[2 * (i - (h_len - 1) + start)];
		y[2 * i] = real_tree(xx, h, h_len, 0);
		y[2 * i + 1] = real_tree(xx, h, h_len, 1);
	}
	return len;
}

/* arch/x86/convolve.c:135-152 */
int orc_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	(void)x_len; (void)y_len;
	memset(y, 0, (size_t)len * 2 * sizeof(float));
	if (h_len % 4) {
		base_complex(x, h, h_len, y, start, len);
		return len;
	}
	for (int i = 0; i < len; i++) {
		const float *xx = &x[2 * (i - (h_len - 1) + start)];
		float Lr[4], Li[4];
		if (!(h_len % 8)) {
			/* sse_conv_cmplx_8n convolve_sse_3.c:462-537: accumulators A (taps 8n+0..3) and
			 * B (taps 8n+4..7), each sequential over n from zero, then A+B lane-wise */
			for (int j = 0; j < 4; j++) {
				float ar = 0.0f, br = 0.0f, ai = 0.0f, bi = 0.0f;
				for (int n = 0; n < h_len / 8; n++) {
					int ka = 8 * n + j, kb = 8 * n + 4 + j;
					float xr = xx[2 * ka], xi = xx[2 * ka + 1], hr = h[2 * ka], hi = h[2 * ka + 1];
					ar = ar + (hr * xr - hi * xi);
					ai = ai + (hr * xi + hi * xr);
					xr = xx[2 * kb]; xi = xx[2 * kb + 1]; hr = h[2 * kb]; hi = h[2 * kb + 1];
					br = br + (hr * xr - hi * xi);
					bi = bi + (hr * xi + hi * xr);
				}
				Lr[j] = ar + br;
				Li[j] = ai + bi;
			}
		} else {
			/* sse_conv_cmplx_4n :404-459 */
			for (int j = 0; j < 4; j++) {
				float ar = 0.0f, ai = 0.0f;
				for (int n = 0; n < h_len / 4; n++) {
					int k = 4 * n + j;
					float xr = xx[2 * k], xi = xx[2 * k + 1], hr = h[2 * k], hi = h[2 * k + 1];
					ar = ar + (hr * xr - hi * xi);
					ai = ai + (hr * xi + hi * xr);
				}
				Lr[j] = ar;
				Li[j] = ai;
			}
		}
		y[2 * i] = (Lr[0] + Lr[1]) + (Lr[2] + Lr[3]);
		y[2 * i + 1] = (Li[0] + Li[1]) + (Li[2] + Li[3]);
	}
	return len;
}

/* arch/common/convolve_base.c:133-165 */
int orc_base_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	if (bounds_check(x_len, h_len, y_len, start, len) < 0)
		return -1;
	memset(y, 0, (size_t)len * 2 * sizeof(float));
	base_real(x, h, h_len, y, start, len);
	return len;
}

int orc_base_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start,
			      int len)
{
	if (bounds_check(x_len, h_len, y_len, start, len) < 0)
		return -1;
	memset(y, 0, (size_t)len * 2 * sizeof(float));
	base_complex(x, h, h_len, y, start, len);
	return len;
}
