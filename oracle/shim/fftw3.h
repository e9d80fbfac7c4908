/* oracle/shim/fftw3.h — FFTW3 (fftw3f, un-pinned, configure.ac:291) is not
 * installed in this image.  The reference's only use is
 * fftwf_plan_many_dft/execute (arch/common/fft.c:80-83,112): `howmany`
 * independent rank-1 length-n DFTs with arbitrary strides.  This stand-in is
 * the mathematical DFT evaluated in double and rounded to float once: a radix-2
 * FFT for power-of-two n (so that the reference arm of bench.py is not timed on
 * an O(n^2) transform FFTW would never run), the plain O(n^2) sum otherwise,
 * both over an exact twiddle table built at plan time.  Parity at that boundary
 * is therefore "unpinned" (tolerance 1e-4 relative).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stdlib.h>
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
struct oracle_fftwf_plan_s {
	int n, howmany, istride, idist, ostride, odist, sign, log2n;
	fftwf_complex *in, *out;
	double *tw;   /* [n][2]: e^{sign * 2 pi i m / n} */
	double *work; /* [n][2] */
	int *rev;     /* bit-reversal permutation (power-of-two n) */
};
typedef struct oracle_fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

static inline fftwf_plan fftwf_plan_many_dft(int rank, const int *n, int howmany,
					     fftwf_complex *in, const int *inembed, int istride, int idist,
					     fftwf_complex *out, const int *onembed, int ostride, int odist,
					     int sign, unsigned flags)
{
	(void)inembed; (void)onembed; (void)flags;
	if (rank != 1)
		return NULL;
	fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
	p->n = n[0]; p->howmany = howmany;
	p->istride = istride; p->idist = idist;
	p->ostride = ostride; p->odist = odist;
	p->sign = sign; p->in = in; p->out = out;
	const int N = p->n;
	p->tw = (double *)malloc(sizeof(double) * 2 * N);
	p->work = (double *)malloc(sizeof(double) * 2 * N);
	p->rev = (int *)malloc(sizeof(int) * N);
	for (int m = 0; m < N; m++) {
		const double ph = sign * 2.0 * M_PI * (double)m / (double)N;
		p->tw[2 * m] = cos(ph); p->tw[2 * m + 1] = sin(ph);
	}
	p->log2n = -1;
	if (N >= 2 && (N & (N - 1)) == 0) {
		int l = 0;
		while ((1 << l) < N) l++;
		p->log2n = l;
		for (int i = 0; i < N; i++) {
			int r = 0;
			for (int b = 0; b < l; b++) if (i & (1 << b)) r |= 1 << (l - 1 - b);
			p->rev[i] = r;
		}
	}
	return p;
}

static inline void fftwf_execute(const fftwf_plan p)
{
	const int n = p->n;
	double *tr = p->work;
	const double *tw = p->tw;
	for (int t = 0; t < p->howmany; t++) {
		const fftwf_complex *x = p->in + (size_t)t * p->idist;
		fftwf_complex *y = p->out + (size_t)t * p->odist;
		if (p->log2n > 0) {
			/* iterative radix-2 decimation in time, double precision */
			for (int j = 0; j < n; j++) {
				const int r = p->rev[j];
				tr[2 * r] = x[(size_t)j * p->istride][0]; tr[2 * r + 1] = x[(size_t)j * p->istride][1];
			}
			for (int len = 2; len <= n; len <<= 1) {
				const int half = len >> 1, step = n / len;
				for (int i = 0; i < n; i += len) {
					for (int k = 0; k < half; k++) {
						const double wr = tw[2 * k * step], wi = tw[2 * k * step + 1];
						double *a = tr + 2 * (i + k), *b = tr + 2 * (i + k + half);
						const double br = b[0] * wr - b[1] * wi, bi = b[0] * wi + b[1] * wr;
						b[0] = a[0] - br; b[1] = a[1] - bi;
						a[0] += br; a[1] += bi;
					}
				}
			}
		} else {
			for (int k = 0; k < n; k++) {
				double ar = 0.0, ai = 0.0;
				for (int j = 0; j < n; j++) {
					const int m = (int)(((long)j * k) % n);
					const double c = tw[2 * m], s = tw[2 * m + 1];
					const double xr = x[(size_t)j * p->istride][0], xi = x[(size_t)j * p->istride][1];
					ar += xr * c - xi * s;
					ai += xr * s + xi * c;
				}
				tr[2 * k] = ar; tr[2 * k + 1] = ai;
			}
		}
		for (int k = 0; k < n; k++) {
			y[(size_t)k * p->ostride][0] = (float)tr[2 * k];
			y[(size_t)k * p->ostride][1] = (float)tr[2 * k + 1];
		}
	}
}

static inline void fftwf_destroy_plan(fftwf_plan p) { if (p) { free(p->tw); free(p->work); free(p->rev); } free(p); }
static inline void *fftwf_malloc(size_t n) { void *p = NULL; return posix_memalign(&p, 64, n) ? NULL : p; }
static inline void fftwf_free(void *p) { free(p); }
#ifdef __cplusplus
}
#endif
#endif
