/* oracle/shim/fftw3.h — FFTW3 (fftw3f, un-pinned, configure.ac:291) is not
 * installed in this image.  The reference's only use is
 * fftwf_plan_many_dft/execute (arch/common/fft.c:80-83,112): `howmany`
 * independent rank-1 length-n DFTs with arbitrary strides.  This stand-in is
 * the mathematical DFT accumulated in double (O(n^2)); parity at that boundary
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
	int n, howmany, istride, idist, ostride, odist, sign;
	fftwf_complex *in, *out;
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
	return p;
}

static inline void fftwf_execute(const fftwf_plan p)
{
	const int n = p->n;
	double *tr = (double *)malloc(sizeof(double) * 2 * n);
	for (int t = 0; t < p->howmany; t++) {
		const fftwf_complex *x = p->in + (size_t)t * p->idist;
		fftwf_complex *y = p->out + (size_t)t * p->odist;
		for (int k = 0; k < n; k++) {
			double ar = 0.0, ai = 0.0;
			for (int j = 0; j < n; j++) {
				double ph = p->sign * 2.0 * M_PI * (double)(((long)j * k) % n) / (double)n;
				double c = cos(ph), s = sin(ph);
				double xr = x[(size_t)j * p->istride][0], xi = x[(size_t)j * p->istride][1];
				ar += xr * c - xi * s;
				ai += xr * s + xi * c;
			}
			tr[2 * k] = ar; tr[2 * k + 1] = ai;
		}
		for (int k = 0; k < n; k++) {
			y[(size_t)k * p->ostride][0] = (float)tr[2 * k];
			y[(size_t)k * p->ostride][1] = (float)tr[2 * k + 1];
		}
	}
	free(tr);
}

static inline void fftwf_destroy_plan(fftwf_plan p) { free(p); }
static inline void *fftwf_malloc(size_t n) { void *p = NULL; return posix_memalign(&p, 64, n) ? NULL : p; }
static inline void fftwf_free(void *p) { free(p); }
#ifdef __cplusplus
}
#endif
#endif
