/*
 * oracle/oracle_filterbank.c — plain-C restatement of Resampler.cpp, ChannelizerBase.cpp,
 * Channelizer.cpp and Synthesis.cpp.  TEST INFRASTRUCTURE ONLY (see oracle_trx.h).
 *
 * The M-point DFT across branches is FFTW3 in the reference (arch/common/fft.c:55-114, un-pinned
 * third-party dependency absent from /root/reference): parity at that boundary is UNPINNED; the
 * oracle uses the mathematical forward DFT accumulated in double (tolerance 1e-4 relative).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_trx.h"

void orc_resampler_proto(int p, int q, int filt_len, float bw, float *parts);

struct orc_resampler {
	int p, q, filt_len;
	float *parts; /* [p][filt_len] real taps, reversed */
};

/* Resampler::Resampler/init (Resampler.cpp:152-181) */
orc_resampler *orc_resampler_create(int p, int q, int filt_len, float bw)
{
	if (p <= 0 || q <= 0 || filt_len <= 0)
		return NULL;
	orc_resampler *r = (orc_resampler *)calloc(1, sizeof(*r));
	r->p = p; r->q = q; r->filt_len = filt_len;
	r->parts = (float *)malloc(sizeof(float) * p * filt_len);
	orc_resampler_proto(p, q, filt_len, bw, r->parts);
	return r;
}

void orc_resampler_destroy(orc_resampler *r)
{
	if (!r) return;
	free(r->parts);
	free(r);
}

int orc_resampler_taps(orc_resampler *r, int path, float *out)
{
	memcpy(out, r->parts + (size_t)path * r->filt_len, sizeof(float) * r->filt_len);
	return r->filt_len;
}

/* Resampler::rotate (Resampler.cpp:131-150): one convolve_real(len=1) per output,
 * in_index[i] = (q*i)/p, out_path[i] = (q*i)%p (:160-165), MAX_OUTPUT_LEN 16384 */
int orc_resampler_rotate(orc_resampler *r, const float *in_with_hist, int hist, int in_len, float *out, int out_len)
{
	if (out_len > 4096 * 4)
		return -1;
	if (in_len % r->q || out_len % r->p || in_len / r->q != out_len / r->p)
		return -1; /* check_vec_len (debug builds only in the reference, :98-129) */
	const float *in = in_with_hist + 2 * hist;
	float *h = (float *)malloc(sizeof(float) * 2 * r->filt_len);
	int cur = -1;
	for (int i = 0; i < out_len; i++) {
		int n = (int)(((long)r->q * i) / r->p), path = (int)(((long)r->q * i) % r->p);
		if (path != cur) {
			for (int k = 0; k < r->filt_len; k++) { h[2 * k] = r->parts[path * r->filt_len + k]; h[2 * k + 1] = 0.0f; }
			cur = path;
		}
		orc_convolve_real(in, in_len, h, r->filt_len, &out[2 * i], out_len - i, n, 1);
	}
	free(h);
	return out_len;
}

/* ---- Channelizer / Synthesis ---- */
struct orc_chan {
	int m, block_len, h_len, synth;
	float *sub;  /* [m][h_len] real taps, reversed (ChannelizerBase.cpp:68-132) */
	float *hist; /* [m][h_len] complex */
};

static float cb_sinc(float x)
{
	if (x == 0.0f)
		return 0.999999999999f;
	return sin(M_PI * x) / (M_PI * x);
}

static orc_chan *chan_create(int m, int block_len, int h_len, int synth)
{
	orc_chan *c = (orc_chan *)calloc(1, sizeof(*c));
	c->m = m; c->block_len = block_len; c->h_len = h_len; c->synth = synth;
	size_t plen = (size_t)m * h_len;
	float *proto = (float *)malloc(sizeof(float) * plen);
	float sum = 0.0f, scale;
	float midpt = (float)(plen - 1.0) / 2.0;
	float a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
	for (size_t i = 0; i < plen; i++) {
		proto[i] = cb_sinc(((float)i - midpt) / (float)m);
		proto[i] *= a0 - a1 * cos(2 * M_PI * i / (plen - 1)) + a2 * cos(4 * M_PI * i / (plen - 1)) -
			    a3 * cos(6 * M_PI * i / (plen - 1));
		sum += proto[i];
	}
	scale = (float)m / sum;
	c->sub = (float *)malloc(sizeof(float) * plen);
	for (int i = 0; i < h_len; i++)
		for (int n = 0; n < m; n++)
			c->sub[(size_t)n * h_len + (h_len - 1 - i)] = proto[(size_t)i * m + n] * scale;
	free(proto);
	c->hist = (float *)calloc((size_t)m * h_len * 2, sizeof(float));
	return c;
}

orc_chan *orc_channelizer_create(int m, int block_len, int h_len) { return chan_create(m, block_len, h_len, 0); }
orc_chan *orc_synthesis_create(int m, int block_len, int h_len) { return chan_create(m, block_len, h_len, 1); }

void orc_chan_destroy(orc_chan *c)
{
	if (!c) return;
	free(c->sub); free(c->hist); free(c);
}

int orc_chan_taps(orc_chan *c, int branch, float *out)
{
	memcpy(out, c->sub + (size_t)branch * c->h_len, sizeof(float) * c->h_len);
	return c->h_len;
}

/* forward DFT across the m branches for each time index (fft.c:55-85 plan geometry):
 * in[k*block_len + t] -> out[k*ostride + t] */
static void dft_branches(const float *in, int m, int block_len, float *out, int ostride)
{
	double *tr = (double *)malloc(sizeof(double) * 2 * m);
	for (int t = 0; t < block_len; t++) {
		for (int k = 0; k < m; k++) {
			double ar = 0.0, ai = 0.0;
			for (int j = 0; j < m; j++) {
				double ph = -2.0 * M_PI * (double)(((long)j * k) % m) / (double)m;
				double c = cos(ph), s = sin(ph);
				double xr = in[2 * ((size_t)j * block_len + t)], xi = in[2 * ((size_t)j * block_len + t) + 1];
				ar += xr * c - xi * s;
				ai += xr * s + xi * c;
			}
			tr[2 * k] = ar; tr[2 * k + 1] = ai;
		}
		for (int k = 0; k < m; k++) {
			out[2 * ((size_t)k * ostride + t)] = (float)tr[2 * k];
			out[2 * ((size_t)k * ostride + t) + 1] = (float)tr[2 * k + 1];
		}
	}
	free(tr);
}

/* per-branch FIR with history splice (Channelizer.cpp:85-94 / Synthesis.cpp:100-108):
 * x: [m][h_len + block_len] with the first h_len slots receiving the history */
static void branch_fir(orc_chan *c, float *x, float *y)
{
	const int m = c->m, bl = c->block_len, hl = c->h_len, row = hl + bl;
	float *h = (float *)malloc(sizeof(float) * 2 * hl);
	for (int i = 0; i < m; i++) {
		float *xi = x + (size_t)i * row * 2;
		memcpy(xi, c->hist + (size_t)i * hl * 2, sizeof(float) * 2 * hl);
		memcpy(c->hist + (size_t)i * hl * 2, xi + 2 * bl, sizeof(float) * 2 * hl); /* &hInputs[i][2*(blockLen-hLen)] */
		for (int k = 0; k < hl; k++) { h[2 * k] = c->sub[(size_t)i * hl + k]; h[2 * k + 1] = 0.0f; }
		orc_convolve_real(xi + 2 * hl, bl, h, hl, y + (size_t)i * bl * 2, bl, 0, bl);
	}
	free(h);
}

/* Channelizer::rotate (Channelizer.cpp:74-99) */
int orc_channelizer_rotate(orc_chan *c, const float *in, int m, int block_len, float *out)
{
	if (!c || c->synth || m != c->m || block_len != c->block_len)
		return -1;
	const int hl = c->h_len, row = hl + block_len;
	float *x = (float *)calloc((size_t)m * row * 2, sizeof(float));
	float *y = (float *)calloc((size_t)m * block_len * 2, sizeof(float));
	float *o = (float *)calloc((size_t)m * row * 2, sizeof(float));
	/* deinterleave :37-48: in[i*m+n] -> branch m-1-n sample i */
	for (int i = 0; i < block_len; i++)
		for (int n = 0; n < m; n++) {
			x[2 * ((size_t)(m - 1 - n) * row + hl + i)] = in[2 * ((size_t)i * m + n)];
			x[2 * ((size_t)(m - 1 - n) * row + hl + i) + 1] = in[2 * ((size_t)i * m + n) + 1];
		}
	branch_fir(c, x, y);
	dft_branches(y, m, block_len, o + 2 * hl, row);
	for (int ch = 0; ch < m; ch++)
		memcpy(out + (size_t)ch * block_len * 2, o + 2 * ((size_t)ch * row + hl), sizeof(float) * 2 * block_len);
	free(x); free(y); free(o);
	return 0;
}

/* Synthesis::rotate (Synthesis.cpp:85-114) */
int orc_synthesis_rotate(orc_chan *c, const float *in, int m, int block_len, float *out)
{
	if (!c || !c->synth || m != c->m || block_len != c->block_len)
		return -1;
	const int hl = c->h_len, row = hl + block_len;
	float *x = (float *)calloc((size_t)m * row * 2, sizeof(float));
	float *y = (float *)calloc((size_t)m * block_len * 2, sizeof(float));
	dft_branches(in, m, block_len, x + 2 * hl, row);
	branch_fir(c, x, y);
	/* interleave :38-49 */
	for (int i = 0; i < block_len; i++)
		for (int n = 0; n < m; n++) {
			out[2 * ((size_t)i * m + n)] = y[2 * ((size_t)n * block_len + i)];
			out[2 * ((size_t)i * m + n) + 1] = y[2 * ((size_t)n * block_len + i) + 1];
		}
	free(x); free(y);
	return 0;
}
