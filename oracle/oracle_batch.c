/*
 * oracle/oracle_batch.c — pthread batch loops over the per-burst oracle functions, with the same
 * array layout as the product's C ABI (include/trxb200.h) and as oracle/ref_capi.cpp.
 * TEST INFRASTRUCTURE ONLY (see oracle_trx.h).
 */
#include <pthread.h>
#include <string.h>
#include "oracle_trx.h"

typedef void (*item_fn)(int b, void *ctx);
struct span { int lo, hi; item_fn f; void *ctx; };

static void *span_run(void *a)
{
	struct span *s = (struct span *)a;
	for (int b = s->lo; b < s->hi; b++)
		s->f(b, s->ctx);
	return NULL;
}

static void parallel_for(int n, int nthreads, item_fn f, void *ctx)
{
	if (nthreads > 256) nthreads = 256;
	if (nthreads <= 1 || n < 2 * nthreads) {
		struct span s = { 0, n, f, ctx };
		span_run(&s);
		return;
	}
	pthread_t th[256];
	struct span sp[256];
	int per = (n + nthreads - 1) / nthreads, nt = 0;
	for (int t = 0; t < nthreads; t++) {
		int lo = t * per, hi = lo + per > n ? n : lo + per;
		if (lo >= hi) break;
		struct span s = { lo, hi, f, ctx };
		sp[nt] = s;
		pthread_create(&th[nt], NULL, span_run, &sp[nt]);
		nt++;
	}
	for (int t = 0; t < nt; t++)
		pthread_join(th[t], NULL);
}

struct dd {
	const float *bursts; int stride, blen; const uint8_t *type, *tsc; const uint16_t *max_toa; float thresh; int sps;
	int32_t *rc; float *amp, *toa; uint8_t *tsc_out; float *ci; uint8_t *flags; float *soft; int soft_stride;
	int32_t *nsoft; int do_detect, do_demod;
};

static void dd_item(int b, void *ctx)
{
	struct dd *d = (struct dd *)ctx;
	const ocf *burst = (const ocf *)(d->bursts + (size_t)b * d->stride * 2);
	orc_ebp ebp;
	int r;
	if (d->do_detect) {
		int fl = 0;
		memset(&ebp, 0, sizeof(ebp));
		r = orc_detect_any_burst(burst, d->blen, d->tsc[b], d->thresh, d->sps, d->type[b], d->max_toa[b], &ebp, &fl);
		d->rc[b] = r;
		if (d->flags) d->flags[b] = (uint8_t)fl;
	} else {
		r = d->rc[b];
		ebp.amp.r = d->amp[2 * b]; ebp.amp.i = d->amp[2 * b + 1];
		ebp.toa = d->toa[b]; ebp.tsc = 0; ebp.ci = d->ci[b];
	}
	if (d->do_demod) {
		d->nsoft[b] = 0;
		if (r > 0) {
			float soft[444];
			int ns = orc_demod_any_burst(burst, d->blen, r, d->sps, &ebp, soft);
			if (ns > 0) {
				memcpy(d->soft + (size_t)b * d->soft_stride, soft,
				       sizeof(float) * (ns < d->soft_stride ? ns : d->soft_stride));
				d->nsoft[b] = ns;
			}
		}
	}
	if (d->do_detect) {
		d->amp[2 * b] = ebp.amp.r; d->amp[2 * b + 1] = ebp.amp.i;
		d->toa[b] = ebp.toa; d->tsc_out[b] = ebp.tsc;
	}
	d->ci[b] = ebp.ci;
}

int orc_detect_batch(const float *bursts, int stride, int blen, int n, const uint8_t *type, const uint8_t *tsc,
		     const uint16_t *max_toa, float thresh, int sps, int32_t *rc, float *amp, float *toa,
		     uint8_t *tsc_out, float *ci, uint8_t *flags, int nthreads)
{
	orc_setup();
	struct dd d = { bursts, stride, blen, type, tsc, max_toa, thresh, sps, rc, amp, toa, tsc_out, ci, flags, NULL, 0, NULL, 1, 0 };
	parallel_for(n, nthreads, dd_item, &d);
	return n;
}

/* detectSCHBurst(SCH_DETECT_FULL) per burst (single-threaded: a test-size helper) */
int orc_detect_sch_batch(const float *bursts, int stride, int blen, int n, float thresh, int32_t *rc, float *amp, float *toa,
			 float *ci, uint8_t *flags)
{
	orc_setup();
	for (int b = 0; b < n; b++) {
		orc_ebp ebp;
		int fl = 0;
		memset(&ebp, 0, sizeof(ebp));
		rc[b] = orc_detect_sch_burst((const ocf *)(bursts + (size_t)b * stride * 2), blen, thresh, 4, &ebp, &fl);
		amp[2 * b] = ebp.amp.r; amp[2 * b + 1] = ebp.amp.i;
		toa[b] = ebp.toa; ci[b] = ebp.ci;
		if (flags) flags[b] = (uint8_t)fl;
	}
	return n;
}

/* detectSCHBurst(SCH_DETECT_BUFFER) per capture of in_len samples */
int orc_detect_sch_buffer_batch(const float *bursts, int stride, int in_len, int n, float thresh, int32_t *rc, float *amp, float *toa,
				float *ci, uint8_t *flags)
{
	orc_setup();
	for (int b = 0; b < n; b++) {
		orc_ebp ebp;
		int fl = 0;
		memset(&ebp, 0, sizeof(ebp));
		rc[b] = orc_detect_sch_buffer((const ocf *)(bursts + (size_t)b * stride * 2), in_len, thresh, &ebp, &fl);
		amp[2 * b] = ebp.amp.r; amp[2 * b + 1] = ebp.amp.i;
		toa[b] = ebp.toa; ci[b] = ebp.ci;
		if (flags) flags[b] = (uint8_t)fl;
	}
	return n;
}

int orc_demod_batch(const float *bursts, int stride, int blen, int n, const int32_t *rc, const float *amp,
		    const float *toa, float *ci, int sps, float *soft, int soft_stride, int32_t *nsoft, int nthreads)
{
	orc_setup();
	struct dd d = { bursts, stride, blen, NULL, NULL, NULL, 0, sps, (int32_t *)rc, (float *)amp, (float *)toa, NULL, ci, NULL,
			soft, soft_stride, nsoft, 0, 1 };
	parallel_for(n, nthreads, dd_item, &d);
	return n;
}

int orc_detect_demod_batch(const float *bursts, int stride, int blen, int n, const uint8_t *type, const uint8_t *tsc,
			   const uint16_t *max_toa, float thresh, int sps, int32_t *rc, float *amp, float *toa,
			   uint8_t *tsc_out, float *ci, uint8_t *flags, float *soft, int soft_stride, int32_t *nsoft,
			   int nthreads)
{
	orc_setup();
	struct dd d = { bursts, stride, blen, type, tsc, max_toa, thresh, sps, rc, amp, toa, tsc_out, ci, flags, soft, soft_stride,
			nsoft, 1, 1 };
	parallel_for(n, nthreads, dd_item, &d);
	return n;
}

struct md { const uint8_t *bits; int nbits; float *out; int edge; };
static void md_item(int b, void *ctx)
{
	struct md *m = (struct md *)ctx;
	ocf *o = (ocf *)(m->out + (size_t)b * 1250);
	if (m->edge) orc_modulate_edge(m->bits + (size_t)b * m->nbits, m->nbits, 4, 0, o, 625);
	else orc_modulate_burst(m->bits + (size_t)b * m->nbits, m->nbits, 8, 4, 0, o, 625);
}

int orc_modulate_gmsk_batch(const uint8_t *bits, int nbits, int n, float *out, int nthreads)
{
	orc_setup();
	struct md m = { bits, nbits, out, 0 };
	parallel_for(n, nthreads, md_item, &m);
	return n;
}

int orc_modulate_edge_batch(const uint8_t *bits, int nbits, int n, float *out, int nthreads)
{
	orc_setup();
	struct md m = { bits, nbits, out, 1 };
	parallel_for(n, nthreads, md_item, &m);
	return n;
}
