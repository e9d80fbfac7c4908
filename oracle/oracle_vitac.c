/*
 * oracle/oracle_vitac.c — plain-C restatement of Transceiver52M/grgsm_vitac/
 * (gr-gsm derived CIR search, matched filter and 16-state MLSE Viterbi).
 * TEST INFRASTRUCTURE ONLY (see oracle_trx.h).
 *
 * std::complex<float> arithmetic is restated component-wise: operator* is
 * (ac-bd, ad+bc); abs() is hypotf, which glibc evaluates as
 * (float)sqrt((double)x*x + (double)y*y); conj(r)/gr_complex(L,0) is component-wise /L
 * (SURVEY.md Appendix B).
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include "oracle_trx.h"

#define OSR 4		 /* grgsm_vitac.h:63 */
#define CIR_LEN 5	 /* constants.h:74 CHAN_IMP_RESP_LENGTH */
#define N_TRAIN_BITS 26
#define N_SYNC_BITS 64
#define N_ACCESS_BITS 41
#define TRAIN_POS (3 + 58 + 5) /* constants.h:66 */
#define TRAIN_BEGINNING 5

/* constants.h:80-91,113-125 */
static const unsigned char SYNC_BITS[N_SYNC_BITS] = {
	1, 0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1,
	0, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1 };
static const unsigned char ACCESS_BITS[N_ACCESS_BITS] = {
	0, 1, 0, 0, 1, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 0,
	0, 1, 1, 1, 1, 0, 0, 0 };
/* NB: row 1 differs from GSM::gTrainingSequence[1] (GSMCommon.cpp:37) at bit 20 - the reference's
 * gr-gsm table says ...00010010111, the 3GPP sequence is ...00010110111.  Restated as found. */
static const char *TRAIN_STR[9] = {
	"00100101110000100010010111", "00101101110111100010010111", "01000011101110100100001110",
	"01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
	"10100111110110001010011111", "11101111000100101110111100", "01110001011100010111000101" /* dummy */
};

static ocf norm_seq[9][N_TRAIN_BITS], acc_seq[N_ACCESS_BITS], sch_seq[N_SYNC_BITS];
static volatile int v_ready = 0;
static pthread_mutex_t v_lock = PTHREAD_MUTEX_INITIALIZER;

static inline ocf cmul(ocf a, ocf b) { ocf r = { a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r }; return r; }
static inline float cabsf_(ocf a) { return (float)sqrt((double)a.r * a.r + (double)a.i * a.i); }

/* gmsk_mapper grgsm_vitac.cpp:125-146 */
static void gmsk_mapper(const unsigned char *in, int n, ocf *out, ocf start)
{
	const ocf j = { 0.0f, 1.0f };
	out[0] = start;
	int prev = 2 * in[0] - 1;
	for (int i = 1; i < n; i++) {
		int cur = 2 * in[i] - 1;
		int enc = cur * prev;
		ocf e = { (float)enc, 0.0f };
		out[i] = cmul(cmul(j, e), out[i - 1]);
		prev = cur;
	}
}

/* initvita grgsm_vitac.cpp:51-80 */
static void vitac_setup(void)
{
	if (v_ready)
		return;
	pthread_mutex_lock(&v_lock);
	if (!v_ready) {
		ocf mj = { 0.0f, -1.0f };
		gmsk_mapper(SYNC_BITS, N_SYNC_BITS, sch_seq, mj);
		for (int i = 0; i < N_SYNC_BITS; i++) sch_seq[i].i = -sch_seq[i].i;
		gmsk_mapper(ACCESS_BITS, N_ACCESS_BITS, acc_seq, mj);
		for (int i = 0; i < N_ACCESS_BITS; i++) acc_seq[i].i = -acc_seq[i].i;
		for (int t = 0; t < 9; t++) {
			unsigned char b[N_TRAIN_BITS];
			for (int i = 0; i < N_TRAIN_BITS; i++) b[i] = TRAIN_STR[t][i] == '1';
			ocf sp = { b[0] == 0 ? 1.0f : -1.0f, 0.0f };
			gmsk_mapper(b, N_TRAIN_BITS, norm_seq[t], sp);
			for (int i = 0; i < N_TRAIN_BITS; i++) norm_seq[t][i].i = -norm_seq[t][i].i;
		}
		__sync_synchronize();
		v_ready = 1;
	}
	pthread_mutex_unlock(&v_lock);
}

int orc_get_vitac_table(int which, int idx, float *out)
{
	vitac_setup();
	const ocf *p; int n;
	if (which == 0) { p = norm_seq[idx]; n = N_TRAIN_BITS; }
	else if (which == 1) { p = acc_seq; n = N_ACCESS_BITS; }
	else { p = sch_seq; n = N_SYNC_BITS; }
	memcpy(out, p, sizeof(ocf) * n);
	return n;
}

/* correlate_sequence grgsm_vitac.cpp:148-156 */
static ocf correlate_sequence(const ocf *seq, int len, const ocf *in)
{
	ocf r = { 0.0f, 0.0f };
	for (int i = 0; i < len; i++) {
		ocf p = cmul(seq[i], in[i * OSR]);
		r.r += p.r; r.i += p.i;
	}
	ocf o = { r.r / (float)len, -r.i / (float)len };
	return o;
}

/* get_chan_imp_resp grgsm_vitac.cpp:183-235 */
static int get_cir(const ocf *in, ocf *cir, int s0, int s1, const ocf *tseq, int tlen, float *corr_max)
{
	const int nwin = s1 - s0, wl = CIR_LEN * OSR;
	ocf *cb = (ocf *)malloc(sizeof(ocf) * nwin);
	float *pb = (float *)malloc(sizeof(float) * nwin);
	float *we = (float *)malloc(sizeof(float) * nwin);
	int nwe = 0;
	for (int i = 0; i < nwin; i++) {
		cb[i] = correlate_sequence(tseq, tlen, &in[s0 + i]);
		pb[i] = (float)pow((double)cabsf_(cb[i]), 2.0);
	}
	float ws = 0;
	for (int i = 0; i < wl; i++) ws += pb[i];
	we[nwe++] = ws;
	for (int i = wl; i < nwin; i++) {
		ws += pb[i] - pb[i - wl];
		we[nwe++] = ws;
	}
	int best = 0;
	for (int i = 1; i < nwe; i++) /* std::max_element: first maximum */
		if (we[best] < we[i]) best = i;
	float mc = 0;
	for (int i = 0; i < wl; i++) {
		ocf c = cb[best + i];
		if (cabsf_(c) > mc) mc = cabsf_(c);
		cir[i] = c;
	}
	*corr_max = mc;
	free(cb); free(pb); free(we);
	return s0 + best;
}

/* get_norm_chan_imp_resp :265-274 / get_access_imp_resp :246-255 */
static int get_norm_cir(const ocf *in, ocf *cir, float *cm, int bcc)
{
	const int c = TRAIN_POS;
	return get_cir(in, cir, (c - 5) * OSR + 1, (c + 5 + CIR_LEN) * OSR, &norm_seq[bcc][TRAIN_BEGINNING],
		       N_TRAIN_BITS - 2 * TRAIN_BEGINNING, cm) - c * OSR;
}
static int get_access_cir(const ocf *in, ocf *cir, float *cm, int max_delay)
{
	const int c = 8 + 5;
	return get_cir(in, cir, (c - 5) * OSR + 1, (c + 5 + CIR_LEN + max_delay) * OSR, &acc_seq[TRAIN_BEGINNING],
		       N_ACCESS_BITS - 2 * TRAIN_BEGINNING, cm) - c * OSR;
}

/* viterbi_detector viterbi_detector.cc:63-392.  The 32 unrolled ACS statements follow one pattern:
 *   imag step, new state s, p = s>>1:  even s: c1 = old[p]   + x - inc[a[p]],  c2 = old[p+8] + x + inc[7-a[p]]
 *                                     odd  s: c1 = old[p]   - x + inc[a[p]],  c2 = old[p+8] - x - inc[7-a[p]]
 *       a = {2,3,0,1,6,7,4,5}
 *   real step:                        even s: c1 = old[p]   - x - inc[7-p],   c2 = old[p+8] - x + inc[p]
 *                                     odd  s: c1 = old[p]   + x + inc[7-p],   c2 = old[p+8] + x - inc[p]
 * each evaluated left to right in float32. */
void orc_viterbi(const float *input, int n, const float *rhh_, int start_state, float *output)
{
	const ocf *in = (const ocf *)input, *rhh = (const ocf *)rhh_;
	static const int A[8] = { 2, 3, 0, 1, 6, 7, 4, 5 };
	static const unsigned parity[16] = { 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0 };
	float inc[8], pm1[16], pm2[16], *oldm = pm1, *newm = pm2, *tmp;
	float (*trans)[16] = (float(*)[16])malloc(sizeof(float) * 16 * (n > 0 ? n : 1));
	int real_imag = 0;
	for (int i = 0; i < 16; i++) pm1[i] = (-10e30);
	pm1[start_state] = 0;
	inc[0] = -rhh[1].i - rhh[2].r - rhh[3].i + rhh[4].r;
	inc[1] = rhh[1].i - rhh[2].r - rhh[3].i + rhh[4].r;
	inc[2] = -rhh[1].i + rhh[2].r - rhh[3].i + rhh[4].r;
	inc[3] = rhh[1].i + rhh[2].r - rhh[3].i + rhh[4].r;
	inc[4] = -rhh[1].i - rhh[2].r + rhh[3].i + rhh[4].r;
	inc[5] = rhh[1].i - rhh[2].r + rhh[3].i + rhh[4].r;
	inc[6] = -rhh[1].i + rhh[2].r + rhh[3].i + rhh[4].r;
	inc[7] = rhh[1].i + rhh[2].r + rhh[3].i + rhh[4].r;
	int k = 0;
	while (k < n) {
		real_imag = 1;
		float x = in[k].i;
		for (int s = 0; s < 16; s++) {
			int p = s >> 1;
			float c1, c2;
			if (!(s & 1)) { c1 = oldm[p] + x - inc[A[p]]; c2 = oldm[p + 8] + x + inc[7 - A[p]]; }
			else { c1 = oldm[p] - x + inc[A[p]]; c2 = oldm[p + 8] - x - inc[7 - A[p]]; }
			float d = c2 - c1;
			newm[s] = (d < 0) ? c1 : c2;
			trans[k][s] = d;
		}
		tmp = oldm; oldm = newm; newm = tmp;
		k++;
		if (k == n) break;
		real_imag = 0;
		x = in[k].r;
		for (int s = 0; s < 16; s++) {
			int p = s >> 1;
			float c1, c2;
			if (!(s & 1)) { c1 = oldm[p] - x - inc[7 - p]; c2 = oldm[p + 8] - x + inc[p]; }
			else { c1 = oldm[p] + x + inc[7 - p]; c2 = oldm[p + 8] + x - inc[p]; }
			float d = c2 - c1;
			newm[s] = (d < 0) ? c1 : c2;
			trans[k][s] = d;
		}
		tmp = oldm; oldm = newm; newm = tmp;
		k++;
	}
	/* stop states {4,12} viterbi_detector.cc:342-350 */
	unsigned best = 4;
	if (oldm[12] > oldm[4]) best = 12;
	/* traceback :371-391 */
	unsigned state = best;
	int out_bit = 0;
	k = n;
	while (k > 0) {
		k--;
		int decision = trans[k][state] > 0;
		output[k] = (decision != out_bit) ? -trans[k][state] : trans[k][state];
		out_bit = out_bit ^ real_imag ^ parity[state];
		state = (state >> 1) + (decision ? 8 : 0);
		real_imag = !real_imag;
	}
	free(trans);
}

/* detect_burst_generic grgsm_vitac.cpp:82-103 with autocorrelation :159-166 and mafi :168-181 */
static void detect_burst(const ocf *in, const ocf *cir, int burst_start, int8_t *out, int ss, int nbits)
{
	const int nt = CIR_LEN * OSR;
	ocf rt[CIR_LEN * OSR], rhh[CIR_LEN];
	for (int k = nt - 1; k >= 0; k--) {
		ocf a = { 0.0f, 0.0f };
		for (int i = k; i < nt; i++) {
			ocf c = { cir[i - k].r, -cir[i - k].i };
			ocf p = cmul(cir[i], c);
			a.r += p.r; a.i += p.i;
		}
		rt[k] = a;
	}
	for (int i = 0; i < CIR_LEN; i++) { rhh[i].r = rt[i * OSR].r; rhh[i].i = -rt[i * OSR].i; }
	ocf *fb = (ocf *)malloc(sizeof(ocf) * nbits);
	float *o = (float *)malloc(sizeof(float) * nbits);
	const ocf *x = in + burst_start;
	for (int n = 0; n < nbits; n++) {
		int a = n * OSR;
		ocf acc = { 0.0f, 0.0f };
		for (int ii = 0; ii < nt; ii++) {
			if (a + ii >= nbits * OSR) break;
			ocf p = cmul(x[a + ii], cir[ii]);
			acc.r += p.r; acc.i += p.i;
		}
		fb[n] = acc;
	}
	orc_viterbi((const float *)fb, nbits, (const float *)rhh, ss, o);
	for (int i = 0; i < nbits; i++) out[i] = o[i] > 0 ? -127 : 127;
	free(fb); free(o);
}

/* get_sch_chan_imp_resp grgsm_vitac.cpp:283-296: the SCH burst of a tracked cell (search centre SYNC_POS + 5, ten symbols
 * back, SYNC_SEARCH_RANGE = 30 forward, 54 of the 64 training symbols); corr_max is computed but not returned there */
static int get_sch_cir(const ocf *in, ocf *cir, float *cm)
{
	const int c = (3 + 39) + TRAIN_BEGINNING; /* SYNC_POS + TRAIN_BEGINNING, constants.h:53 */
	return get_cir(in, cir, (c - 10) * OSR, (c + 30) * OSR, &sch_seq[TRAIN_BEGINNING], N_SYNC_BITS - 2 * TRAIN_BEGINNING, cm) - c * OSR;
}

struct vjob { const float *bufs; int stride, offset, lo, hi, is_ab; const uint8_t *tsc; int max_delay, clo, chi;
	      int8_t *bits; int32_t *start; float *cmax, *cir; };

static void *vworker(void *arg)
{
	struct vjob *j = (struct vjob *)arg;
	const int nbits = j->is_ab == 1 ? 88 : 148; /* is_ab: 0 normal burst, 1 access burst, 2 SCH burst (ms_rx_lower.cpp:173-177) */
	for (int b = j->lo; b < j->hi; b++) {
		const ocf *in = (const ocf *)(j->bufs + (size_t)b * j->stride * 2) + j->offset;
		ocf cir[CIR_LEN * OSR];
		float cm = 0;
		int st = j->is_ab == 2 ? get_sch_cir(in, cir, &cm)
				       : j->is_ab ? get_access_cir(in, cir, &cm, j->max_delay) : get_norm_cir(in, cir, &cm, j->tsc[b]);
		if (st < j->clo) st = j->clo;
		if (st > j->chi) st = j->chi;
		detect_burst(in, cir, st, j->bits + (size_t)b * nbits, 3, nbits);
		j->start[b] = st;
		j->cmax[b] = cm;
		if (j->cir) memcpy(j->cir + (size_t)b * 40, cir, sizeof(cir));
	}
	return NULL;
}

/* detect_burst_nb / detect_burst_ab :105-123 with a given channel estimate and start (single-threaded helper) */
int orc_vitac_detect_ss_batch(const float *bufs, int stride, int offset, int n, int is_ab, const float *cir, const int32_t *start,
			      int ss, int8_t *bits)
{
	vitac_setup();
	const int nbits = is_ab ? 88 : 148;
	for (int b = 0; b < n; b++)
		detect_burst((const ocf *)(bufs + (size_t)b * stride * 2) + offset, (const ocf *)(cir + (size_t)b * 40), start[b],
			     bits + (size_t)b * nbits, ss, nbits);
	return n;
}
int orc_vitac_detect_batch(const float *bufs, int stride, int offset, int n, int is_ab, const float *cir, const int32_t *start,
			   int8_t *bits)
{
	return orc_vitac_detect_ss_batch(bufs, stride, offset, n, is_ab, cir, start, 3, bits);
}

int orc_vitac_batch(const float *bufs, int stride, int offset, int n, int is_ab, const uint8_t *tsc, int max_delay,
		    int clamp_lo, int clamp_hi, int8_t *bits, int32_t *start_out, float *corr_max, float *cir_out,
		    int nthreads)
{
	vitac_setup();
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	pthread_t th[256];
	struct vjob jobs[256];
	int per = (n + nthreads - 1) / nthreads, nt = 0;
	for (int t = 0; t < nthreads; t++) {
		int lo = t * per, hi = lo + per > n ? n : lo + per;
		if (lo >= hi) break;
		struct vjob j = { bufs, stride, offset, lo, hi, is_ab, tsc, max_delay, clamp_lo, clamp_hi, bits, start_out,
				  corr_max, cir_out };
		jobs[nt] = j;
		if (nthreads == 1) vworker(&jobs[nt]);
		else pthread_create(&th[nt], NULL, vworker, &jobs[nt]);
		nt++;
	}
	if (nthreads > 1)
		for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
	return n;
}

/* get_sch_buffer_chan_imp_resp grgsm_vitac.cpp:298-309 + detect_burst_nb at the position found (ms_rx_lower.cpp:168-177): the
 * first SCH acquisition over a capture of `len` samples (12 frames there).  Search windows 0 .. len - 8 * N_SYNC_BITS;
 * the returned start is relative to the burst's first sample (window - (SYNC_POS + 5) * OSR) and may be negative: rows
 * carry `offset` samples of head-room, the start is limited to what the row holds. */
int orc_vitac_sch_buffer_batch(const float *bufs, int stride, int offset, int len, int n, int8_t *bits, int32_t *start_out,
			       float *corr_max, float *cir_out)
{
	vitac_setup();
	const int c = (3 + 39) + TRAIN_BEGINNING;
	if (len - N_SYNC_BITS * 8 < CIR_LEN * OSR) return -1;
	for (int b = 0; b < n; b++) {
		const ocf *in = (const ocf *)(bufs + (size_t)b * stride * 2) + offset;
		ocf cir[CIR_LEN * OSR];
		float cm = 0;
		int st = get_cir(in, cir, 0, len - N_SYNC_BITS * 8, &sch_seq[TRAIN_BEGINNING], N_SYNC_BITS - 2 * TRAIN_BEGINNING, &cm) - c * OSR;
		start_out[b] = st;
		corr_max[b] = cm;
		if (cir_out) memcpy(cir_out + (size_t)b * 40, cir, sizeof(cir));
		if (bits) {
			int lo = -offset, hi = stride - offset - 148 * OSR;
			int sd = st < lo ? lo : (st > hi ? hi : st);
			detect_burst(in + sd, cir, 0, bits + (size_t)b * 148, 3, 148);
		}
	}
	return n;
}
