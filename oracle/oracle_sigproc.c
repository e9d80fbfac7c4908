/*
 * oracle/oracle_sigproc.c — plain-C restatement of Transceiver52M/sigProcLib.cpp
 * (table setup, GMSK/8-PSK modulators, burst detection, soft demodulation).
 * TEST INFRASTRUCTURE ONLY (see oracle_trx.h).  Every function cites the
 * reference lines it restates; float32 operation order follows the reference's
 * x86-64/SSE3 build (no FMA).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "oracle_trx.h"

#define CLIP_THRESH 30000.0f /* sigProcLib.cpp:49 */
static const float M_PI_F = (float)M_PI; /* sigProcLib.cpp:55 */

static orc_tables T;
static volatile int g_ready = 0;
static __thread int t_in_setup = 0; /* setup re-enters the public modulators on the same thread */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

/* ---- Complex.h helpers (Complex.h:73-112,144-150) ---- */
static inline ocf cmul(ocf a, ocf b) { ocf r = { a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r }; return r; }
static inline ocf cscale(ocf a, float s) { ocf r = { a.r * s, a.i * s }; return r; }
static inline float norm2(ocf a) { return a.i * a.i + a.r * a.r; }
static inline ocf cinv(ocf a) { float n = norm2(a); ocf r = { a.r / n, -a.i / n }; return r; }
static inline ocf cdiv(ocf a, ocf b) { return cmul(a, cinv(b)); }
static inline float cabs_(ocf a) { return sqrtf(norm2(a)); } /* ::sqrt(float) overload, Complex.h:121 */

/* ---- GSM constants (GSM/GSMCommon.cpp:35-68) ---- */
static const char *TSC_STR[8] = {
	"00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110",
	"01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
	"10100111110110001010011111", "11101111000100101110111100",
};
static const char *EDGE_TSC_STR[8] = {
	"111111001111111001111001001001111111111111001111111111001111111001111001001001",
	"111111001111001001111001001001111001001001001111111111001111001001111001001001",
	"111001111111111111001001001111001001001111001111111001111111111111001001001111",
	"111001111111111001001001001111001001111001111111111001111111111001001001001111",
	"111111111001001111001111001001001111111001111111111111111001001111001111001001",
	"111001111111001001001111001111001001111111111111111001111111001001001111001111",
	"001111001111111001001001001001111001001111111111001111001111111001001001001001",
	"001001001111001001001001111111111001111111001111001001001111001001001001111111",
};
static const char *DUMMY_TSC_STR = "01110001011100010111000101";
static const char *RACH_STR[3] = {
	"01001011011111111001100110101010001111000",
	"01010100111110001000011000101111001001101",
	"11101111001001110101011000001101101110111",
};
static const char *SCH_STR = "1011100101100010000001000000111100101101010001010111011000011011";

static int str_bits(const char *s, uint8_t *out)
{
	int n = 0;
	for (; s[n]; n++)
		out[n] = s[n] == '1';
	return n;
}

/* sigProcLib.cpp:66-75 */
static const ocf psk8_table[8] = {
	{ -0.70710678f, 0.70710678f }, { 0.0f, -1.0f }, { 0.0f, 1.0f }, { 0.70710678f, -0.70710678f },
	{ -1.0f, 0.0f }, { -0.70710678f, -0.70710678f }, { 0.70710678f, 0.70710678f }, { 1.0f, 0.0f },
};

/* ---- sinc table (sigProcLib.cpp:981-998) ---- */
static void gen_sinc_table(void)
{
	for (int i = 0; i < ORC_SINC_TABLESIZE; i++) {
		double x = (double)i / ORC_SINC_TABLESIZE * 8 * M_PI;
		double y = sin(x) / x;
		T.sinc[i] = isnan(y) ? 1.0 : y;
	}
	T.sinc[ORC_SINC_TABLESIZE] = 0.0f; /* static storage, never written */
}

static float sinc_tab(float x)
{
	float ax = fabsf(x);
	if ((double)ax >= 8 * M_PI)
		return 0.0f;
	int index = (int)floorf((float)((double)ax / (8 * M_PI) * ORC_SINC_TABLESIZE));
	return T.sinc[index];
}

/* ---- rotation tables (sigProcLib.cpp:191-216) ---- */
static void gen_rot_tables(void)
{
	double phase = 0.0;
	for (int i = 0; i < 625; i++) {
		T.rot4[i].r = cos(phase); T.rot4[i].i = sin(phase);
		T.rrot4[i].r = cos(-phase); T.rrot4[i].i = sin(-phase);
		phase += M_PI / 2.0 / 4.0;
	}
	phase = 0.0;
	for (int i = 0; i < 157; i++) {
		T.rot1[i].r = cos(phase); T.rot1[i].i = sin(phase);
		T.rrot1[i].r = cos(-phase); T.rrot1[i].i = sin(-phase);
		phase += M_PI / 2.0;
	}
}

/* ---- pulses (sigProcLib.cpp:405-543) ---- */
static void gen_pulses(void)
{
	static const double c0[16] = { 0.0, 4.46348606e-03, 2.84385729e-02, 1.03184855e-01, 2.56065552e-01,
				       4.76375085e-01, 7.05961177e-01, 8.71291644e-01, 9.29453645e-01,
				       8.71291644e-01, 7.05961177e-01, 4.76375085e-01, 2.56065552e-01,
				       1.03184855e-01, 2.84385729e-02, 4.46348606e-03 };
	static const double c1[8] = { 0.0, 8.16373112e-03, 2.84385729e-02, 5.64158904e-02,
				      7.05463553e-02, 5.64158904e-02, 2.84385729e-02, 8.16373112e-03 };
	static const double inv[5] = { 0.15884, -0.43176, 1.00000, -0.42608, 0.14882 };
	for (int i = 0; i < 16; i++) T.pulse4_c0[i] = (float)c0[i];
	for (int i = 0; i < 8; i++) T.pulse4_c1[i] = (float)c1[i];
	for (int i = 0; i < 5; i++) T.c0_inv[i] = (float)inv[i];

	/* sps = 1: :520-533 */
	int len = 4, sps = 1;
	float center = (float)(len - 1.0) / 2.0;
	for (int i = 0; i < len; i++) {
		float arg = ((float)i - center) / (float)sps;
		T.pulse1_c0[i] = (float)(0.96 * exp(-1.1380 * arg * arg - 0.527 * arg * arg * arg * arg));
	}
	float energy = 0.0f; /* vectorNorm2 :178-186, norm2 = i*i + r*r with i = 0 */
	for (int i = 0; i < len; i++)
		energy += 0.0f * 0.0f + T.pulse1_c0[i] * T.pulse1_c0[i];
	float avg = sqrtf(energy / sps);
	for (int i = 0; i < len; i++)
		T.pulse1_c0[i] /= avg;
}

/* ---- fractional delay filterbank (sigProcLib.cpp:1005-1044) ---- */
static void gen_delay_filters(void)
{
	const int h_len = 20;
	float a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
	for (int i = 0; i < ORC_DELAYFILTS; i++) {
		float sum = 0.0f;
		float *h = T.delay[i];
		for (int n = 0; n < h_len; n++) {
			float k = (float)n;
			int pos = h_len - 1 - n; /* *--itr from end */
			float s = sinc_tab(M_PI_F * (k - (float)h_len / 2.0 - (float)i / ORC_DELAYFILTS));
			/* Complex *= Real: r *= a (Complex.h:183-188), a computed in double then narrowed to Real */
			float w = (float)(a0 - a1 * cos(2 * M_PI * n / (h_len - 1)) + a2 * cos(4 * M_PI * n / (h_len - 1)) -
					  a3 * cos(6 * M_PI * n / (h_len - 1)));
			h[pos] = s * w;
			sum += h[pos];
		}
		for (int n = 0; n < h_len; n++)
			h[n] /= sum;
	}
}

/* ---- Resampler(1,4) partition (Resampler.cpp:47-96), used by downsampleBurst ---- */
static float rs_sinc(float x)
{
	if (x == 0.0)
		return 0.9999999999;
	return sin(M_PI * x) / (M_PI * x);
}

void orc_resampler_proto(int p, int q, int filt_len, float bw, float *parts /*[p][filt_len] reversed*/)
{
	int plen = p * filt_len;
	float *proto = (float *)malloc(sizeof(float) * plen);
	float sum = 0.0f, scale, cutoff;
	float a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
	cutoff = p > q ? (float)p : (float)q;
	float midpt = (plen - 1) / 2.0;
	for (int i = 0; i < plen; i++) {
		proto[i] = rs_sinc(((float)i - midpt) / cutoff * bw);
		proto[i] *= a0 - a1 * cos(2 * M_PI * i / (plen - 1)) + a2 * cos(4 * M_PI * i / (plen - 1)) -
			    a3 * cos(6 * M_PI * i / (plen - 1));
		sum += proto[i];
	}
	scale = p / sum;
	for (int i = 0; i < filt_len; i++)
		for (int n = 0; n < p; n++)
			parts[n * filt_len + (filt_len - 1 - i)] = proto[i * p + n] * scale; /* stored reversed */
	free(proto);
}

/* ---- span-mode convolve wrapper (sigProcLib.cpp:297-398) ---- */
int orc_convolve_sv(const ocf *x, int x_len, int x_head, const float *h, int h_len, int h_real, int h_aligned,
		    int span, int start, int len, ocf *y)
{
	int head = 0, tail = 0;
	(void)x_head; /* head-room only decides copy-vs-in-place in the reference; zeros either way */
	switch (span) {
	case 0: start = 0; head = h_len - 1; len = x_len; break;
	case 1: start = h_len / 2; head = start; tail = start; len = x_len; break;
	case 2:
		if (start < h_len - 1) head = h_len - start;
		if (start + len > x_len) tail = start + len - x_len;
		break;
	default: return -1;
	}
	/* work on a padded copy: max(head, h_len) zeros in front so every tap is addressable */
	int pad = head > h_len ? head : h_len;
	int tot = pad + x_len + tail + h_len;
	ocf *buf = (ocf *)calloc(tot, sizeof(ocf));
	memcpy(buf + pad, x, sizeof(ocf) * x_len);
	int rc;
	const float *xp = (const float *)(buf + pad);
	if (h_real && h_aligned)
		rc = orc_convolve_real(xp, x_len, h, h_len, (float *)y, len, start, len);
	else if (!h_real && h_aligned)
		rc = orc_convolve_complex(xp, x_len, h, h_len, (float *)y, len, start, len);
	else if (h_real)
		rc = orc_base_convolve_real(xp, x_len + tail, h, h_len, (float *)y, len, start, len);
	else
		rc = orc_base_convolve_complex(xp, x_len + tail, h, h_len, (float *)y, len, start, len);
	free(buf);
	return rc < 0 ? -1 : len;
}

/* real taps -> interleaved complex layout used by the kernels */
static void taps_cx(const float *h, int n, float *out)
{
	for (int i = 0; i < n; i++) { out[2 * i] = h[i]; out[2 * i + 1] = 0.0f; }
}

/* ---- GMSK rotate (sigProcLib.cpp:218-260) ---- */
static void gmsk_rotate(ocf *x, int n, int sps, int is_real)
{
	const ocf *rot = sps == 1 ? T.rot1 : T.rot4;
	for (int i = 0; i < n; i++)
		x[i] = is_real ? cscale(rot[i], x[i].r) : cmul(rot[i], x[i]);
}

/* rotateBurst (sigProcLib.cpp:558-580): unshaped, complex rotate path, 1-tap "empty" pulse via base_convolve_real */
static int rotate_burst(const uint8_t *bits, int nbits, int guard, int sps, ocf *out, int max_cf)
{
	int blen = sps * (nbits + guard);
	ocf *rot = (ocf *)calloc(blen, sizeof(ocf));
	for (int i = 0; i < nbits; i++)
		rot[i * sps].r = (float)(2.0 * (bits[i] & 1) - 1.0);
	gmsk_rotate(rot, blen, sps, 0);
	for (int i = 0; i < blen && i < max_cf; i++) { /* y = 0 + x*1.0 (mac_real) */
		out[i].r = 0.0f + rot[i].r * 1.0f;
		out[i].i = 0.0f + rot[i].i * 1.0f;
	}
	free(rot);
	return blen;
}

/* modulateBurstBasic (sigProcLib.cpp:938-967) */
static int modulate_basic(const uint8_t *bits, int nbits, int guard, int sps, ocf *out, int max_cf)
{
	const float *pulse = sps == 1 ? T.pulse1_c0 : T.pulse4_c0;
	int plen = sps == 1 ? 4 : 16;
	int blen = sps * (nbits + guard);
	ocf *b = (ocf *)calloc(blen, sizeof(ocf));
	for (int i = 0; i < nbits; i++)
		b[i * sps].r = (float)(2.0 * (bits[i] & 1) - 1.0);
	gmsk_rotate(b, blen, sps, 1);
	float hc[32];
	taps_cx(pulse, plen, hc);
	ocf *y = (ocf *)calloc(blen, sizeof(ocf));
	orc_convolve_sv(b, blen, plen, hc, plen, 1, 1, 0, 0, 0, y);
	memcpy(out, y, sizeof(ocf) * (blen < max_cf ? blen : max_cf));
	free(b); free(y);
	return blen;
}

/* modulateBurstLaurent (sigProcLib.cpp:595-670) */
static int modulate_laurent(const uint8_t *bits, int nbits, ocf *out, int max_cf)
{
	const int sps = 4, blen = 625;
	if (nbits > 156 || nbits < 2)
		return -1;
	ocf c0[625], c1[625], y0[625], y1[625];
	memset(c0, 0, sizeof(c0));
	memset(c1, 0, sizeof(c1));
	int p = 0;
	c0[p].r = (float)(2.0 * (0x00 & 0x01) - 1.0); p += sps;
	for (int i = 0; i < nbits; i++) { c0[p].r = (float)(2.0 * (bits[i] & 1) - 1.0); p += sps; }
	c0[p].r = (float)(2.0 * (0x00 & 0x01) - 1.0);
	gmsk_rotate(c0, blen, sps, 1);

	int q = sps * 2;
	float phase = 2.0 * ((0x01 & 0x01) ^ (0x01 & 0x01)) - 1.0;
	ocf j = { 0.0f, phase };
	c1[q] = cmul(c0[q], j); q += sps;
	for (int i = 2; i < nbits; i++) {
		phase = 2.0 * ((bits[i - 1] & 1) ^ (bits[i - 2] & 1)) - 1.0;
		j.r = 0.0f; j.i = phase;
		c1[q] = cmul(c0[q], j); q += sps;
	}
	phase = 2.0 * ((bits[nbits - 1] & 1) ^ (bits[nbits - 2] & 1)) - 1.0;
	j.r = 0.0f; j.i = phase;
	c1[q] = cmul(c0[q], j);

	float h0[32], h1[16];
	taps_cx(T.pulse4_c0, 16, h0);
	taps_cx(T.pulse4_c1, 8, h1);
	orc_convolve_sv(c0, blen, 16, h0, 16, 1, 1, 0, 0, 0, y0);
	orc_convolve_sv(c1, blen, 8, h1, 8, 1, 1, 0, 0, 0, y1);
	for (int i = 0; i < blen; i++) { y0[i].r += y1[i].r; y0[i].i += y1[i].i; }
	memcpy(out, y0, sizeof(ocf) * (blen < max_cf ? blen : max_cf));
	return blen;
}

/* modulateBurst (sigProcLib.cpp:970-979) */
int orc_modulate_burst(const uint8_t *bits, int nbits, int guard, int sps, int empty, ocf *out, int max_cf)
{
	orc_setup();
	if (empty)
		return rotate_burst(bits, nbits, guard, sps, out, max_cf);
	if (sps == 4)
		return modulate_laurent(bits, nbits, out, max_cf);
	return modulate_basic(bits, nbits, guard, sps, out, max_cf);
}

/* mapEdgeSymbols/rotateEdgeBurst/shapeEdgeBurst/modulateEdgeBurst (sigProcLib.cpp:672-763,917-936) */
int orc_modulate_edge(const uint8_t *bits, int nbits, int sps, int empty, ocf *out, int max_cf)
{
	orc_setup();
	if ((sps != 4) && !empty)
		return -1;
	if (nbits % 3)
		return -1;
	int nsym = nbits / 3;
	ocf *sym = (ocf *)malloc(sizeof(ocf) * (nsym + 1));
	for (int i = 0; i < nsym; i++) {
		unsigned idx = (bits[3 * i] & 1) | ((bits[3 * i + 1] & 1) << 1) | ((bits[3 * i + 2] & 1) << 2);
		sym[i] = psk8_table[idx];
	}
	int rc;
	if (empty) {
		int blen = nsym * sps;
		ocf *b = (ocf *)calloc(blen, sizeof(ocf));
		for (int i = 0; i < nsym; i++) {
			float phase = i * 3.0f * M_PI / 8.0f;
			ocf rot = { cosf(phase), sinf(phase) }; /* cos(float) resolves to the float overload */
			b[i * sps] = cmul(sym[i], rot);
		}
		memcpy(out, b, sizeof(ocf) * (blen < max_cf ? blen : max_cf));
		free(b);
		rc = blen;
	} else {
		int nsamps = 625;
		if (nsym * 4 > nsamps)
			nsym = 156;
		ocf b[625], y[625];
		memset(b, 0, sizeof(b));
		for (int i = 0; i < nsym; i++) {
			float phase = i * 3.0f * M_PI / 8.0f;
			ocf rot = { cosf(phase), sinf(phase) };
			b[4 + 4 * i] = cmul(sym[i], rot);
		}
		float h0[32];
		taps_cx(T.pulse4_c0, 16, h0);
		orc_convolve_sv(b, 625, 16, h0, 16, 1, 1, 0, 0, 0, y);
		memcpy(out, y, sizeof(ocf) * (625 < max_cf ? 625 : max_cf));
		rc = 625;
	}
	free(sym);
	return rc;
}

/* ---- peak machinery ---- */
/* interpolatePoint sigProcLib.cpp:1100-1118 (complex input path) */
static ocf interpolate_point(const ocf *sig, int size, float ix)
{
	int start = (int)(floor(ix) - 10);
	if (start < 0) start = 0;
	int end = (int)(floor(ix) + 11);
	if ((unsigned)end > (unsigned)(size - 1)) end = size - 1;
	ocf p = { 0.0f, 0.0f };
	for (int i = start; i < end; i++) {
		float s = sinc_tab(M_PI_F * (i - ix));
		p.r += sig[i].r * s;
		p.i += sig[i].i * s;
	}
	return p;
}

/* fastPeakDetect :1120-1139 */
static ocf fast_peak(const ocf *x, int n, float *index)
{
	float max = 0.0f;
	ocf amp = { 0.0f, 0.0f };
	int idx = -1;
	for (int i = 0; i < n; i++) {
		float v = norm2(x[i]);
		if (v > max) { max = v; idx = i; amp = x[i]; }
	}
	*index = (float)idx;
	return amp;
}

static int near_tie(float a, float b)
{
	float m = fabsf(a) > fabsf(b) ? fabsf(a) : fabsf(b);
	return fabsf(a - b) <= 4.0f * 1.1920929e-7f * m;
}

/* peakDetect :1141-1186 */
static ocf peak_detect(const ocf *x, int n, float *peak_index, int *tie)
{
	float maxp = 0.0f, max_index = -1;
	for (int i = 0; i < n; i++) {
		float sp = norm2(x[i]);
		if (sp > maxp) { maxp = sp; max_index = i; }
	}
	float early = max_index - 1, late = max_index + 1;
	float incr = 0.5;
	while (incr > 1.0 / 1024.0) {
		ocf e = interpolate_point(x, n, early), l = interpolate_point(x, n, late);
		float ne = norm2(e), nl = norm2(l);
		if (tie && near_tie(ne, nl)) *tie = 1;
		if (ne < nl) early += incr;
		else if (ne > nl) early -= incr;
		else break;
		incr /= 2.0;
		late = early + 2.0;
	}
	max_index = early + 1.0;
	*peak_index = max_index;
	return interpolate_point(x, n, max_index);
}

/* computePeakRatio :1541-1571 (sps = 1) */
static float peak_ratio(const ocf *corr, int n, int sps, float toa, ocf amp)
{
	int num = 0;
	float avg = 0.0f;
	if (toa < 0.0 || toa > n)
		return 0.0f;
	int pk = (int)rint(toa);
	for (int i = 2 * sps; i <= 5 * sps; i++) {
		if (pk - i >= 0) { avg += norm2(corr[pk - i]); num++; }
		if (pk + i < n) { avg += norm2(corr[pk + i]); num++; }
	}
	if (num < 5)
		return 0.0f;
	float rms = sqrtf(avg / (float)num) + 0.00001;
	return cabs_(amp) / rms;
}

/* computeCI :1608-1639 */
static float compute_ci(const ocf *burst, int blen, const orc_corrseq *sync, float toa, int start, ocf xcorr)
{
	const int N = sync->len;
	const int ps = start + 1 - N + (int)roundf(toa);
	if (ps < 0) return 0;
	if (ps + N > blen) return 0;
	float S = 0.0f;
	for (int i = 0, j = ps; i < N; i++, j++)
		S += norm2(burst[j]);
	S /= N;
	float C = norm2(xcorr) / ((N - 1) * cabs_(sync->gain));
	return 3.0103f * log2f(C / (S - C));
}

/* downsampleBurst :1587-1601 + Resampler::rotate (Resampler.cpp:131-150) for (1,4) */
int orc_downsample_burst(const ocf *in, int blen, ocf *out)
{
	orc_setup();
	ocf buf[16 + 624];
	float h[32];
	memset(buf, 0, sizeof(buf));
	memcpy(buf + 16, in, sizeof(ocf) * (blen < 624 ? blen : 624));
	taps_cx(T.dnsamp, 16, h);
	for (int i = 0; i < ORC_DEC_LEN; i++)
		orc_convolve_real((const float *)(buf + 16), 624, h, 16, (float *)&out[i], ORC_DEC_LEN - i, 4 * i, 1);
	return ORC_DEC_LEN;
}

/* detectBurst :1649-1709 on the already decimated (1 sps) burst */
static int detect_burst(const ocf *dec, int dlen, const orc_corrseq *sync, float thresh, int start, int len,
			orc_ebp *ebp, int *flags)
{
	const int sps = 1;
	ocf *corr = (ocf *)calloc(len, sizeof(ocf));
	int rc = 1;
	if (orc_convolve_sv(dec, dlen, 64, (const float *)sync->seq, sync->len, 0, 1, 2, start, len, corr) < 0) {
		rc = -1;
		goto out;
	}
	ebp->amp = fast_peak(corr, len, &ebp->toa);
	if ((ebp->toa < 3 * sps) || (ebp->toa > len - 3 * sps)) { rc = 0; goto out; }
	{
		float ratio = peak_ratio(corr, len, sps, ebp->toa, ebp->amp);
		if (flags && fabsf(ratio - thresh) < 1e-5f) *flags |= ORC_FLAG_THRESH_EDGE;
		if (ratio < thresh) { rc = 0; goto out; }
	}
	{
		int tie = 0;
		ocf xcorr = peak_detect(corr, len, &ebp->toa, &tie);
		if (flags && tie) *flags |= ORC_FLAG_BISECT_TIE;
		ebp->ci = compute_ci(dec, dlen, sync, ebp->toa, start, xcorr);
		ebp->amp = cdiv(xcorr, sync->gain);
		ebp->toa = ebp->toa - sync->toa;
	}
out:
	free(corr);
	return rc;
}

/* detectGeneralBurst :1732-1771 (+ maxAmplitude :1711-1722) */
static int detect_general(const ocf *burst, int blen, float thresh, int sps, int target, int head, int tail,
			  const orc_corrseq *sync, orc_ebp *ebp, int *flags)
{
	if (sps != 1 && sps != 4)
		return -ORC_SIGERR_UNSUPPORTED;
	float maxa = 0.0f;
	for (int i = 0; i < blen; i++) {
		if (fabsf(burst[i].r) > maxa) maxa = fabsf(burst[i].r);
		if (fabsf(burst[i].i) > maxa) maxa = fabsf(burst[i].i);
	}
	int clipping = maxa > CLIP_THRESH;
	if (clipping && flags) *flags |= ORC_FLAG_CLIP;
	int start = target - head - 1, len = head + tail;
	int rc;
	if (sps == 4) {
		ocf dec[ORC_DEC_LEN];
		orc_downsample_burst(burst, blen, dec);
		rc = detect_burst(dec, ORC_DEC_LEN, sync, thresh, start, len, ebp, flags);
	} else {
		rc = detect_burst(burst, blen, sync, thresh, start, len, ebp, flags);
	}
	if (rc < 0)
		return -ORC_SIGERR_INTERNAL;
	if (!rc) {
		ebp->amp.r = ebp->amp.i = 0.0f;
		ebp->toa = 0.0f;
		ebp->ci = 0.0f;
		return clipping ? -ORC_SIGERR_CLIP : ORC_SIGERR_NONE;
	}
	ebp->toa -= head;
	return 1;
}

/* detectSCHBurst :1805-1861 in its SCH_DETECT_FULL state (the single-burst search: head = target - 1 symbols,
 * tail = 39 + 3 + 9): downsampleBurst(burst, len * 4, len) is the default 624 -> 156 decimation, the correlation
 * starts at sample 0 of the decimated burst, so the 64-tap sequence reaches 63 samples before it - zeros (convolve's
 * CUSTOM span prepends head-room :325-329).  Returns detectBurst's rc (1 on a hit). */
int orc_detect_sch_burst(const ocf *burst, int blen, float thresh, int sps, orc_ebp *ebp, int *flags)
{
	orc_setup();
	if (sps != 4)
		return -1;
	const int target = 3 + 39 + 64, head = target - 1, tail = 39 + 3 + 9;
	const int start = (target - head) - 1, len = head + tail;
	ocf dec[ORC_DEC_LEN];
	orc_downsample_burst(burst, blen, dec);
	int rc = detect_burst(dec, ORC_DEC_LEN, &T.sch, thresh, start, len, ebp, flags);
	if (rc < 0)
		return -1;
	if (!rc) {
		ebp->amp.r = ebp->amp.i = 0.0f;
		ebp->toa = 0.0f;
		return 0;
	}
	ebp->toa -= head; /* "Subtract forward search bits from delay" */
	return rc;
}

/* detectSCHBurst in its SCH_DETECT_BUFFER state :1805-1861 (first acquisition over a 12-frame capture): the whole capture is
 * decimated (downsampleBurst(burst, len * 4, len), len = 12 * 8 * 625 / 4 = 15000) and correlated with the 64-symbol sequence
 * from start 0 (the head of the correlation reads the zero prefix convolve() prepends, :312-334); toa is reported relative to
 * the burst start (-(3 + 39 + 64)).  in_len: samples of the capture (60000 in the reference). */
int orc_detect_sch_buffer(const ocf *burst, int in_len, float thresh, orc_ebp *ebp, int *flags)
{
	orc_setup();
	const int len = in_len / 4;
	if (len < 64 || len > 4096 * 4) return -1;
	ocf *buf = (ocf *)calloc(16 + (size_t)in_len, sizeof(ocf));
	ocf *dec = (ocf *)calloc(len, sizeof(ocf));
	float h[32];
	memcpy(buf + 16, burst, sizeof(ocf) * (size_t)in_len);
	taps_cx(T.dnsamp, 16, h);
	for (int i = 0; i < len; i++)
		orc_convolve_real((const float *)(buf + 16), in_len, h, 16, (float *)&dec[i], len - i, 4 * i, 1);
	int rc = detect_burst(dec, len, &T.sch, thresh, 0, len, ebp, flags);
	free(buf); free(dec);
	if (rc < 0)
		return -1;
	if (!rc) {
		ebp->amp.r = ebp->amp.i = 0.0f;
		ebp->toa = 0.0f;
		return 0;
	}
	ebp->toa = ebp->toa - (3 + 39 + 64);
	return rc;
}

/* detectAnyBurst :1926-1957 with analyzeTrafficBurst :1887, detectRACHBurst :1782,
 * detectEdgeBurst :1906, detectDummyBurst :1863 */
int orc_detect_any_burst(const ocf *burst, int blen, unsigned tsc, float thresh, int sps, int type, unsigned max_toa,
			 orc_ebp *ebp, int *flags)
{
	orc_setup();
	int rc = 0;
	switch (type) {
	case ORC_EDGE:
		if (tsc > 7) rc = -ORC_SIGERR_UNSUPPORTED;
		else {
			ebp->tsc = tsc;
			rc = detect_general(burst, blen, thresh, sps, 3 + 58 + 16 + 5, 6, 6 + max_toa,
					    &T.edge_midamble[tsc], ebp, flags);
		}
		if (rc > 0)
			break;
		type = ORC_TSC;
		/* fall through */
	case ORC_TSC:
		if (tsc > 7) { rc = -ORC_SIGERR_UNSUPPORTED; break; }
		ebp->tsc = tsc;
		rc = detect_general(burst, blen, thresh, sps, 3 + 58 + 16 + 5, 10, 6 + max_toa, &T.midamble[tsc], ebp,
				    flags);
		break;
	case ORC_EXT_RACH:
	case ORC_RACH: {
		int num = type == ORC_EXT_RACH ? 3 : 1;
		for (int i = 0; i < num; i++) {
			rc = detect_general(burst, blen, thresh, sps, 8 + 40, 8, 8 + max_toa, &T.rach[i], ebp, flags);
			if (rc > 0) { ebp->tsc = i; break; }
		}
		break;
	}
	case ORC_IDLE:
		ebp->tsc = 0;
		rc = detect_general(burst, blen, thresh, sps, 3 + 58 + 16 + 5, 10, 6 + max_toa, &T.dummy, ebp, flags);
		break;
	default:
		break;
	}
	if (rc > 0)
		return type;
	return rc;
}

/* energyDetect :1573-1585 */
float orc_energy_detect(const ocf *burst, int blen, unsigned window)
{
	float energy = 0.0f;
	if (window == 0) return 0.0f;
	if (window > (unsigned)blen) window = blen;
	for (unsigned i = 0; i < window; i++)
		energy += norm2(burst[4 * i]);
	return energy / window;
}

/* delayVector :1046-1098 */
int orc_delay_vector(const ocf *in, int len, float delay, ocf *out)
{
	orc_setup();
	int whole = floor(delay);
	float frac = delay - whole;
	ocf *shift = (ocf *)malloc(sizeof(ocf) * len);
	if (fabs(frac) > 1e-2) {
		int index = floorf(frac * (float)ORC_DELAYFILTS);
		float h[40];
		taps_cx(T.delay[index], 20, h);
		orc_convolve_sv(in, len, 0, h, 20, 1, 1, 1, 0, 0, shift);
	} else {
		memcpy(shift, in, sizeof(ocf) * len);
	}
	if (whole < 0) {
		whole = -whole;
		int w = 0;
		for (int s = whole; s < len; s++) shift[w++] = shift[s];
		for (; w < len; w++) { shift[w].r = 0.0f; shift[w].i = 0.0f; }
	} else {
		int w = len - 1;
		for (int s = len - 1 - whole; s >= 0; s--) shift[w--] = shift[s];
		for (; w >= 0; w--) { shift[w].r = 0.0f; shift[w].i = 0.0f; }
	}
	memcpy(out, shift, sizeof(ocf) * len);
	free(shift);
	return len;
}

/* vectorSlicer :546-556 */
void orc_vector_slicer(float *dst, const float *src, size_t len)
{
	for (size_t i = 0; i < len; i++) {
		dst[i] = 0.5 * (src[i] + 1.0f);
		if (dst[i] > 1.0) dst[i] = 1.0;
		else if (dst[i] < 0.0) dst[i] = 0.0;
	}
}

/* demodCommon :2030-2048: timing recovery and one-tap channel correction; at 4 sps the result is decimated to 1 sps
 * (156 samples), at 1 sps the delayed and scaled burst is the result (blen samples).  Returns the length of dec. */
#define ORC_DEMOD_MAX 160
static int demod_common(const ocf *burst, int blen, int sps, const orc_ebp *ebp, ocf *dec)
{
	ocf *d = (ocf *)malloc(sizeof(ocf) * blen);
	orc_delay_vector(burst, blen, -ebp->toa * (float)sps, d);
	ocf one = { 1.0f, 0.0f };
	ocf scale = cdiv(one, ebp->amp);
	for (int i = 0; i < blen; i++)
		d[i] = cmul(d[i], scale);
	int n = ORC_DEC_LEN;
	if (sps == 1) {
		n = blen;
		memcpy(dec, d, sizeof(ocf) * blen);
	} else {
		orc_downsample_burst(d, blen, dec);
	}
	free(d);
	return n;
}

/* demodGmskBurst :2055-2072 */
static int demod_gmsk(const ocf *burst, int blen, int sps, const orc_ebp *ebp, float *soft)
{
	ocf dec[ORC_DEMOD_MAX];
	const int n = demod_common(burst, blen, sps, ebp, dec);
	for (int i = 0; i < n; i++)
		soft[i] = cmul(T.rrot1[i], dec[i]).r;
	return n;
}

/* demodEdgeBurst :2105-2128 with derotateEdgeBurst :691-711, computeEdgeCI :2074-2093,
 * softSliceEdgeBurst :1962-2006, rotateBurst2 :582-588 */
static int demod_edge(const ocf *burst, int blen, int sps, orc_ebp *ebp, float *soft)
{
	ocf dec[ORC_DEMOD_MAX], eq[ORC_DEMOD_MAX], rot[ORC_DEMOD_MAX];
	const int dlen = demod_common(burst, blen, sps, ebp, dec); /* 156 at 4 sps, the burst's own length at 1 sps */
	float h[10];
	taps_cx(T.c0_inv, 5, h);
	orc_convolve_sv(dec, dlen, 64, h, 5, 1, 0, 1, 0, 0, eq);
	for (int i = 0; i < dlen; i++) {
		float phase = (float)(i % 16) * 3.0f * M_PI / 8.0f;
		ocf r = { cosf(phase), -sinf(phase) };
		rot[i] = cmul(eq[i], r);
	}
	/* computeEdgeCI */
	float err_pwr = 0.0f;
	float step = 2.0f * M_PI_F / 8.0f;
	for (int i = 8; i < dlen - 8; i++) {
		ocf sym = rot[i];
		float phase = step * roundf(atan2f(sym.i, sym.r) / step);
		ocf ideal = { cosf(phase), sinf(phase) };
		ocf err = { ideal.r - sym.r, ideal.i - sym.i };
		err_pwr += norm2(err);
	}
	ebp->ci = 3.0103f * log2f(1.0f * (dlen - 16) / err_pwr);
	/* softSliceEdgeBurst */
	const int nsyms = 148;
	ocf r1 = { (float)cos(-M_PI / 8.0), (float)sin(-M_PI / 8.0) };
	for (int i = 0; i < dlen; i++) rot[i] = cmul(rot[i], r1);
	for (int i = 0; i < nsyms; i++) { soft[3 * i] = -rot[i].i; soft[3 * i + 1] = rot[i].r; }
	for (int i = 0; i < dlen; i++) { rot[i].r = fabsf(rot[i].r); rot[i].i = fabsf(rot[i].i); }
	ocf r2 = { (float)cos(-M_PI / 4.0), (float)sin(-M_PI / 4.0) };
	for (int i = 0; i < dlen; i++) rot[i] = cmul(rot[i], r2);
	for (int i = 0; i < nsyms; i++) soft[3 * i + 2] = -rot[i].i;
	return nsyms * 3;
}

/* demodAnyBurst :2130-2137 */
int orc_demod_any_burst(const ocf *burst, int blen, int type, int sps, orc_ebp *ebp, float *soft)
{
	orc_setup();
	if ((sps != 1 && sps != 4) || (sps == 1 && (blen > ORC_DEMOD_MAX || blen < 148)))
		return -1;
	if (type == ORC_EDGE)
		return demod_edge(burst, blen, sps, ebp, soft);
	return demod_gmsk(burst, blen, sps, ebp, soft);
}

/* ---- correlation sequence generation (sigProcLib.cpp:1227-1527) ---- */
static void gen_corr_seq(orc_corrseq *cs, const uint8_t *full, int nfull, const uint8_t *mid, int nmid,
			 int is_midamble, double toa_off)
{
	ocf shaped[64], ref[64], ac[64];
	memset(shaped, 0, sizeof(shaped));
	rotate_burst(mid, nmid, 0, 1, ref, 64);
	modulate_basic(full, nfull, 0, 1, shaped, 64);
	if (is_midamble) {
		ocf m1 = { -1.0f, 0.0f }, j = { 0.0f, 1.0f };
		for (int i = 0; i < nmid; i++) ref[i] = cmul(ref[i], m1);
		for (int i = 0; i < nfull; i++) shaped[i] = cmul(shaped[i], j);
	}
	for (int i = 0; i < nmid; i++) ref[i].i = -ref[i].i; /* conjugateVector */
	memcpy(cs->seq, ref, sizeof(ocf) * nmid);
	cs->len = nmid;
	orc_convolve_sv(shaped, nfull, 0, (const float *)ref, nmid, 0, 1, 1, 0, 0, ac);
	float toa;
	cs->gain = peak_detect(ac, nfull, &toa, NULL);
	cs->toa = toa - toa_off;
}

static void gen_edge_midamble(orc_corrseq *cs, int tsc)
{
	uint8_t bits[78];
	str_bits(EDGE_TSC_STR[tsc], bits);
	ocf m[16];
	orc_modulate_edge(bits + 15, 48, 1, 1, m, 16);
	for (int i = 0; i < 16; i++) { cs->seq[i].r = m[i].r; cs->seq[i].i = -m[i].i; }
	cs->len = 16;
	ocf g = { -19.6432, 19.5006 };
	float d = 1.18; /* Complex / Real (Complex.h:76): divisor narrowed to float */
	cs->gain.r = g.r / d; cs->gain.i = g.i / d;
	cs->toa = 0;
}

const orc_tables *orc_setup(void)
{
	if (g_ready || t_in_setup)
		return &T;
	pthread_mutex_lock(&g_lock);
	if (!g_ready) {
		t_in_setup = 1;
		uint8_t b[80];
		gen_sinc_table();
		gen_rot_tables();
		gen_pulses();
		for (int i = 0; i < 3; i++) {
			int n = str_bits(RACH_STR[i], b);
			gen_corr_seq(&T.rach[i], b, n, b, 40, 0, 20.5);
		}
		{
			int n = str_bits(SCH_STR, b);
			gen_corr_seq(&T.sch, b, n, b, n, 0, 32.5);
			n = str_bits(DUMMY_TSC_STR, b);
			gen_corr_seq(&T.dummy, b, n, b + 5, 16, 1, 13.5);
		}
		for (int t = 0; t < 8; t++) {
			int n = str_bits(TSC_STR[t], b);
			gen_corr_seq(&T.midamble[t], b, n, b + 5, 16, 1, 13.5);
			gen_edge_midamble(&T.edge_midamble[t], t);
		}
		gen_delay_filters();
		orc_resampler_proto(1, 4, 16, 1.0f, T.dnsamp);
		t_in_setup = 0;
		__sync_synchronize();
		g_ready = 1;
	}
	pthread_mutex_unlock(&g_lock);
	return &T;
}

int orc_get_table(const char *name, int idx, float *out, int max_floats)
{
	orc_setup();
	const float *src = NULL;
	int n = 0;
	float tmp[64];
	const orc_corrseq *cs = NULL;
	int meta = strstr(name, "_meta") != NULL;
#define IS(s) (!strcmp(name, s))
	if (IS("sinc")) { src = T.sinc; n = ORC_SINC_TABLESIZE + 1; }
	else if (IS("rot4")) { src = (float *)T.rot4; n = 1250; }
	else if (IS("rrot4")) { src = (float *)T.rrot4; n = 1250; }
	else if (IS("rot1")) { src = (float *)T.rot1; n = 314; }
	else if (IS("rrot1")) { src = (float *)T.rrot1; n = 314; }
	else if (IS("delay")) { taps_cx(T.delay[idx], 20, tmp); src = tmp; n = 40; }
	else if (IS("pulse4_c0")) { taps_cx(T.pulse4_c0, 16, tmp); src = tmp; n = 32; }
	else if (IS("pulse4_c1")) { taps_cx(T.pulse4_c1, 8, tmp); src = tmp; n = 16; }
	else if (IS("pulse4_c0inv")) { taps_cx(T.c0_inv, 5, tmp); src = tmp; n = 10; }
	else if (IS("pulse1_c0")) { taps_cx(T.pulse1_c0, 4, tmp); src = tmp; n = 8; }
	else if (IS("dnsamp")) { taps_cx(T.dnsamp, 16, tmp); src = tmp; n = 32; }
	else if (!strncmp(name, "midamble", 8)) cs = &T.midamble[idx];
	else if (!strncmp(name, "edge_midamble", 13)) cs = &T.edge_midamble[idx];
	else if (!strncmp(name, "rach", 4)) cs = &T.rach[idx];
	else if (!strncmp(name, "sch", 3)) cs = &T.sch;
	else if (!strncmp(name, "dummy", 5)) cs = &T.dummy;
	else if (IS("psk8")) { src = (const float *)psk8_table; n = 16; }
	else return -1;
#undef IS
	if (cs) {
		if (meta) { tmp[0] = cs->gain.r; tmp[1] = cs->gain.i; tmp[2] = cs->toa; src = tmp; n = 3; }
		else { src = (const float *)cs->seq; n = 2 * cs->len; }
	}
	if (n > max_floats)
		return -n;
	memcpy(out, src, sizeof(float) * n);
	return n;
}

/* convert (arch/x86/convert_sse_3.c:29-103 semantics: cvtps2dq = round-to-nearest-even under the
 * default MXCSR, packs_epi32 = signed saturation) */
void orc_convert_float_short(short *out, const float *in, float scale, int len)
{
	for (int i = 0; i < len; i++) {
		float v = in[i] * scale;
		long r;
		if (!(v == v) || v >= 2147483648.0f || v < -2147483648.0f) r = -2147483648L; /* "integer indefinite" */
		else r = lrintf(v);
		if (r > 32767) r = 32767;
		if (r < -32768) r = -32768;
		out[i] = (short)r;
	}
}

/* base_convert_float_short (arch/common/convert_base.c:20-25): `short = float * scale`, which gcc compiles for x86-64 as
 * cvttss2si (truncation; 0x80000000 for NaN / out of int32 range) followed by keeping the low 16 bits */
void orc_base_convert_float_short(short *out, const float *in, float scale, int len)
{
	for (int i = 0; i < len; i++) {
		float v = in[i] * scale;
		int r;
		if (!(v == v) || v >= 2147483648.0f || v < -2147483648.0f) r = (int)0x80000000;
		else r = (int)v;
		out[i] = (short)(r & 0xffff);
	}
}

/* convert_float_short as an SSE3 host dispatches it (arch/x86/convert.c:63-71, convert_sse_3.c:38-47): whole groups of
 * eight through the SSE routine, the len % 8 tail through the scalar loop */
void orc_convert_float_short_x86(short *out, const float *in, float scale, int len)
{
	const int body = len / 8 * 8;
	orc_convert_float_short(out, in, scale, body);
	orc_base_convert_float_short(out + body, in + body, scale, len - body);
}

void orc_convert_short_float(float *out, const short *in, int len)
{
	for (int i = 0; i < len; i++)
		out[i] = (float)in[i];
}
