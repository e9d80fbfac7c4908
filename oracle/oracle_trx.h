/*
 * oracle/oracle_trx.h — CPU restatement of the osmo-trx Transceiver52M burst-DSP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a from-scratch plain-C
 * restatement of the reference algorithms (each function cites the reference
 * file:line it follows; paths relative to /root/reference).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (osmo_trx_b200/, include/) never does.
 *
 * Pinning: oracle_convolve.c is checked against the reference's own golden
 * vectors (tests/Transceiver52M/convolve_test_golden.h → tests/golden/); every
 * other function is checked bit-for-bit against oracle/_ref/libref_osmotrx.so
 * (the unmodified reference compiled from /root/reference) in tests/test_oracle_vs_ref.py
 * and against fixtures generated from it (the .npz files under tests/golden/).
 *
 * Arithmetic contract: IEEE float32, no FMA contraction (build with
 * -ffp-contract=off), operation order of the reference's SSE3 build.
 */
#ifndef ORACLE_TRX_H
#define ORACLE_TRX_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float r, i; } ocf; /* Complex<float>, Complex.h:29-33 */

/* CorrType sigProcLib.h:29-37, SignalError :39-45 */
enum { ORC_OFF = 0, ORC_TSC = 1, ORC_EXT_RACH = 2, ORC_RACH = 3, ORC_SCH = 4, ORC_EDGE = 5, ORC_IDLE = 6 };
enum { ORC_SIGERR_NONE = 0, ORC_SIGERR_BOUNDS, ORC_SIGERR_CLIP, ORC_SIGERR_UNSUPPORTED, ORC_SIGERR_INTERNAL };

#define ORC_BURST_LEN 625
#define ORC_DEC_LEN 156
#define ORC_SINC_TABLESIZE 1024
#define ORC_DELAYFILTS 64

typedef struct {
	ocf seq[64];
	int len;
	float toa;
	ocf gain;
} orc_corrseq;

/* everything sigProcLibSetup() builds (sigProcLib.cpp:2139-2172) */
typedef struct {
	float sinc[ORC_SINC_TABLESIZE + 1];
	ocf rot4[625], rrot4[625], rot1[157], rrot1[157];
	float pulse4_c0[16], pulse4_c1[8], pulse1_c0[4], c0_inv[5];
	float delay[ORC_DELAYFILTS][20];
	float dnsamp[16]; /* Resampler(1,4) partition 0, stored reversed */
	orc_corrseq midamble[8], edge_midamble[8], rach[3], sch, dummy;
} orc_tables;

const orc_tables *orc_setup(void);
int orc_get_table(const char *name, int idx, float *out, int max_floats);

/* arch/common/convolve.h:4-26 */
int orc_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
int orc_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
int orc_base_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);
int orc_base_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len);

/* static convolve() wrapper sigProcLib.cpp:297-398; span 0 START_ONLY, 1 NO_DELAY, 2 CUSTOM.
 * x_head = number of valid (zero) head-room samples before x (signalVector::getStart()). */
int orc_convolve_sv(const ocf *x, int x_len, int x_head, const float *h /*interleaved complex*/, int h_len, int h_real,
		    int h_aligned, int span, int start, int len, ocf *y);

/* modulators */
int orc_modulate_burst(const uint8_t *bits, int nbits, int guard, int sps, int empty, ocf *out, int max_cf);
int orc_modulate_edge(const uint8_t *bits, int nbits, int sps, int empty, ocf *out, int max_cf);
int orc_modulate_gmsk_batch(const uint8_t *bits, int nbits, int n, float *out, int nthreads);
int orc_modulate_edge_batch(const uint8_t *bits, int nbits, int n, float *out, int nthreads);

/* detection / demodulation */
typedef struct { ocf amp; float toa; uint8_t tsc; float ci; } orc_ebp; /* sigProcLib.h:113-118 */
int orc_detect_any_burst(const ocf *burst, int blen, unsigned tsc, float thresh, int sps, int type, unsigned max_toa,
			 orc_ebp *ebp, int *edge_flags);
int orc_demod_any_burst(const ocf *burst, int blen, int type, int sps, orc_ebp *ebp, float *soft /*>=444*/);
/* detectSCHBurst sigProcLib.cpp:1805-1861, SCH_DETECT_FULL */
int orc_detect_sch_buffer(const ocf *burst, int in_len, float thresh, orc_ebp *ebp, int *edge_flags);
int orc_detect_sch_burst(const ocf *burst, int blen, float thresh, int sps, orc_ebp *ebp, int *edge_flags);
int orc_detect_batch(const float *bursts, int stride, int blen, int n, const uint8_t *type, const uint8_t *tsc,
		     const uint16_t *max_toa, float thresh, int sps, int32_t *rc, float *amp, float *toa,
		     uint8_t *tsc_out, float *ci, uint8_t *flags, int nthreads);
int orc_demod_batch(const float *bursts, int stride, int blen, int n, const int32_t *rc, const float *amp,
		    const float *toa, float *ci, int sps, float *soft, int soft_stride, int32_t *nsoft, int nthreads);
int orc_detect_demod_batch(const float *bursts, int stride, int blen, int n, const uint8_t *type, const uint8_t *tsc,
			   const uint16_t *max_toa, float thresh, int sps, int32_t *rc, float *amp, float *toa,
			   uint8_t *tsc_out, float *ci, uint8_t *flags, float *soft, int soft_stride, int32_t *nsoft,
			   int nthreads);
float orc_energy_detect(const ocf *burst, int blen, unsigned window);
int orc_delay_vector(const ocf *in, int len, float delay, ocf *out);
void orc_vector_slicer(float *dst, const float *src, size_t len);
int orc_downsample_burst(const ocf *in, int blen, ocf *out /*156*/);
void orc_convert_float_short(short *out, const float *in, float scale, int len); /* SSE semantics */
void orc_base_convert_float_short(short *out, const float *in, float scale, int len); /* scalar: truncation, low 16 bits */
void orc_convert_float_short_x86(short *out, const float *in, float scale, int len); /* SSE body + scalar tail (len % 8) */
void orc_convert_short_float(float *out, const short *in, int len);

/* receive chain around the hot path (oracle_pull.c): int16 slot -> TRXD uplink datagram */
int orc_trxd_pack(int version, uint32_t fn, uint8_t tn, double rssi, double toa, int idle, int is_8psk, uint8_t tsc,
		  float ci, const float *soft01, int nbits, uint8_t *pkt);
int orc_pull_burst(const int16_t *iq, int type, unsigned tsc, unsigned max_toa, uint32_t fn, uint8_t tn, float thresh,
		   double full_scale, double rssi_offset, int version, uint8_t *pkt, int32_t *rc_out, float *energy_out,
		   orc_ebp *ebp_out, int *flags_out);
int orc_pull_batch(const int16_t *iq, int stride, int n, const uint8_t *type, const uint8_t *tsc, const uint16_t *max_toa,
		   const uint32_t *fn, const uint8_t *tn, float thresh, double full_scale, double rssi_offset, int version,
		   int32_t *rc, float *energy, uint8_t *pkt, int pkt_stride, uint16_t *pkt_len, uint8_t *flags, float *amp,
		   float *toa, float *ci, uint8_t *tsc_out, int nthreads);

/* flags written by detection (bit set) */
#define ORC_FLAG_THRESH_EDGE 1 /* |peakRatio - thresh| < 1e-5 */
#define ORC_FLAG_BISECT_TIE 2  /* early/late powers within 4 ulp at some bisection step */
#define ORC_FLAG_CLIP 4	       /* max |I|,|Q| > 30000 */
#define ORC_FLAG_PKT_TRUNC 8   /* pull path: datagram row too small for the burst (8-PSK), nothing emitted */

/* Resampler.h:31-61 */
typedef struct orc_resampler orc_resampler;
orc_resampler *orc_resampler_create(int p, int q, int filt_len, float bw);
void orc_resampler_destroy(orc_resampler *);
int orc_resampler_rotate(orc_resampler *, const float *in_with_hist, int hist, int in_len, float *out, int out_len);
int orc_resampler_taps(orc_resampler *, int path, float *out /*filt_len*/);

/* Channelizer.h / Synthesis.h */
typedef struct orc_chan orc_chan;
orc_chan *orc_channelizer_create(int m, int block_len, int h_len);
orc_chan *orc_synthesis_create(int m, int block_len, int h_len);
void orc_chan_destroy(orc_chan *);
int orc_channelizer_rotate(orc_chan *, const float *in, int m, int block_len, float *out);
int orc_synthesis_rotate(orc_chan *, const float *in, int m, int block_len, float *out);
int orc_chan_taps(orc_chan *, int branch, float *out /*h_len*/);

/* grgsm_vitac */
int orc_get_vitac_table(int which, int idx, float *out);
int orc_vitac_sch_buffer_batch(const float *bufs, int stride, int offset, int len, int n, int8_t *bits, int32_t *start_out,
			       float *corr_max, float *cir_out);
int orc_vitac_batch(const float *bufs, int stride, int offset, int n, int is_ab, const uint8_t *tsc, int max_delay,
		    int clamp_lo, int clamp_hi, int8_t *bits, int32_t *start_out, float *corr_max, float *cir_out,
		    int nthreads);
void orc_viterbi(const float *input, int n, const float *rhh, int start_state, float *out);

#ifdef __cplusplus
}
#endif
#endif
