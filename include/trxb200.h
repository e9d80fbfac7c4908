/*
 * trxb200.h — C ABI of the B200-native batched burst-DSP library (libtrxb200.so).
 *
 * This is the drop-in boundary for the Transceiver52M burst signal-processing hot path of
 * osmo-trx.  The reference has no FFI for this path: sigProcLib / Resampler / Channelizer /
 * Synthesis / grgsm_vitac are statically linked C++ called once per burst
 * (Transceiver.cpp:392-396,768-786; radioInterfaceMulti.cpp:264-343).  The entry points below are
 * what a binding for that path has to offer: the same operations, batched over N independent
 * bursts (ARFCN x timeslot x frame), plain pointers and sizes, no C++ or torch types.
 * Citations are file:line in the reference tree.  osmo_trx_b200/host/ re-creates the reference's
 * per-burst C++ API (sigProcLib.h etc.) on top of this ABI; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - complex samples are interleaved float32 (re,im) = `Complex<float>` (Complex.h:29-33);
 *     `stride` arguments are in complex samples.
 *   - `*_batch` functions take DEVICE pointers, enqueue on the context's stream and return
 *     without synchronising; `*_host` functions take HOST pointers and include the copies.
 *   - return value: 0 (TRXB200_OK) or a negative TRXB200_E* code.  Per-burst results use the
 *     reference's own conventions (rc > 0 CorrType, 0 no burst, < 0 -SignalError,
 *     sigProcLib.h:29-45,127-129).
 *   - there is no CPU fallback: every entry point fails with TRXB200_ENODEV/TRXB200_ECUDA when
 *     no sm_100 device is usable.
 */
#ifndef TRXB200_H
#define TRXB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRXB200_ABI_VERSION 1

/* error codes */
#define TRXB200_OK 0
#define TRXB200_EINVAL (-1)   /* bad argument */
#define TRXB200_ENODEV (-2)   /* no usable CUDA device */
#define TRXB200_ECUDA (-3)    /* CUDA runtime error (see trxb200_last_error) */
#define TRXB200_ENOMEM (-4)
#define TRXB200_EBOUNDS (-5)  /* convolve bounds_check failure (convolve_base.c:114-131) */

/* CorrType (sigProcLib.h:29-37) and SignalError (sigProcLib.h:39-45) */
enum trxb200_corr_type { TRXB200_OFF = 0, TRXB200_TSC = 1, TRXB200_EXT_RACH = 2, TRXB200_RACH = 3,
			 TRXB200_SCH = 4, TRXB200_EDGE = 5, TRXB200_IDLE = 6 };
enum trxb200_sigerr { TRXB200_SIGERR_NONE = 0, TRXB200_SIGERR_BOUNDS = 1, TRXB200_SIGERR_CLIP = 2,
		      TRXB200_SIGERR_UNSUPPORTED = 3, TRXB200_SIGERR_INTERNAL = 4 };

/* per-burst edge-case flags (north_star parity rule: such bursts are counted and reported) */
#define TRXB200_FLAG_THRESH_EDGE 1 /* |peak-to-average - threshold| < 1e-5 */
#define TRXB200_FLAG_BISECT_TIE 2  /* early/late powers within 4 ulp in some TOA bisection step */
#define TRXB200_FLAG_CLIP 4	   /* undetected burst with max(|I|,|Q|) > 30000, i.e. rc == -SIGERR_CLIP
				    * (sigProcLib.cpp:49,1746-1764; what Transceiver.cpp:772 counts as rx_clipping) */

#define TRXB200_BURST_LEN 625	/* samples per slot at 4 sps (radioInterface.cpp:257-258) */
#define TRXB200_GMSK_SOFT 156	/* demodGmskBurst output length (sigProcLib.cpp:2055-2072) */
#define TRXB200_NB_BITS 148
#define TRXB200_EDGE_SOFT 444	/* demodEdgeBurst output length (sigProcLib.cpp:1962-2006) */

typedef struct trxb200_ctx trxb200_ctx;

/* ---- life cycle: replaces convolve_init()+convert_init() (osmo-trx.cpp:648-649),
 *      sigProcLibSetup() (sigProcLib.cpp:2139-2172) and initvita() (grgsm_vitac.cpp:51-80) ---- */
int trxb200_init(int device, trxb200_ctx **out);
void trxb200_destroy(trxb200_ctx *ctx); /* sigProcLibDestroy() sigProcLib.cpp:137-176 */
int trxb200_abi_version(void);
const char *trxb200_last_error(trxb200_ctx *ctx);
/* A new context launches on its own non-blocking stream.  trxb200_set_stream() makes all subsequent
 * calls use the given cudaStream_t instead (NULL = the CUDA default stream); trxb200_use_own_stream()
 * switches back.
 * One context = one unit of device scratch (correlation intermediates, pull scratch, filterbank history), so the work
 * of a context never overlaps itself: switching streams orders the new stream behind everything the context enqueued
 * on the previous one (event record + wait, no host synchronisation).  Pipelines that want two batches in flight at
 * once use two contexts.  A context is bound to the device it was created on; every entry point selects that device
 * for the duration of the call and restores the caller's current device.  A context is not thread-safe: one caller
 * thread at a time (the reference's per-channel RX/TX threads each own one, Transceiver.cpp:308-322). */
int trxb200_set_stream(trxb200_ctx *ctx, void *cuda_stream);
int trxb200_use_own_stream(trxb200_ctx *ctx);
void *trxb200_get_stream(trxb200_ctx *ctx);
int trxb200_sync(trxb200_ctx *ctx);
int trxb200_device(trxb200_ctx *ctx);
int trxb200_sm_count(trxb200_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t trxb200_launch_count(trxb200_ctx *ctx);
/* Device memory helpers so that a binding needs no CUDA runtime of its own (cgo / JNI / ctypes / the C++ mirror in
 * osmo_trx_b200/host).  Copies are ordered on the context's stream; copy_to_host also waits for the stream, i.e.
 * for every kernel enqueued before it. */
int trxb200_dev_alloc(trxb200_ctx *ctx, size_t bytes, void **out);
int trxb200_dev_free(trxb200_ctx *ctx, void *p);
int trxb200_copy_to_device(trxb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int trxb200_copy_to_host(trxb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int trxb200_memset_device(trxb200_ctx *ctx, void *dst, int value, size_t bytes);
/* per-kernel device time of the detect / demod kernels launched between begin and end, measured with CUDA
 * events on the launching stream (used by bench.py for the roofline line; adds event overhead, so it is
 * not left on during throughput runs).  end() synchronises and writes "name:total_ms:launches;..." */
int trxb200_profile_begin(trxb200_ctx *ctx);
int trxb200_profile_end(trxb200_ctx *ctx, char *out, int cap);
/* host copy of a setup table for bit-exact checks against sigProcLib.cpp:52-135 statics.
 * names: sinc rot4 rrot4 rot1 rrot1 delay pulse4_c0 pulse4_c1 pulse4_c0inv pulse1_c0 dnsamp psk8
 *        midamble edge_midamble rach sch dummy (+ "_meta" = gain.re, gain.im, toa)
 *        vitac_norm vitac_access vitac_sch.  Returns the number of floats written or < 0. */
int trxb200_get_table(trxb200_ctx *ctx, const char *name, int idx, float *out, int max_floats);

/* ---- modulators: modulateBurst(bits, guard, sps=4) -> modulateBurstLaurent (sigProcLib.cpp:595-670,
 *      970-979) and modulateEdgeBurst(bits, sps=4) (sigProcLib.cpp:713-763,917-936).
 *      bits: u8[n][bits_stride], one bit per byte, only bit 0 is read (BitVector semantics,
 *      BitVector.cpp:59-68); out: complex[n][out_stride], 625 samples written per burst. ---- */
int trxb200_modulate_gmsk_batch(trxb200_ctx *ctx, const uint8_t *bits, int nbits, int bits_stride, int n,
				float *out, int out_stride);
int trxb200_modulate_edge_batch(trxb200_ctx *ctx, const uint8_t *bits, int nbits, int bits_stride, int n,
				float *out, int out_stride);

/* modulateBurst(bits, guard, sps, emptyPulse) outside the 4-sps Laurent case and modulateEdgeBurst(bits, sps, true)
 * (sigProcLib.cpp:558-580,672-689,938-979) - the forms the reference uses to build its correlation references and to
 * transmit at 1 sps.  mode 0: modulateBurstBasic (sps = 1: one Gaussian pulse, GSMPulse1); mode 1: rotateBurst
 * (emptyPulse: NRZ impulses times the GMSK rotator, sps 1 or 4); mode 2: rotateEdgeBurst (Gray-mapped 8-PSK symbols
 * times e^{j i 3pi/8} at stride sps; guard must be 0).  Writes sps * (symbols + guard) samples per burst. */
int trxb200_modulate_basic_batch(trxb200_ctx *ctx, const uint8_t *bits, int nbits, int bits_stride, int n, int guard, int sps,
				 int mode, float *out, int out_stride);

/* ---- detection: detectAnyBurst (sigProcLib.cpp:1926-1957) at sps = 4 for every burst b:
 *      rc[b] = detectAnyBurst(bursts[b], tsc[b], thresh, 4, type[b], max_toa[b], &ebp)
 *      amp[b] (re,im), toa[b], tsc_out[b], ci[b] = ebp fields (sigProcLib.h:113-118).
 *      max_toa_bound >= every max_toa[b] (sizes on-chip buffers; larger values give -SIGERR_BOUNDS).
 *      flags may be NULL. ---- */
int trxb200_detect_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, const uint8_t *type,
			 const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh, int32_t *rc,
			 float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags);

/* Sizing hints for the detection kernels.  max_seq_len: the longest sync sequence following batches may
 * correlate against - 16 when only TSC / EDGE / IDLE bursts are submitted, 40 (default) when RACH / EXT_RACH
 * may occur (generateRACHSequence uses 40 symbols, sigProcLib.cpp:1420).  max_attempts: detection rounds
 * scheduled per batch - 1 when only TSC / RACH / IDLE bursts are submitted, 2 with EDGE (detectAnyBurst retries
 * an EDGE miss as TSC, sigProcLib.cpp:1933-1941), 3 (default) with EXT_RACH (three sync sequences, :1793-1800).
 * A burst that needs more than the configured sizes is reported as -SIGERR_BOUNDS, never processed wrongly. */
int trxb200_detect_config(trxb200_ctx *ctx, int max_seq_len, int max_attempts);

/* ---- detectSCHBurst(burst, thresh, 4, SCH_DETECT_FULL, &ebp) (sigProcLib.cpp:1805-1861), the single-burst search
 *      of the MS side: 64-symbol synchronisation sequence, 156 correlation outputs (head = 105, tail = 51 symbols).
 *      rc: 1 detected / 0 not (detectBurst's value, as the reference returns it); amp, toa (relative to the expected
 *      position, head subtracted), ci as in estim_burst_params; flags may be NULL.  The NARROW state reads past its
 *      8-sample decimated vector in the reference and the BUFFER state searches a 12-frame capture: neither is offered. ---- */
int trxb200_detect_sch_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, float thresh, int32_t *rc, float *amp,
			     float *toa, float *ci, uint8_t *flags);

/* detectSCHBurst(burst, thresh, 4, SCH_DETECT_BUFFER, &ebp) (sigProcLib.cpp:1805-1861): the first SCH acquisition over a
 * capture of in_len samples per row (12 frames = 60000 in ms/ms_rx_lower.cpp:213-219; in_len a multiple of 4, at most
 * 4 * 16384 as Resampler::rotate allows).  The capture is decimated, correlated with the 64-symbol sequence at every
 * symbol position from 0, and detectBurst's peak logic runs over the whole vector; toa[b] is the burst's first symbol
 * inside the capture (peak - (3 + 39 + 64)).  rc 1 / 0 / -1 as the reference; amp and toa are 0 when nothing is found. */
int trxb200_detect_sch_buffer_batch(trxb200_ctx *ctx, const float *bufs, int stride, int in_len, int n, float thresh,
				    int32_t *rc, float *amp, float *toa, float *ci, uint8_t *flags);

/* ---- demodulation: demodAnyBurst (sigProcLib.cpp:2130-2137) for every burst with rc[b] > 0
 *      (rc[b] is the CorrType returned by detection).  soft: f32[n][soft_stride]; GMSK bursts get
 *      `n_gmsk_soft` values (148 = what Transceiver.cpp:799-803 consumes, or 156 = the full
 *      SoftVector), EDGE bursts 444.  ci is updated for EDGE bursts (computeEdgeCI, :2074-2093,2118).
 *      Bursts with rc[b] <= 0 leave their soft row untouched. ---- */
int trxb200_demod_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, const int32_t *rc,
			const float *amp, const float *toa, float *ci, float *soft, int soft_stride,
			int n_gmsk_soft);

/* ---- the receive path at ONE sample per symbol (rx_sps = 1): detectAnyBurst / demodAnyBurst called with sps == 1
 *      (sigProcLib.cpp:1659-1662: the correlator reads the burst itself, no decimation; :2038-2042: the burst delayed by
 *      -toa and scaled by 1 / amp is the demodulator's 1-sps vector).  bursts: complex[n][stride], blen samples per
 *      burst (a slot is 156 or 157 symbols; 148 <= blen <= 160).  Outputs as the 4-sps calls; GMSK bursts yield blen
 *      soft values (n_gmsk_soft <= blen of them are written), 8-PSK bursts 444, and computeEdgeCI runs over blen - 16
 *      symbols.  Outside every BASELINE configuration: built for coverage of the reference's interface, not tuned. ---- */
int trxb200_detect_sps1_batch(trxb200_ctx *ctx, const float *bursts, int stride, int blen, int n, const uint8_t *type,
			      const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh, int32_t *rc,
			      float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags);
int trxb200_demod_sps1_batch(trxb200_ctx *ctx, const float *bursts, int stride, int blen, int n, const int32_t *rc,
			     const float *amp, const float *toa, float *ci, float *soft, int soft_stride, int n_gmsk_soft);

/* ---- fused detect + demod, the Transceiver::pullRadioVector sequence (Transceiver.cpp:768,786) ---- */
int trxb200_detect_demod_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, const uint8_t *type,
			       const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh,
			       int32_t *rc, float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags,
			       float *soft, int soft_stride, int n_gmsk_soft);

/* same, HOST buffers: chunks the batch and overlaps H2D / kernels / D2H on three internal streams; returns after
 * everything has landed in the host outputs.  Per chunk one copy of the samples and one of the packed per-burst
 * inputs go down, one copy of the soft rows and one of the packed per-burst results come back.  Page-locked caller
 * buffers (cudaHostAlloc / cudaHostRegister) are read and written in place; pageable ones are staged through internal
 * pinned buffers (one extra host memcpy per chunk) so that the copies stay asynchronous either way. */
int trxb200_detect_demod_host(trxb200_ctx *ctx, const float *bursts, int stride, int n, const uint8_t *type,
			      const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh,
			      int32_t *rc, float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags,
			      float *soft, int soft_stride, int n_gmsk_soft);

/* ---- the receive chain around the hot path: int16 slot in, TRXD uplink datagram out ----
 * Per slot b, the DSP statements of Transceiver::pullRadioVector (Transceiver.cpp:665-815) with what feeds and
 * follows them:
 *   convert_short_float of the radio's int16 I/Q (RadioInterface::pullBuffer, radioInterface.cpp:345-349;
 *   arch/x86/convert.c:37-79); slot b is the 625 samples at iq + b*stride (radioInterface.cpp:272-291 slices the
 *   stream into consecutive 625-sample slots at 4 sps, i.e. stride 625 on a contiguous stream);
 *   type OFF: nothing is computed or emitted (:713-716);
 *   energy[b] = energyDetect(burst, 20*sps) (:725), RSSI = 20*log10(rx_full_scale / sqrt(energy)) + rssi_offset (:742-751);
 *   type IDLE: idle indication (:754); otherwise rc[b] = detectAnyBurst(...) (:768), rc <= 0: idle indication,
 *   -SIGERR_CLIP reported in rc (:769-782); rc > 0: demodAnyBurst (:786) and vectorSlicer (:803);
 *   the datagram trxd_send_burst_ind_v0 / _v1 would write (proto_trxd.c:69-117): header (tn, fn, rssi, toa*256,
 *   and for v1 idle/modulation/tsc/ci in cB) followed by 148 (GMSK) or 444 (8-PSK) soft bits normalised to 0..255
 *   (v0: two more bytes, both written as 0; v0 emits nothing for idle slots).
 * pkt: u8[n][pkt_stride], pkt_len[b] = datagram length (0: nothing to send).  pkt_stride >= 159 (v1) / 158 (v0);
 * rows shorter than 455 / 454 cannot hold an 8-PSK burst: such a burst gets pkt_len 0 and TRXB200_FLAG_PKT_TRUNC.
 * amp, toa, ci, tsc_out (the estim_burst_params fields) and flags are optional (NULL).  iq must be 4-byte aligned.
 * rx_full_scale = RadioInterface::fullScaleInputValue(), rssi_offset = Transceiver::rssiOffset(chan).
 * The noise-level average over IDLE slots (mNoises, :744-748) is caller state: feed it from energy[]. ---- */
typedef struct trxb200_pull_args {
	const int16_t *iq;
	int stride; /* complex samples between consecutive slots, >= 625 */
	int n;
	const uint8_t *type, *tsc;
	const uint16_t *max_toa;
	const uint32_t *fn; /* TDMA frame number per slot */
	const uint8_t *tn;  /* timeslot number per slot */
	int max_toa_bound;
	float thresh;
	double rx_full_scale, rssi_offset;
	int trxd_version; /* 0 or 1 */
	int32_t *rc;
	float *energy;
	uint8_t *pkt;
	int pkt_stride;
	uint16_t *pkt_len;
	uint8_t *flags;	 /* optional */
	float *amp, *toa, *ci; /* optional */
	uint8_t *tsc_out;      /* optional */
} trxb200_pull_args;
#define TRXB200_FLAG_PKT_TRUNC 8
#define TRXB200_TRXD_V1_HDR 11
#define TRXB200_TRXD_V0_HDR 8
/* DEVICE pointers; enqueues on the context's stream */
int trxb200_pull_batch(trxb200_ctx *ctx, const trxb200_pull_args *args);
/* HOST pointers; H2D / kernels / D2H pipelined over three internal streams, two copies down and two back per chunk
 * (slots + packed per-slot inputs; datagram rows + packed per-slot results); pageable caller memory is staged through
 * internal pinned buffers, page-locked memory is used in place (this is the e2e path of bench.py) */
int trxb200_pull_host(trxb200_ctx *ctx, const trxb200_pull_args *args);

/* ---- burst-type scheduler: Transceiver::expectedCorrType (Transceiver.cpp:513-601) and the search window
 *      pullRadioVector derives from it (:757-758: max_toa = RACH / EXT_RACH ? mMaxExpectedDelayAB : mMaxExpectedDelayNB).
 *      Per slot (fn, tn, chan) -> CorrType + max_toa, ready to be handed to trxb200_pull_batch / detect_batch as the
 *      `type` and `max_toa` arrays.  All pointers are device pointers.
 *        chan_type  u8[n_chan][8]  ChannelCombination per timeslot (Transceiver.h:131-148: FILL 0, I 1 .. XIII 13,
 *                                  NONE 14, LOOPBACK 15), i.e. TransceiverState::chanType
 *        handover   u8[8]          per timeslot, bit s set = mHandover[tn][s] (Transceiver.h:225)
 *        chan       u16[n] or NULL (all slots belong to channel 0); a channel index >= n_chan gives OFF
 *        max_toa    u16[n] or NULL ---- */
typedef struct trxb200_sched_cfg {
	int n_chan;
	const uint8_t *chan_type;
	const uint8_t *handover;
	int ext_rach;	/* cfg->ext_rach */
	int egprs;	/* cfg->egprs */
	int max_toa_nb; /* mMaxExpectedDelayNB */
	int max_toa_ab; /* mMaxExpectedDelayAB */
} trxb200_sched_cfg;
int trxb200_expected_corr_type_batch(trxb200_ctx *ctx, const trxb200_sched_cfg *cfg, const uint32_t *fn, const uint8_t *tn,
				     const uint16_t *chan, int n, uint8_t *type, uint16_t *max_toa);

/* ---- small per-burst helpers of sigProcLib.h used around detection ---- */
/* energyDetect(burst, window) (sigProcLib.cpp:1573-1585): mean |x|^2 of `window` samples at stride 4 */
int trxb200_energy_detect_batch(trxb200_ctx *ctx, const float *bursts, int stride, int blen, int n,
				unsigned window, float *energy);
/* vectorSlicer (sigProcLib.cpp:546-556) over a flat array */
int trxb200_vector_slicer(trxb200_ctx *ctx, float *dst, const float *src, size_t len);
/* delayVector (sigProcLib.cpp:1046-1098): per-burst delay[b] in samples */
int trxb200_delay_vector_batch(trxb200_ctx *ctx, const float *in, int stride, int len, int n,
			       const float *delay, float *out, int out_stride);

/* ---- FIR stages: convolve_real / convolve_complex / base_convolve_* (arch/common/convolve.h:4-26),
 *      y[b][i] = sum_k x[b][i + start - (h_len-1) + k] * h[k], i < len, one shared h for the batch.
 *      x points at sample 0 of burst 0; samples before it (head-room) are read when start < h_len-1,
 *      exactly like the reference.  `base` != 0 selects the strictly sequential summation of
 *      arch/common/convolve_base.c (used for unaligned taps), otherwise the SSE3 summation tree of
 *      arch/x86/convolve_sse_3.c is reproduced.  Returns TRXB200_EBOUNDS where the reference
 *      returns -1. ---- */
int trxb200_convolve_real_batch(trxb200_ctx *ctx, const float *x, int x_len, int x_stride, const float *h,
				int h_len, float *y, int y_len, int y_stride, int start, int len, int n, int base);
int trxb200_convolve_complex_batch(trxb200_ctx *ctx, const float *x, int x_len, int x_stride, const float *h,
				   int h_len, float *y, int y_len, int y_stride, int start, int len, int n,
				   int base);

/* ---- int16 <-> float (arch/common/convert.h; SSE semantics: round-to-nearest-even + saturation) ---- */
int trxb200_convert_float_short(trxb200_ctx *ctx, int16_t *out, const float *in, float scale, size_t len);
/* mode 0: as above; 1: exactly what convert_float_short() of an SSE3 host returns for any length (arch/x86/convert.c:63-71:
 * SSE for whole groups of eight, the scalar loop - C truncation - for the len % 8 tail, convert_sse_3.c:38-47);
 * 2: base_convert_float_short (arch/common/convert_base.c:20-25: truncation, low 16 bits) */
int trxb200_convert_float_short_mode(trxb200_ctx *ctx, int16_t *out, const float *in, float scale, size_t len, int mode);
int trxb200_convert_short_float(trxb200_ctx *ctx, float *out, const int16_t *in, size_t len);

/* ---- grgsm_vitac MLSE (grgsm_vitac.h:65-82): per burst get_norm_chan_imp_resp / get_access_imp_resp
 *      -> clamp start to [clamp_lo, clamp_hi] (ms_upper.cpp:224-225 / Transceiver.cpp:631,635) ->
 *      detect_burst_nb / detect_burst_ab.  bufs: complex[n][stride]; the burst starts `offset`
 *      samples into each row so that negative starts stay addressable.  bits: i8[n][148 or 88]
 *      (+-127, pre-flipped, grgsm_vitac.cpp:101-102); cir may be NULL (else complex[n][20]).
 *      is_ab: 0 normal burst, 1 access burst, 2 SCH burst of a tracked cell (get_sch_chan_imp_resp :283-296 +
 *      detect_burst_nb, ms_rx_lower.cpp:173-177; tsc unused, corr_max is the search's maximum although the reference's
 *      function keeps it to itself). ---- */
int trxb200_vitac_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int n, int is_ab,
			const uint8_t *tsc, int max_delay, int clamp_lo, int clamp_hi, int8_t *bits,
			int32_t *start, float *corr_max, float *cir);

/* detect_burst_nb / detect_burst_ab (grgsm_vitac.cpp:105-123) on their own: the caller supplies the channel estimate
 * (complex[n][20], e.g. from an earlier trxb200_vitac_batch) and the burst start per burst; matched filter + Viterbi
 * only.  start_in is clamped to [clamp_lo, clamp_hi].  is_ab: 0 = 148 decisions, 1 = 88. */
int trxb200_vitac_detect_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int n, int is_ab,
			       const float *cir_in, const int32_t *start_in, int clamp_lo, int clamp_hi, int8_t *bits);
/* the five-argument detect_burst_nb / detect_burst_ab (grgsm_vitac.cpp:105-116): `ss` = the Viterbi detector's start state,
 * 0..15 (the reference indexes a 16-entry metric array with it unchecked; values outside are rejected here) */
int trxb200_vitac_detect_ss_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int n, int is_ab,
				  const float *cir_in, const int32_t *start_in, int clamp_lo, int clamp_hi, int start_state,
				  int8_t *bits);

/* First SCH acquisition of the MS side: get_sch_buffer_chan_imp_resp (grgsm_vitac/grgsm_vitac.cpp:298-309) over a capture of
 * `len` samples per row (12 frames in ms/ms_rx_lower.cpp:160-177: the 54 inner symbols of the extended training sequence are
 * correlated at every sample position 0 .. len - 512, the strongest 20-window wins) followed, when `bits` is given, by
 * detect_burst_nb at the position found (ms_rx_lower.cpp:173-177).  Row b holds its capture at bufs + 2*(b*stride + offset);
 * start[b] is the burst's first sample relative to that (position - 47 symbols; may be negative), cir complex[n][20],
 * corr_max[n].  The demodulated position is start limited to the row ([-offset, stride - offset - 592]), where the reference
 * reads outside its buffer. ---- */
int trxb200_vitac_sch_buffer_batch(trxb200_ctx *ctx, const float *bufs, int stride, int offset, int len, int n,
				   int8_t *bits, int32_t *start, float *corr_max, float *cir);

/* ---- Resampler (Resampler.h:31-61): rational p/q polyphase resampler, filt_len taps per path.
 *      rotate: in points at the first NEW input sample of each stream; `filt_len` samples of history
 *      precede it in memory (Resampler.cpp:131-150 reads before `in`).  n_streams independent streams
 *      (channels) are processed per call. ---- */
typedef struct trxb200_resampler trxb200_resampler;
int trxb200_resampler_create(trxb200_ctx *ctx, int p, int q, int filt_len, float bw, trxb200_resampler **out);
void trxb200_resampler_destroy(trxb200_resampler *r);
int trxb200_resampler_rotate(trxb200_resampler *r, const float *in, int in_len, int in_stride, float *out,
			     int out_len, int out_stride, int n_streams);
/* The same for streams of any length (no MAX_OUTPUT_LEN): because the path indices of Resampler.cpp:131-150 repeat every
 * p outputs / q inputs, one long call equals the reference's consecutive block calls whenever the blocks are whole periods,
 * each block's history being the tail of the one before - which is how radioInterfaceMulti.cpp:283-309 calls it. */
int trxb200_resampler_rotate_stream(trxb200_resampler *r, const float *in, int in_len, int in_stride, float *out,
				    int out_len, int out_stride, int n_streams);
int trxb200_resampler_taps(trxb200_resampler *r, int path, float *out_host);

/* ---- Channelizer / Synthesis (Channelizer.h:13-31, Synthesis.h:13-32, ChannelizerBase.cpp):
 *      M-channel critically sampled polyphase filterbanks; history is carried inside the object
 *      (ChannelizerBase hist[], Channelizer.cpp:87-88).  n_blocks consecutive blocks per call:
 *      channelizer in: complex[n_blocks][block_len*m] -> out: complex[m][n_blocks*block_len]
 *      synthesis   in: complex[m][n_blocks*block_len] -> out: complex[n_blocks][block_len*m] ---- */
typedef struct trxb200_filterbank trxb200_filterbank;
int trxb200_channelizer_create(trxb200_ctx *ctx, int m, int block_len, int h_len, trxb200_filterbank **out);
int trxb200_synthesis_create(trxb200_ctx *ctx, int m, int block_len, int h_len, trxb200_filterbank **out);
void trxb200_filterbank_destroy(trxb200_filterbank *fb);
int trxb200_filterbank_reset(trxb200_filterbank *fb);
int trxb200_channelizer_rotate(trxb200_filterbank *fb, const float *in, float *out, int n_blocks);
int trxb200_synthesis_rotate(trxb200_filterbank *fb, const float *in, float *out, int n_blocks);
/* The same channelizer step for any number of time samples (total_t >= h_len wideband rows of m samples; block_len only
 * names the reference's call size) with an output row pitch: channel c's samples land at out[c * out_stride + t], so the
 * caller can leave room in front of each row for the 16 samples of history Resampler::rotate reads before its input
 * (radioInterfaceMulti.cpp:283-309 keeps them in the channel's ring buffer).
 * trxb200_channelizer_prime sets the carried history from the last h_len rows of `prev_in` ([n_prev_t][m]) without
 * producing output: a rank that processes a time block in the middle of a stream re-reads that halo from the source
 * (Channelizer.cpp:87-88 is the state it replaces). */
int trxb200_channelizer_rotate_strided(trxb200_filterbank *fb, const float *in, long total_t, float *out, long out_stride);
int trxb200_channelizer_prime(trxb200_filterbank *fb, const float *prev_in, long n_prev_t);
int trxb200_filterbank_taps(trxb200_filterbank *fb, int branch, float *out_host);

#ifdef __cplusplus
}
#endif
#endif /* TRXB200_H */
