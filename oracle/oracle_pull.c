/*
 * oracle/oracle_pull.c — CPU restatement of the receive chain AROUND the hot path:
 * int16 ingest, the pre-detection power measurement, detect + demodulate, soft-bit slicing and the TRXD
 * uplink datagram.  TEST INFRASTRUCTURE ONLY (see oracle_trx.h).
 *
 * Follows, per received slot:
 *   RadioInterface::pullBuffer      radioInterface.cpp:345-349  convert_short_float of the device's int16 I/Q
 *   Transceiver::pullRadioVector    Transceiver.cpp:665-815     energyDetect / RSSI / detectAnyBurst /
 *                                                               demodAnyBurst / vectorSlicer / idle handling
 *   trxd_send_burst_ind_v0 / _v1    proto_trxd.c:27-117         header fields + soft bits normalised to 0..255
 * Pinned bit-for-bit against the reference's own functions (oracle/_ref: ref_pull_batch drives
 * convert_short_float, energyDetect, detectAnyBurst, demodAnyBurst, vectorSlicer and the unmodified
 * trxd_send_burst_ind_v0/_v1, whose write() lands in a pipe) in tests/test_oracle_cpu.py.
 */
#include <math.h>
#include <string.h>
#include <pthread.h>
#include "oracle_trx.h"

/* `uint8_t = double` as gcc/x86-64 evaluates it (proto_trxd.c:43 `v0->rssi = bi->rssi`): cvttsd2si to a
 * 32-bit integer (0x80000000 for NaN / out of range), then the low byte. */
static uint8_t dbl_to_u8_x86(double v)
{
	int32_t t;
	if (!(v > -2147483649.0 && v < 2147483648.0)) t = INT32_MIN;
	else t = (int32_t)v;
	return (uint8_t)((uint32_t)t & 0xffu);
}

static int32_t dbl_to_i32_x86(double v)
{
	if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
	return (int32_t)v;
}

static void store16be(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)(v & 0xff); }
static void store32be(uint8_t *p, uint32_t v)
{
	p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

/* trxd_send_burst_ind_v0 proto_trxd.c:69-92 / _v1 :94-117.  Returns the datagram length (0: nothing is sent).
 * soft01: vectorSlicer output (0..1), nbits of them. */
int orc_trxd_pack(int version, uint32_t fn, uint8_t tn, double rssi, double toa, int idle, int is_8psk, uint8_t tsc,
		  float ci, const float *soft01, int nbits, uint8_t *pkt)
{
	int len = 0;
	if (version == 0 && idle)
		return 0; /* :73-74 v0 has no idle indications */
	/* trxd_fill_common :27-33: tn:3 | reserved:1 | version:4 (little-endian bit-field order, proto_trxd.h:46-49) */
	pkt[0] = (uint8_t)((tn & 7) | ((version & 15) << 4));
	store32be(pkt + 1, fn);
	/* trxd_fill_v0_specific :35-44 */
	pkt[5] = dbl_to_u8_x86(rssi);
	store16be(pkt + 6, (uint16_t)dbl_to_i32_x86(toa * 256.0 + 0.5));
	len = 8;
	if (version == 1) {
		/* trxd_fill_v1_specific :46-60: tsc:3 | modulation:4 | idle:1; ci in centiBels ((float)ci * 10 is a float product) */
		const int16_t ci_cb = (int16_t)dbl_to_i32_x86((double)(ci * 10) + 0.5);
		const int mod = is_8psk ? 4 : 0; /* TRXD_MODULATION_8PSK(0) / _GMSK(0), proto_trxd.h:76-77 */
		pkt[8] = (uint8_t)((tsc & 7) | (mod << 3) | ((idle ? 1 : 0) << 7));
		store16be(pkt + 9, (uint16_t)ci_cb);
		len = 11;
		if (idle)
			return len;
	}
	/* trxd_fill_burst_normalized255 :62-67 */
	for (int i = 0; i < nbits; i++)
		pkt[len + i] = (uint8_t)round(soft01[i] * 255.0);
	len += nbits;
	if (version == 0) {
		pkt[len] = 0;     /* :84-86 "uninitialised byte" (excluded from parity), then a NUL */
		pkt[len + 1] = 0;
		len += 2;
	}
	return len;
}

/* One slot through pullRadioVector's DSP.  iq: 625 complex int16.  Returns the datagram length. */
int orc_pull_burst(const int16_t *iq, int type, unsigned tsc, unsigned max_toa, uint32_t fn, uint8_t tn, float thresh,
		   double full_scale, double rssi_offset, int version, uint8_t *pkt, int32_t *rc_out, float *energy_out,
		   orc_ebp *ebp_out, int *flags_out)
{
	float x[2 * ORC_BURST_LEN];
	float soft[444], soft01[444];
	orc_ebp ebp;
	double rssi = 0.0, toa = 0.0;
	int rc = 0, idle = 1, nbits = 0, is_8psk = 0, fl = 0;
	uint8_t tsc_o = 0;
	float ci = 0.0f, avg, pw;
	memset(&ebp, 0, sizeof(ebp));
	*rc_out = 0; *energy_out = 0.0f; *ebp_out = ebp; *flags_out = 0;
	if (type == ORC_OFF)
		return 0; /* Transceiver.cpp:713-716: no processing at all, -ENOENT, nothing sent */
	orc_convert_short_float(x, iq, 2 * ORC_BURST_LEN); /* radioInterface.cpp:345-349 */
	pw = orc_energy_detect((const ocf *)x, ORC_BURST_LEN, 20 * 4); /* :725, one diversity path */
	*energy_out = pw;
	avg = sqrtf(pw / 1.0f);						   /* :742 */
	rssi = 20.0 * log10(full_scale / avg) + rssi_offset;		   /* :751 */
	if (type != ORC_IDLE) {						   /* :754-755 */
		rc = orc_detect_any_burst((const ocf *)x, ORC_BURST_LEN, tsc, thresh, 4, type, max_toa, &ebp, &fl); /* :768 */
		if (rc > 0) {
			int ns = orc_demod_any_burst((const ocf *)x, ORC_BURST_LEN, rc, 4, &ebp, soft); /* :786 */
			toa = ebp.toa; tsc_o = ebp.tsc; ci = ebp.ci;					 /* :789-791 */
			is_8psk = (ns == 444);								 /* :794-800 */
			nbits = is_8psk ? 444 : 148;
			orc_vector_slicer(soft01, soft, nbits);						 /* :803 */
			idle = 0;
		}
	}
	*rc_out = rc; *ebp_out = ebp; *flags_out = fl;
	return orc_trxd_pack(version, fn, tn, rssi, toa, idle, is_8psk, tsc_o, ci, soft01, nbits, pkt);
}

struct pull_job {
	const int16_t *iq; int stride, n; const uint8_t *type, *tsc; const uint16_t *max_toa; const uint32_t *fn; const uint8_t *tn;
	float thresh; double full_scale, rssi_offset; int version;
	int32_t *rc; float *energy; uint8_t *pkt; int pkt_stride; uint16_t *pkt_len; uint8_t *flags; float *amp, *toa, *ci; uint8_t *tsc_out;
	int lo, hi;
};

static void *pull_run(void *a)
{
	struct pull_job *j = (struct pull_job *)a;
	uint8_t buf[11 + 444 + 2];
	for (int b = j->lo; b < j->hi; b++) {
		orc_ebp ebp;
		int fl = 0;
		int len = orc_pull_burst(j->iq + (size_t)b * j->stride * 2, j->type[b], j->tsc[b], j->max_toa[b], j->fn[b], j->tn[b],
					 j->thresh, j->full_scale, j->rssi_offset, j->version, buf, &j->rc[b], &j->energy[b], &ebp, &fl);
		if (len > j->pkt_stride) { len = 0; fl |= 8; } /* row too small for an 8-PSK burst: flagged, nothing emitted */
		memcpy(j->pkt + (size_t)b * j->pkt_stride, buf, len);
		j->pkt_len[b] = (uint16_t)len;
		if (j->flags) j->flags[b] = (uint8_t)fl;
		if (j->amp) { j->amp[2 * b] = ebp.amp.r; j->amp[2 * b + 1] = ebp.amp.i; }
		if (j->toa) j->toa[b] = ebp.toa;
		if (j->ci) j->ci[b] = ebp.ci;
		if (j->tsc_out) j->tsc_out[b] = ebp.tsc;
	}
	return NULL;
}

/* Batch form with the array layout of trxb200_pull_batch (include/trxb200.h). */
int orc_pull_batch(const int16_t *iq, int stride, int n, const uint8_t *type, const uint8_t *tsc, const uint16_t *max_toa,
		   const uint32_t *fn, const uint8_t *tn, float thresh, double full_scale, double rssi_offset, int version,
		   int32_t *rc, float *energy, uint8_t *pkt, int pkt_stride, uint16_t *pkt_len, uint8_t *flags, float *amp,
		   float *toa, float *ci, uint8_t *tsc_out, int nthreads)
{
	pthread_t th[256];
	struct pull_job jobs[256];
	int nt = 0;
	orc_setup();
	if (nthreads > 256) nthreads = 256;
	if (nthreads < 1 || n < 2 * nthreads) nthreads = 1;
	const int per = (n + nthreads - 1) / nthreads;
	for (int t = 0; t < nthreads; t++) {
		struct pull_job j = { iq, stride, n, type, tsc, max_toa, fn, tn, thresh, full_scale, rssi_offset, version,
				      rc, energy, pkt, pkt_stride, pkt_len, flags, amp, toa, ci, tsc_out, t * per,
				      (t + 1) * per > n ? n : (t + 1) * per };
		if (j.lo >= j.hi) break;
		jobs[nt] = j;
		if (nthreads == 1) pull_run(&jobs[nt]);
		else pthread_create(&th[nt], NULL, pull_run, &jobs[nt]);
		nt++;
	}
	if (nthreads > 1)
		for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
	return n;
}
