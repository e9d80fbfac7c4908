/*
 * oracle/ref_capi.cpp — C-ABI driver around the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (osmo_trx_b200/, include/)
 * may include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs load the resulting
 * oracle/_ref/libref_osmotrx.so.
 *
 * The reference's hot path has no FFI of its own (statically linked C++,
 * SURVEY.md §8(b)), so this file is the thinnest possible batch loop over its
 * per-burst API.  It #includes sigProcLib.cpp so that the file-static tables
 * (sigProcLib.cpp:52-135) can be dumped for bit-for-bit table checks.
 * No reference source is copied: the TUs are compiled where they lie under
 * /root/reference (see oracle/Makefile).
 */
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <functional>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <complex>
#include <algorithm>

#include "sigProcLib.cpp" /* reference TU, pulled in for its statics */

#include "Channelizer.h"
#include "Synthesis.h"
#include "grgsm_vitac/grgsm_vitac.h"
#include "grgsm_vitac/viterbi_detector.h"

extern "C" {
#include "convert.h"
#include "proto_trxd.h" /* reference TRXD uplink datagram writers (proto_trxd.c, compiled unmodified) */
}
#include <unistd.h>
#include <fcntl.h>
#include <cerrno>

int gVectorDebug = 0; /* declared CommonLibs/Vector.h:45, defined nowhere in the tree */

static void noop_free(void *) {}

namespace {

/* Non-owning view of caller memory as a signalVector. */
struct BurstView {
	signalVector v;
	BurstView(const float *p, size_t n) : v((complex *)p, 0, n, NULL, noop_free) {}
};

// Persistent worker pool (created on first use, grown on demand): the batch entry points are called many times by the
// benchmark's reference arm, and a pool with dynamic chunking keeps thread start-up and static-partition imbalance out of
// the measurement.  Items are handed out in chunks through an atomic counter.
class Pool {
public:
	static Pool &get() { static Pool *p = new Pool; return *p; } // never destroyed: its threads live until exit
	void run(int n, int nthreads, const std::function<void(int)> &f)
	{
		std::unique_lock<std::mutex> lk(mu_);
		while ((int)th_.size() < nthreads - 1) th_.emplace_back([this, id = (int)th_.size()]() { worker(id); });
		fn_ = &f; n_ = n; next_.store(0); active_ = nthreads - 1; pending_ = nthreads - 1;
		chunk_ = std::max(1, std::min(256, n / (8 * nthreads)));
		gen_++;
		lk.unlock();
		cv_.notify_all();
		drain();
		lk.lock();
		done_.wait(lk, [this]() { return pending_ == 0; });
		fn_ = nullptr;
	}

private:
	void drain()
	{
		for (;;) {
			const int lo = next_.fetch_add(chunk_);
			if (lo >= n_) break;
			const int hi = std::min(n_, lo + chunk_);
			for (int i = lo; i < hi; i++) (*fn_)(i);
		}
	}
	void worker(int id)
	{
		unsigned seen = 0;
		for (;;) {
			std::unique_lock<std::mutex> lk(mu_);
			cv_.wait(lk, [&]() { return gen_ != seen; });
			seen = gen_;
			const bool take = id < active_;
			lk.unlock();
			if (take) {
				drain();
				lk.lock();
				if (--pending_ == 0) done_.notify_all();
			}
		}
	}
	std::mutex mu_;
	std::condition_variable cv_, done_;
	std::vector<std::thread> th_;
	const std::function<void(int)> *fn_ = nullptr;
	std::atomic<int> next_{ 0 };
	int n_ = 0, chunk_ = 1, active_ = 0, pending_ = 0;
	unsigned gen_ = 0;
};

template <class F> void parallel_for(int n, int nthreads, F f)
{
	if (nthreads <= 1 || n < 2 * nthreads) {
		for (int i = 0; i < n; i++)
			f(i);
		return;
	}
	const std::function<void(int)> fn = f;
	Pool::get().run(n, nthreads, fn);
}

void fill_bits(BitVector &bv, const uint8_t *bits, int n)
{
	for (int i = 0; i < n; i++)
		bv[i] = bits[i];
}

int copy_sv(const signalVector *sv, float *out, int max_cf)
{
	if (!sv)
		return -1;
	int n = (int)sv->size();
	if (n > max_cf)
		n = max_cf;
	memcpy(out, sv->begin(), sizeof(float) * 2 * n);
	return (int)sv->size();
}

bool g_setup_done = false;

} // namespace

extern "C" {

/* osmo-trx.cpp:648-649 + Transceiver.cpp:207,212 */
int ref_setup(void)
{
	if (g_setup_done)
		return 0;
	convolve_init();
	convert_init();
	if (!sigProcLibSetup())
		return -1;
	initvita();
	g_setup_done = true;
	return 0;
}

/* ---- table dumps (sigProcLib.cpp:52-135 statics) ---- */
int ref_get_table(const char *name, int idx, float *out, int max_floats)
{
	const float *src = NULL;
	int n = 0;
	float tmp[8];
	std::string s(name);
	auto sv = [&](const signalVector *v) { src = (const float *)v->begin(); n = 2 * (int)v->size(); };
	auto cs = [&](const CorrelationSequence *c, bool meta) {
		if (!meta) {
			sv(c->sequence);
		} else {
			tmp[0] = c->gain.real(); tmp[1] = c->gain.imag(); tmp[2] = c->toa;
			src = tmp; n = 3;
		}
	};
	if (s == "sinc") { src = sincTable; n = TABLESIZE + 1; }
	else if (s == "rot4") sv(GMSKRotation4);
	else if (s == "rrot4") sv(GMSKReverseRotation4);
	else if (s == "rot1") sv(GMSKRotation1);
	else if (s == "rrot1") sv(GMSKReverseRotation1);
	else if (s == "delay") { if (idx < 0 || idx >= DELAYFILTS) return -1; sv(delayFilters[idx]); }
	else if (s == "pulse4_c0") sv(GSMPulse4->c0);
	else if (s == "pulse4_c1") sv(GSMPulse4->c1);
	else if (s == "pulse4_c0inv") sv(GSMPulse4->c0_inv);
	else if (s == "pulse1_c0") sv(GSMPulse1->c0);
	else if (s == "midamble" || s == "midamble_meta") { if (idx < 0 || idx > 7) return -1; cs(gMidambles[idx], s == "midamble_meta"); }
	else if (s == "edge_midamble" || s == "edge_midamble_meta") { if (idx < 0 || idx > 7) return -1; cs(gEdgeMidambles[idx], s == "edge_midamble_meta"); }
	else if (s == "rach" || s == "rach_meta") { if (idx < 0 || idx > 2) return -1; cs(gRACHSequences[idx], s == "rach_meta"); }
	else if (s == "sch" || s == "sch_meta") cs(gSCHSequence, s == "sch_meta");
	else if (s == "dummy" || s == "dummy_meta") cs(gDummySequence, s == "dummy_meta");
	else if (s == "psk8") { src = (const float *)psk8_table; n = 16; }
	else return -1;
	if (n > max_floats)
		return -n;
	memcpy(out, src, sizeof(float) * n);
	return n;
}

/* vitac tables (grgsm_vitac.cpp:46-48): which = 0 norm[idx], 1 access, 2 sch */
extern gr_complex d_acc_training_seq[N_ACCESS_BITS];
extern gr_complex d_sch_training_seq[N_SYNC_BITS];
extern gr_complex d_norm_training_seq[TRAIN_SEQ_NUM][N_TRAIN_BITS];
int ref_get_vitac_table(int which, int idx, float *out)
{
	const gr_complex *p; int n;
	if (which == 0) { p = d_norm_training_seq[idx]; n = N_TRAIN_BITS; }
	else if (which == 1) { p = d_acc_training_seq; n = N_ACCESS_BITS; }
	else { p = d_sch_training_seq; n = N_SYNC_BITS; }
	memcpy(out, p, sizeof(float) * 2 * n);
	return n;
}

/* ---- modulators ---- */
/* modulateBurst(bits, guard, sps, empty) sigProcLib.cpp:970-979; out has room for max_cf complex */
int ref_modulate_burst(const uint8_t *bits, int nbits, int guard, int sps, int empty, float *out, int max_cf)
{
	BitVector bv(nbits);
	fill_bits(bv, bits, nbits);
	signalVector *sv = modulateBurst(bv, guard, sps, empty != 0);
	int rc = copy_sv(sv, out, max_cf);
	delete sv;
	return rc;
}

int ref_modulate_gmsk_batch(const uint8_t *bits, int nbits, int n, float *out, int nthreads)
{
	parallel_for(n, nthreads, [&](int b) {
		BitVector bv(nbits);
		fill_bits(bv, bits + (size_t)b * nbits, nbits);
		signalVector *sv = modulateBurst(bv, 8, 4, false);
		copy_sv(sv, out + (size_t)b * 1250, 625);
		delete sv;
	});
	return n;
}

/* modulateEdgeBurst(bits, sps, empty) sigProcLib.cpp:917-936 */
int ref_modulate_edge(const uint8_t *bits, int nbits, int sps, int empty, float *out, int max_cf)
{
	BitVector bv(nbits);
	fill_bits(bv, bits, nbits);
	signalVector *sv = modulateEdgeBurst(bv, sps, empty != 0);
	int rc = copy_sv(sv, out, max_cf);
	delete sv;
	return rc;
}

int ref_modulate_edge_batch(const uint8_t *bits, int nbits, int n, float *out, int nthreads)
{
	parallel_for(n, nthreads, [&](int b) {
		BitVector bv(nbits);
		fill_bits(bv, bits + (size_t)b * nbits, nbits);
		signalVector *sv = modulateEdgeBurst(bv, 4, false);
		copy_sv(sv, out + (size_t)b * 1250, 625);
		delete sv;
	});
	return n;
}

/* ---- detection / demodulation ---- */
/* detectAnyBurst sigProcLib.cpp:1926-1957.  bursts: [n][stride] complex, burst length `blen` (625). */
int ref_detect_batch(const float *bursts, int stride, int blen, int n, const uint8_t *type, const uint8_t *tsc,
		     const uint16_t *max_toa, float thresh, int sps, int32_t *rc, float *amp, float *toa,
		     uint8_t *tsc_out, float *ci, int nthreads)
{
	parallel_for(n, nthreads, [&](int b) {
		BurstView bv(bursts + (size_t)b * stride * 2, blen);
		struct estim_burst_params ebp;
		ebp.amp = 0.0f; ebp.toa = 0.0f; ebp.tsc = 0; ebp.ci = 0.0f;
		rc[b] = detectAnyBurst(bv.v, tsc[b], thresh, sps, (CorrType)type[b], max_toa[b], &ebp);
		amp[2 * b] = ebp.amp.real(); amp[2 * b + 1] = ebp.amp.imag();
		toa[b] = ebp.toa; tsc_out[b] = ebp.tsc; ci[b] = ebp.ci;
	});
	return n;
}

/* detectSCHBurst(burst, thresh, 4, SCH_DETECT_FULL, &ebp) sigProcLib.cpp:1805-1861 per burst */
int ref_detect_sch_batch(const float *bursts, int stride, int blen, int n, float thresh, int32_t *rc, float *amp, float *toa,
			 float *ci)
{
	for (int b = 0; b < n; b++) {
		BurstView bv(bursts + (size_t)b * stride * 2, blen);
		struct estim_burst_params ebp;
		ebp.amp = 0.0f; ebp.toa = 0.0f; ebp.tsc = 0; ebp.ci = 0.0f;
		rc[b] = detectSCHBurst(bv.v, thresh, 4, sch_detect_type::SCH_DETECT_FULL, &ebp);
		amp[2 * b] = ebp.amp.real(); amp[2 * b + 1] = ebp.amp.imag();
		toa[b] = ebp.toa; ci[b] = ebp.ci;
	}
	return n;
}

/* detectSCHBurst(burst, thresh, 4, SCH_DETECT_BUFFER, &ebp) sigProcLib.cpp:1805-1861 per capture (in_len = 60000 samples) */
int ref_detect_sch_buffer_batch(const float *bursts, int stride, int in_len, int n, float thresh, int32_t *rc, float *amp, float *toa,
				float *ci)
{
	for (int b = 0; b < n; b++) {
		BurstView bv(bursts + (size_t)b * stride * 2, in_len);
		struct estim_burst_params ebp;
		ebp.amp = 0.0f; ebp.toa = 0.0f; ebp.tsc = 0; ebp.ci = 0.0f;
		rc[b] = detectSCHBurst(bv.v, thresh, 4, sch_detect_type::SCH_DETECT_BUFFER, &ebp);
		amp[2 * b] = ebp.amp.real(); amp[2 * b + 1] = ebp.amp.imag();
		toa[b] = ebp.toa; ci[b] = ebp.ci;
	}
	return n;
}

/* demodAnyBurst sigProcLib.cpp:2130-2137 for bursts with rc > 0; soft: [n][soft_stride] floats
 * (156 GMSK / 444 EDGE written; rest untouched); nsoft[b] = returned SoftVector size (0 if skipped). */
int ref_demod_batch(const float *bursts, int stride, int blen, int n, const int32_t *rc, const float *amp,
		    const float *toa, float *ci, int sps, float *soft, int soft_stride, int32_t *nsoft, int nthreads)
{
	parallel_for(n, nthreads, [&](int b) {
		nsoft[b] = 0;
		if (rc[b] <= 0)
			return;
		BurstView bv(bursts + (size_t)b * stride * 2, blen);
		struct estim_burst_params ebp;
		ebp.amp = complex(amp[2 * b], amp[2 * b + 1]);
		ebp.toa = toa[b]; ebp.tsc = 0; ebp.ci = ci[b];
		SoftVector *sv = demodAnyBurst(bv.v, (CorrType)rc[b], sps, &ebp);
		if (!sv)
			return;
		int ns = std::min((int)sv->size(), soft_stride);
		memcpy(soft + (size_t)b * soft_stride, sv->begin(), sizeof(float) * ns);
		nsoft[b] = (int)sv->size();
		ci[b] = ebp.ci;
		delete sv;
	});
	return n;
}

/* detect + demod the way Transceiver::pullRadioVector does (Transceiver.cpp:768,786) */
int ref_detect_demod_batch(const float *bursts, int stride, int blen, int n, const uint8_t *type, const uint8_t *tsc,
			   const uint16_t *max_toa, float thresh, int sps, int32_t *rc, float *amp, float *toa,
			   uint8_t *tsc_out, float *ci, float *soft, int soft_stride, int32_t *nsoft, int nthreads)
{
	parallel_for(n, nthreads, [&](int b) {
		BurstView bv(bursts + (size_t)b * stride * 2, blen);
		struct estim_burst_params ebp;
		ebp.amp = 0.0f; ebp.toa = 0.0f; ebp.tsc = 0; ebp.ci = 0.0f;
		nsoft[b] = 0;
		int r = detectAnyBurst(bv.v, tsc[b], thresh, sps, (CorrType)type[b], max_toa[b], &ebp);
		rc[b] = r;
		if (r > 0) {
			SoftVector *sv = demodAnyBurst(bv.v, (CorrType)r, sps, &ebp);
			if (sv) {
				int ns = std::min((int)sv->size(), soft_stride);
				memcpy(soft + (size_t)b * soft_stride, sv->begin(), sizeof(float) * ns);
				nsoft[b] = (int)sv->size();
				delete sv;
			}
		}
		amp[2 * b] = ebp.amp.real(); amp[2 * b + 1] = ebp.amp.imag();
		toa[b] = ebp.toa; tsc_out[b] = ebp.tsc; ci[b] = ebp.ci;
	});
	return n;
}

/* ---- the receive chain around the hot path: int16 slot -> TRXD uplink datagram ----
 * The DSP statements of Transceiver::pullRadioVector (Transceiver.cpp:665-815; the method itself needs the
 * radio FIFO, libosmocore and the state machine, so its statements are replayed here around the reference's own
 * functions), preceded by RadioInterface::pullBuffer's convert_short_float (radioInterface.cpp:345-349) and
 * followed by the UNMODIFIED trxd_send_burst_ind_v0/_v1 (proto_trxd.c:69-117).  capture != 0: the datagram is
 * written to a pipe and copied out (parity); capture == 0: written to /dev/null (timing: the reference pays one
 * write() per burst, to a UDP socket in production). */
int ref_pull_batch(const int16_t *iq, int stride, int n, const uint8_t *type, const uint8_t *tsc, const uint16_t *max_toa,
		   const uint32_t *fn, const uint8_t *tn, float thresh, double full_scale, double rssi_offset, int version,
		   int32_t *rc, float *energy, uint8_t *pkt, int pkt_stride, uint16_t *pkt_len, float *amp, float *toa,
		   float *ci, uint8_t *tsc_out, int capture, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	if (n < 2 * nthreads) nthreads = 1;
	const int per = (n + nthreads - 1) / nthreads;
	auto worker = [&](int lo, int hi) {
		int fds[2] = { -1, -1 };
		if (capture) {
			if (pipe(fds) != 0) return;
		} else {
			fds[1] = open("/dev/null", O_WRONLY);
		}
		std::vector<float> xbuf(2 * 625);
		for (int b = lo; b < hi; b++) {
			struct trx_ul_burst_ind bi;
			struct estim_burst_params ebp;
			ebp.amp = 0.0f; ebp.toa = 0.0f; ebp.tsc = 0; ebp.ci = 0.0f;
			rc[b] = 0; energy[b] = 0.0f; pkt_len[b] = 0;
			if (amp) { amp[2 * b] = 0; amp[2 * b + 1] = 0; }
			if (toa) toa[b] = 0;
			if (ci) ci[b] = 0;
			if (tsc_out) tsc_out[b] = 0;
			/* Transceiver.cpp:693-704 */
			bi.nbits = 0; bi.fn = fn[b]; bi.tn = tn[b]; bi.rssi = 0.0; bi.toa = 0.0; bi.noise = 0.0;
			bi.idle = false; bi.modulation = MODULATION_GMSK; bi.tss = 0; bi.tsc = 0; bi.ci = 0.0;
			const CorrType ctype = (CorrType)type[b];
			if (ctype == OFF) /* :713-716 */
				continue;
			convert_short_float(xbuf.data(), iq + (size_t)b * stride * 2, 2 * 625);
			BurstView bv(xbuf.data(), 625);
			float max = -1.0, avg = 0.0;
			int max_i = -1;
			{ /* :723-731, one diversity path */
				float pow = energyDetect(bv.v, 20 * 4);
				if (pow > max) { max = pow; max_i = 0; }
				avg += pow;
				energy[b] = pow;
			}
			bool idle = true;
			int r = 0;
			if (max_i >= 0) {
				avg = sqrt(avg / (size_t)1);				      /* :742 */
				bi.rssi = 20.0 * log10(full_scale / avg) + rssi_offset;     /* :751 */
				if (ctype != IDLE) {					      /* :754 */
					r = detectAnyBurst(bv.v, tsc[b], thresh, 4, ctype, max_toa[b], &ebp); /* :768 */
					if (r > 0) {
						SoftVector *rxBurst = demodAnyBurst(bv.v, (CorrType)r, 4, &ebp); /* :786 */
						bi.toa = ebp.toa; bi.tsc = ebp.tsc; bi.ci = ebp.ci;
						if (rxBurst->size() == 444) { bi.modulation = MODULATION_8PSK; bi.nbits = 444; }
						else { bi.modulation = MODULATION_GMSK; bi.nbits = 148; }
						vectorSlicer(bi.rx_burst, rxBurst->begin(), bi.nbits);	  /* :803 */
						delete rxBurst;
						idle = false;
					}
				}
			}
			bi.idle = idle; /* ret_idle :810-814 */
			rc[b] = r;
			if (amp) { amp[2 * b] = ebp.amp.real(); amp[2 * b + 1] = ebp.amp.imag(); }
			if (toa) toa[b] = ebp.toa;
			if (ci) ci[b] = ebp.ci;
			if (tsc_out) tsc_out[b] = ebp.tsc;
			/* driveReceiveFIFO :1244-1250 */
			if (version == 0) trxd_send_burst_ind_v0(0, fds[1], &bi);
			else trxd_send_burst_ind_v1(0, fds[1], &bi);
			if (capture) {
				/* v0 drops idle indications without writing: poll the pipe without blocking */
				uint8_t buf[1024];
				int fl = fcntl(fds[0], F_GETFL, 0);
				fcntl(fds[0], F_SETFL, fl | O_NONBLOCK);
				ssize_t got = read(fds[0], buf, sizeof(buf));
				if (got < 0) got = 0;
				if (got > pkt_stride) got = 0;
				memcpy(pkt + (size_t)b * pkt_stride, buf, (size_t)got);
				pkt_len[b] = (uint16_t)got;
			}
		}
		if (fds[0] >= 0) close(fds[0]);
		if (fds[1] >= 0) close(fds[1]);
	};
	if (nthreads == 1) {
		worker(0, n);
	} else {
		std::vector<std::thread> th;
		for (int t = 0; t < nthreads; t++) {
			int lo = t * per, hi = std::min(n, lo + per);
			if (lo >= hi) break;
			th.emplace_back(worker, lo, hi);
		}
		for (auto &t : th) t.join();
	}
	return n;
}

/* helpers exposed for unit parity */
float ref_energy_detect(const float *burst, int blen, unsigned window)
{
	BurstView bv(burst, blen);
	return energyDetect(bv.v, window);
}

int ref_delay_vector(const float *in, int len, float delay, float *out)
{
	BurstView bv(in, len);
	signalVector *sv = delayVector(&bv.v, NULL, delay);
	int rc = copy_sv(sv, out, len);
	delete sv;
	return rc;
}

void ref_vector_slicer(float *dst, const float *src, size_t len) { vectorSlicer(dst, src, len); }

int ref_downsample_burst(const float *in, int blen, float *out)
{
	BurstView bv(in, blen);
	signalVector *sv = downsampleBurst(bv.v);
	int rc = copy_sv(sv, out, 156);
	delete sv;
	return rc;
}

/* generic static convolve() wrapper (sigProcLib.cpp:297-398) for span-mode parity:
 * h_real/h_aligned select the 4 dispatch cases. span: 0 START_ONLY 1 NO_DELAY 2 CUSTOM */
int ref_convolve_sv(const float *x, int x_len, const float *h, int h_len, int h_real, int h_aligned, int span,
		    int start, int len, float *y, int max_cf)
{
	signalVector xv(x_len);
	memcpy(xv.begin(), x, sizeof(float) * 2 * x_len);
	complex *hd = (complex *)convolve_h_alloc(h_len);
	memcpy(hd, h, sizeof(float) * 2 * h_len);
	signalVector hv(hd, 0, h_len, convolve_h_alloc, free);
	hv.isReal(h_real != 0);
	hv.setAligned(h_aligned != 0);
	signalVector *yv = convolve(&xv, &hv, NULL, (ConvType)span, start, len);
	int rc = copy_sv(yv, y, max_cf);
	delete yv;
	return rc;
}

/* ---- raw C kernels (arch/common/convolve.h:4-26) ---- */
int ref_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return convolve_real(x, x_len, h, h_len, y, y_len, start, len);
}
int ref_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return convolve_complex(x, x_len, h, h_len, y, y_len, start, len);
}
int ref_base_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return base_convolve_real(x, x_len, h, h_len, y, y_len, start, len);
}
int ref_base_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start,
			      int len)
{
	return base_convolve_complex(x, x_len, h, h_len, y, y_len, start, len);
}
void *ref_convolve_h_alloc(size_t n) { return convolve_h_alloc(n); }
void ref_free(void *p) { free(p); }

void ref_convert_float_short(short *out, const float *in, float scale, int len) { convert_float_short(out, in, scale, len); }
void ref_convert_short_float(float *out, const short *in, int len) { convert_short_float(out, in, len); }
void ref_base_convert_float_short(short *out, const float *in, float scale, int len) { base_convert_float_short(out, in, scale, len); }
/* the reference's dispatcher is what an SSE3 host runs for any length */
void ref_convert_float_short_x86(short *out, const float *in, float scale, int len) { convert_float_short(out, in, scale, len); }

/* ---- Resampler (Resampler.h:31-61) ---- */
void *ref_resampler_create(int p, int q, int filt_len, float bw)
{
	Resampler *r = new Resampler(p, q, filt_len);
	if (!r->init(bw)) {
		delete r;
		return NULL;
	}
	return r;
}
void ref_resampler_destroy(void *r) { delete (Resampler *)r; }
/* `in` must have filt_len complex samples of history before it (caller supplies in_hist = pointer to start
 * of a buffer [hist | block]); returns rotate()'s rc */
int ref_resampler_rotate(void *r, const float *in_with_hist, int hist, int in_len, float *out, int out_len)
{
	return ((Resampler *)r)->rotate(in_with_hist + 2 * hist, in_len, out, out_len);
}

/* ---- Channelizer / Synthesis (Channelizer.h:13-31, Synthesis.h:13-32) ---- */
void *ref_channelizer_create(int m, int block_len, int h_len)
{
	Channelizer *c = new Channelizer(m, block_len, h_len);
	if (!c->init()) {
		delete c;
		return NULL;
	}
	return c;
}
void ref_channelizer_destroy(void *c) { delete (Channelizer *)c; }
/* in: [block_len*m] complex; out: [m][block_len] complex */
int ref_channelizer_rotate(void *c_, const float *in, int m, int block_len, float *out)
{
	Channelizer *c = (Channelizer *)c_;
	if (!c->rotate(in, (size_t)m * block_len))
		return -1;
	for (int ch = 0; ch < m; ch++)
		memcpy(out + (size_t)ch * block_len * 2, c->outputBuffer(ch), sizeof(float) * 2 * block_len);
	return 0;
}

/* The receive side of radioInterfaceMulti (radioInterfaceMulti.cpp:237-310) for nblk blocks of an m-channel wideband stream:
 * per block Channelizer::rotate, then per channel Resampler(p, q)::rotate of the block_len new samples behind the 16 samples of
 * history the caller keeps (there: the tail of the previous block).  out: [m][nblk * out_per_block] complex.  One thread, as
 * the reference's receive thread.  Returns 0 or -1. */
int ref_wideband_rx(const float *wide, int nblk, int m, int block_len, int p, int q, float *out)
{
	Channelizer ch(m, block_len, 16);
	Resampler rs(p, q, 16);
	if (!ch.init() || !rs.init()) return -1;
	const int out_per_block = block_len / q * p;
	std::vector<float> hist((size_t)m * (16 + block_len) * 2, 0.0f); /* per channel: 16 history samples, then the block */
	for (int b = 0; b < nblk; b++) {
		if (!ch.rotate(wide + (size_t)b * m * block_len * 2, (size_t)m * block_len)) return -1;
		for (int c = 0; c < m; c++) {
			float *hc = hist.data() + (size_t)c * (16 + block_len) * 2;
			memcpy(hc + 32, ch.outputBuffer(c), sizeof(float) * 2 * block_len);
			if (rs.rotate(hc + 32, block_len, out + ((size_t)c * nblk + b) * out_per_block * 2, out_per_block) < 0) return -1;
			memcpy(hc, hc + 2 * block_len, sizeof(float) * 32); /* the block's tail is the next block's history */
		}
	}
	return 0;
}

void *ref_synthesis_create(int m, int block_len, int h_len)
{
	Synthesis *s = new Synthesis(m, block_len, h_len);
	if (!s->init()) {
		delete s;
		return NULL;
	}
	return s;
}
void ref_synthesis_destroy(void *s) { delete (Synthesis *)s; }
/* in: [m][block_len] complex; out: [block_len*m] complex */
int ref_synthesis_rotate(void *s_, const float *in, int m, int block_len, float *out)
{
	Synthesis *s = (Synthesis *)s_;
	for (int ch = 0; ch < m; ch++)
		memcpy(s->inputBuffer(ch), in + (size_t)ch * block_len * 2, sizeof(float) * 2 * block_len);
	return s->rotate(out, (size_t)m * block_len) ? 0 : -1;
}

/* ---- grgsm_vitac (grgsm_vitac.h:65-82) ---- */
/* bufs: [n][stride] complex, burst begins at sample `offset` inside each row (rows are zero padded so
 * negative starts are addressable, ms_upper.cpp:164-171). type: 0 NB (tsc per burst), 1 AB. */
/* detect_burst_nb / detect_burst_ab grgsm_vitac.cpp:105-123 with a given channel estimate and start */
int ref_vitac_detect_batch(const float *bufs, int stride, int offset, int n, int is_ab, const float *cir, const int32_t *start,
			   int8_t *bits)
{
	const int nbits = is_ab ? 88 : 148;
	for (int b = 0; b < n; b++) {
		const gr_complex *in = (const gr_complex *)(bufs + (size_t)b * stride * 2) + offset;
		gr_complex c[CHAN_IMP_RESP_LENGTH * 4];
		memcpy(c, cir + (size_t)b * 40, sizeof(c));
		if (is_ab)
			detect_burst_ab(in, c, start[b], (sbit_t *)bits + (size_t)b * nbits);
		else
			detect_burst_nb(in, c, start[b], (sbit_t *)bits + (size_t)b * nbits);
	}
	return n;
}

/* the five-argument forms (start state `ss`, grgsm_vitac.cpp:105-116) */
int ref_vitac_detect_ss_batch(const float *bufs, int stride, int offset, int n, int is_ab, const float *cir, const int32_t *start,
			      int ss, int8_t *bits)
{
	const int nbits = is_ab ? 88 : 148;
	for (int b = 0; b < n; b++) {
		const gr_complex *in = (const gr_complex *)(bufs + (size_t)b * stride * 2) + offset;
		gr_complex c[CHAN_IMP_RESP_LENGTH * 4];
		memcpy(c, cir + (size_t)b * 40, sizeof(c));
		if (is_ab)
			detect_burst_ab(in, c, start[b], (sbit_t *)bits + (size_t)b * nbits, ss);
		else
			detect_burst_nb(in, c, start[b], (sbit_t *)bits + (size_t)b * nbits, ss);
	}
	return n;
}

int ref_vitac_batch(const float *bufs, int stride, int offset, int n, int is_ab, const uint8_t *tsc, int max_delay,
		    int clamp_lo, int clamp_hi, int8_t *bits, int32_t *start_out, float *corr_max, float *cir_out,
		    int nthreads)
{
	const int nbits = is_ab == 1 ? 88 : 148; /* is_ab: 0 normal, 1 access, 2 SCH burst (ms_rx_lower.cpp:173-177) */
	parallel_for(n, nthreads, [&](int b) {
		const gr_complex *in = (const gr_complex *)(bufs + (size_t)b * stride * 2) + offset;
		gr_complex cir[CHAN_IMP_RESP_LENGTH * 4];
		float cmax = 0.0f;
		int st;
		if (is_ab == 2)
			st = get_sch_chan_imp_resp(in, cir); /* keeps its corr_max to itself: reported as 0 */
		else if (is_ab)
			st = get_access_imp_resp(in, cir, &cmax, max_delay);
		else
			st = get_norm_chan_imp_resp(in, cir, &cmax, tsc[b]);
		st = std::max(clamp_lo, std::min(clamp_hi, st));
		if (is_ab == 1)
			detect_burst_ab(in, cir, st, (sbit_t *)bits + (size_t)b * nbits);
		else
			detect_burst_nb(in, cir, st, (sbit_t *)bits + (size_t)b * nbits);
		start_out[b] = st;
		corr_max[b] = cmax;
		if (cir_out)
			memcpy(cir_out + (size_t)b * 40, cir, sizeof(float) * 40);
	});
	return n;
}

/* get_sch_buffer_chan_imp_resp (grgsm_vitac.cpp:298-309) + detect_burst_nb on the position found, as ms_rx_lower.cpp:168-177
 * does for the first SCH acquisition.  The start may be negative (window - 47 symbols): the demodulated position is limited
 * to the row, which carries `offset` samples of head-room. */
int ref_vitac_sch_buffer_batch(const float *bufs, int stride, int offset, int len, int n, int8_t *bits, int32_t *start_out,
			       float *corr_max, float *cir_out)
{
	for (int b = 0; b < n; b++) {
		const gr_complex *in = (const gr_complex *)(bufs + (size_t)b * stride * 2) + offset;
		gr_complex cir[CHAN_IMP_RESP_LENGTH * 4];
		float cmax = 0.0f;
		int st = get_sch_buffer_chan_imp_resp(in, cir, (unsigned)len, &cmax);
		start_out[b] = st;
		corr_max[b] = cmax;
		if (cir_out)
			memcpy(cir_out + (size_t)b * 40, cir, sizeof(float) * 40);
		if (bits) {
			const int lo = -offset, hi = stride - offset - 148 * 4;
			const int sd = std::max(lo, std::min(hi, st));
			detect_burst_nb(in + sd, cir, 0, (sbit_t *)bits + (size_t)b * 148);
		}
	}
	return n;
}

/* raw viterbi for unit parity: input N complex, rhh 5 complex */
void ref_viterbi(const float *input, int n, const float *rhh, int start_state, float *out)
{
	unsigned int stops[2] = { 4, 12 };
	gr_complex r[5];
	memcpy(r, rhh, sizeof(r));
	viterbi_detector((const gr_complex *)input, n, r, start_state, stops, 2, out);
}

} /* extern "C" */
