// sigproc_b200.cpp — host-side C++ mirror of the reference's burst-DSP API on top of libtrxb200.so.
//
// Implements the functions declared in host/include/{sigProcLib,convolve,Resampler,Channelizer,Synthesis,
// grgsm_vitac}.h — the interfaces Transceiver.cpp, radioInterface{Resamp,Multi}.cpp, ms/*.cpp and
// utils/va-test/burst-gen.cpp call — as batches of one over the C ABI.  All signal
// arithmetic happens in the CUDA kernels; this file only marshals buffers.  No CUDA headers are needed:
// device memory goes through trxb200_dev_alloc / trxb200_copy_*.
#include "include/sigProcLib.h"
#include "include/convolve.h"
#include "include/convert.h"
#include "include/Resampler.h"
#include "include/Channelizer.h"
#include "include/Synthesis.h"
#include "include/grgsm_vitac.h"
#include "../../include/trxb200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

trxb200_ctx *g_ctx = nullptr;
std::mutex g_mu; // one stream, one set of staging buffers: per-burst calls are serialised (the batched ABI is the fast path)

// 3GPP TS 45.002 bit patterns used by the filler-burst generators (reference: GSM/GSMCommon.cpp:35-68)
const char *const kTsc[8] = {
	"00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110", "01000111101101000100011110",
	"00011010111001000001101011", "01001110101100000100111010", "10100111110110001010011111", "11101111000100101110111100" };
const char *const kEdgeTsc[8] = {
	"111111001111111001111001001001111111111111001111111111001111111001111001001001",
	"111111001111001001111001001001111001001001001111111111001111001001111001001001",
	"111001111111111111001001001111001001001111001111111001111111111111001001001111",
	"111001111111111001001001001111001001111001111111111001111111111001001001001111",
	"111111111001001111001111001001001111111001111111111111111001001111001111001001",
	"111001111111001001001111001111001001111111111111111001111111001001001111001111",
	"001111001111111001001001001001111001001111111111001111001111111001001001001001",
	"001001001111001001001001111111111001111111001111001001001111001001001001111111" };
const char *const kDummyBurst =
	"0001111101101110110000010100100111000001001000100000001111100011100010111000101110001010111010010100011001100111001111010011111000100101111101010000";
const char *const kRachBurst = "0011101001001011011111111001100110101010001111000110111101111110000111001001010110011000";

// growable device scratch buffer
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	void *get(size_t bytes)
	{
		if (bytes > cap) {
			if (p) trxb200_dev_free(g_ctx, p);
			p = nullptr;
			cap = 0;
			if (trxb200_dev_alloc(g_ctx, bytes, &p) != TRXB200_OK) return nullptr;
			cap = bytes;
		}
		return p;
	}
	void drop() { if (p && g_ctx) trxb200_dev_free(g_ctx, p); p = nullptr; cap = 0; }
};
DevBuf d_in, d_out, d_a, d_b, d_c, d_d, d_e, d_f, d_g;

bool ok(int rc, const char *what)
{
	if (rc == TRXB200_OK) return true;
	fprintf(stderr, "sigproc_b200: %s failed (%d): %s\n", what, rc, g_ctx ? trxb200_last_error(g_ctx) : "no context");
	return false;
}

signalVector *modulate(const BitVector &bits, bool edge)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || bits.size() == 0) return nullptr;
	const int nbits = (int)bits.size();
	uint8_t *db = (uint8_t *)d_in.get(nbits);
	float *dw = (float *)d_out.get(625 * 8);
	if (!db || !dw) return nullptr;
	if (!ok(trxb200_copy_to_device(g_ctx, db, bits.begin(), nbits), "copy")) return nullptr;
	const int rc = edge ? trxb200_modulate_edge_batch(g_ctx, db, nbits, nbits, 1, dw, 625)
			    : trxb200_modulate_gmsk_batch(g_ctx, db, nbits, nbits, 1, dw, 625);
	if (!ok(rc, edge ? "modulate_edge" : "modulate_gmsk")) return nullptr;
	signalVector *out = new signalVector(625);
	if (!ok(trxb200_copy_to_host(g_ctx, out->begin(), dw, 625 * 8), "copy")) { delete out; return nullptr; }
	return out;
}

// copies `n` complex samples of a burst (zero padded to 625) to the device; returns the device row
float *upload_burst(const signalVector &b, DevBuf &buf)
{
	float *d = (float *)buf.get(625 * 8);
	if (!d) return nullptr;
	const size_t n = b.size() < 625 ? b.size() : 625;
	if (n < 625 && trxb200_memset_device(g_ctx, d, 0, 625 * 8) != TRXB200_OK) return nullptr;
	if (trxb200_copy_to_device(g_ctx, d, b.begin(), n * 8) != TRXB200_OK) return nullptr;
	return d;
}

} // namespace

trxb200_ctx *sigProcLibContext() { return g_ctx; }

bool sigProcLibSetup()
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_ctx) return true;
	const char *dev = getenv("TRXB200_DEVICE");
	const int rc = trxb200_init(dev ? atoi(dev) : 0, &g_ctx);
	if (rc != TRXB200_OK) {
		fprintf(stderr, "sigProcLibSetup: no usable sm_100 device (trxb200_init = %d); there is no CPU path\n", rc);
		g_ctx = nullptr;
		return false;
	}
	return true;
}

void sigProcLibDestroy(void)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return;
	for (DevBuf *b : { &d_in, &d_out, &d_a, &d_b, &d_c, &d_d, &d_e, &d_f, &d_g }) b->drop();
	trxb200_destroy(g_ctx);
	g_ctx = nullptr;
}

void vectorSlicer(float *dest, const float *src, size_t len)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !len) return;
	float *d = (float *)d_in.get(len * 4);
	if (!d) return;
	if (!ok(trxb200_copy_to_device(g_ctx, d, src, len * 4), "copy")) return;
	if (!ok(trxb200_vector_slicer(g_ctx, d, d, len), "vector_slicer")) return;
	ok(trxb200_copy_to_host(g_ctx, dest, d, len * 4), "copy");
}

// modulateBurstBasic / rotateBurst / rotateEdgeBurst (sigProcLib.cpp:558-580,672-689,938-967): mode as in
// trxb200_modulate_basic_batch; the result has sps * (symbols + guard) samples
static signalVector *modulate_basic(const BitVector &bits, int guard, int sps, int mode)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || bits.size() == 0 || guard < 0) return nullptr;
	const int nbits = (int)bits.size();
	const int olen = sps * ((mode == 2 ? nbits / 3 : nbits) + guard);
	uint8_t *db = (uint8_t *)d_in.get(nbits);
	float *dw = (float *)d_out.get((size_t)olen * 8);
	if (!db || !dw) return nullptr;
	if (!ok(trxb200_copy_to_device(g_ctx, db, bits.begin(), nbits), "copy")) return nullptr;
	if (!ok(trxb200_modulate_basic_batch(g_ctx, db, nbits, nbits, 1, guard, sps, mode, dw, olen), "modulate_basic")) return nullptr;
	signalVector *out = new signalVector(olen);
	if (!ok(trxb200_copy_to_host(g_ctx, out->begin(), dw, (size_t)olen * 8), "copy")) { delete out; return nullptr; }
	return out;
}

signalVector *modulateBurst(const BitVector &wBurst, int guardPeriodLength, int sps, bool emptyPulse)
{
	if (emptyPulse) return (sps == 1 || sps == 4) ? modulate_basic(wBurst, guardPeriodLength, sps, 1) : nullptr; // rotateBurst
	if (sps == 4) return modulate(wBurst, false); // Laurent; the guard argument is ignored (sigProcLib.cpp:977)
	if (sps == 1) return modulate_basic(wBurst, guardPeriodLength, 1, 0); // modulateBurstBasic
	return nullptr;
}

signalVector *modulateEdgeBurst(const BitVector &bits, int sps, bool emptyPulse)
{
	if (bits.size() % 3) return nullptr;
	if (emptyPulse) return (sps == 1 || sps == 4) ? modulate_basic(bits, 0, sps, 2) : nullptr; // rotateEdgeBurst
	if (sps != 4) return nullptr; // sigProcLib.cpp:922
	return modulate(bits, true);
}

signalVector *genRandNormalBurst(int tsc, int sps, int tn)
{
	if (tsc < 0 || tsc > 7 || tn < 0 || tn > 7 || (sps != 1 && sps != 4)) return nullptr; // :772
	BitVector bits(148); // 3 tail, 57 data, stealing, 26 TSC, stealing, 57 data, 3 tail
	for (int i = 3; i < 60; i++) bits[i] = (char)(rand() % 2);
	for (int i = 0; i < 26; i++) bits[61 + i] = (char)(kTsc[tsc][i] == '1');
	for (int i = 88; i < 145; i++) bits[i] = (char)(rand() % 2);
	return modulateBurst(bits, 8 + !(tn % 4), sps);
}

signalVector *genRandAccessBurst(int delay, int sps, int tn)
{
	if (tn < 0 || tn > 7 || (sps != 1 && sps != 4) || delay < 0 || delay > 68) return nullptr; // :815
	BitVector bits(88 + delay); // delay zeros, 49 head+sync bits, 36 random, 3 tail
	for (int i = 0; i < 49; i++) bits[delay + i] = (char)(kRachBurst[i] == '1');
	for (int i = 49; i < 85; i++) bits[delay + i] = (char)(rand() % 2);
	return modulateBurst(bits, 68 - delay + !(tn % 4), sps);
}

signalVector *generateDummyBurst(int sps, int tn)
{
	if ((sps != 1 && sps != 4) || tn < 0 || tn > 7) return nullptr; // :858
	return modulateBurst(BitVector(kDummyBurst), 8 + !(tn % 4), sps);
}

signalVector *generateEmptyBurst(int sps, int tn)
{
	if (tn < 0 || tn > 7) return nullptr;
	if (sps == 4) return new signalVector(625);
	if (sps == 1) return new signalVector(148 + 8 + !(tn % 4));
	return nullptr;
}

signalVector *generateEdgeBurst(int tsc)
{
	if (tsc < 0 || tsc > 7) return nullptr;
	BitVector bits(444); // 3 tail symbols (7), 58 random, 26 TSC, 58 random, 3 tail; 3 bits per symbol, LSB first
	auto put = [&](int sym, unsigned v) { for (int k = 0; k < 3; k++) bits[3 * sym + k] = (char)((v >> k) & 1); };
	for (int s = 0; s < 3; s++) put(s, 7);
	for (int s = 3; s < 61; s++) put(s, (unsigned)(rand() % 8));
	for (int i = 0; i < 78; i++) bits[183 + i] = (char)(kEdgeTsc[tsc][i] == '1');
	for (int s = 87; s < 145; s++) put(s, (unsigned)(rand() % 8));
	for (int s = 145; s < 148; s++) put(s, 7);
	return modulateEdgeBurst(bits, 4);
}

void scaleVector(signalVector &x, complex scale)
{
	// one complex multiply per sample; done through the FIR entry point with a single complex tap so that the
	// arithmetic stays on the device like everything else
	if (!x.size()) return;
	std::vector<float> y(2 * x.size());
	const float h[2] = { scale.real(), scale.imag() };
	if (base_convolve_complex((const float *)x.begin(), (int)x.size(), h, 1, y.data(), (int)x.size(), 0, (int)x.size()) < 0) return;
	memcpy((void *)x.begin(), y.data(), x.size() * 8);
}

signalVector *delayVector(const signalVector *in, signalVector *out, float delay)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !in || !in->size()) return nullptr;
	const int len = (int)in->size();
	float *di = (float *)d_in.get((size_t)len * 8), *dd = (float *)d_a.get(4), *dout = (float *)d_out.get((size_t)len * 8);
	if (!di || !dd || !dout) return nullptr;
	if (!ok(trxb200_copy_to_device(g_ctx, di, in->begin(), (size_t)len * 8), "copy")) return nullptr;
	if (!ok(trxb200_copy_to_device(g_ctx, dd, &delay, 4), "copy")) return nullptr;
	if (!ok(trxb200_delay_vector_batch(g_ctx, di, len, len, 1, dd, dout, len), "delay_vector")) return nullptr;
	signalVector *res = out ? out : new signalVector(len);
	if (!ok(trxb200_copy_to_host(g_ctx, res->begin(), dout, (size_t)len * 8), "copy")) { if (!out) delete res; return nullptr; }
	return res;
}

float energyDetect(const signalVector &rxBurst, unsigned windowLength)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !rxBurst.size()) return 0.0f;
	const int len = (int)rxBurst.size();
	float *di = (float *)d_in.get((size_t)len * 8), *de = (float *)d_a.get(4);
	if (!di || !de) return 0.0f;
	if (!ok(trxb200_copy_to_device(g_ctx, di, rxBurst.begin(), (size_t)len * 8), "copy")) return 0.0f;
	if (!ok(trxb200_energy_detect_batch(g_ctx, di, len, len, 1, windowLength, de), "energy_detect")) return 0.0f;
	float e = 0.0f;
	ok(trxb200_copy_to_host(g_ctx, &e, de, 4), "copy");
	return e;
}

std::vector<BurstResult> detectDemodBursts(const std::vector<const signalVector *> &bursts, const std::vector<CorrType> &type,
					   const std::vector<unsigned> &tsc, const std::vector<unsigned> &max_toa, float threshold)
{
	std::lock_guard<std::mutex> lk(g_mu);
	const size_t n = bursts.size();
	std::vector<BurstResult> res(n);
	for (auto &r : res) { r.rc = -SIGERR_INTERNAL; r.ebp = estim_burst_params{ complex(0, 0), 0.0f, 0, 0.0f }; }
	if (!g_ctx || !n || type.size() != n || tsc.size() != n || max_toa.size() != n) return res;
	// pinned-free host staging: the host entry point chunks and overlaps copies itself
	std::vector<float> h_b(n * 625 * 2, 0.0f), h_amp(n * 2), h_toa(n), h_ci(n), h_soft(n * 444, 0.0f);
	std::vector<uint8_t> h_type(n), h_tsc(n), h_tsc_out(n), h_flags(n);
	std::vector<uint16_t> h_mt(n);
	std::vector<int32_t> h_rc(n);
	unsigned bound = 0;
	for (size_t k = 0; k < n; k++) {
		const size_t m = bursts[k]->size() < 625 ? bursts[k]->size() : 625;
		memcpy(&h_b[k * 1250], bursts[k]->begin(), m * 8);
		h_type[k] = (uint8_t)type[k];
		h_tsc[k] = (uint8_t)(tsc[k] > 255 ? 255 : tsc[k]);
		h_mt[k] = (uint16_t)(max_toa[k] > 65535 ? 65535 : max_toa[k]);
		if (max_toa[k] > bound) bound = max_toa[k];
	}
	const int rc = trxb200_detect_demod_host(g_ctx, h_b.data(), 625, (int)n, h_type.data(), h_tsc.data(), h_mt.data(), (int)bound,
						 threshold, h_rc.data(), h_amp.data(), h_toa.data(), h_tsc_out.data(), h_ci.data(),
						 h_flags.data(), h_soft.data(), 444, 148);
	if (!ok(rc, "detect_demod_host")) return res;
	for (size_t k = 0; k < n; k++) {
		res[k].rc = h_rc[k];
		res[k].ebp.amp = complex(h_amp[2 * k], h_amp[2 * k + 1]);
		res[k].ebp.toa = h_toa[k];
		res[k].ebp.tsc = h_tsc_out[k];
		res[k].ebp.ci = h_ci[k];
		if (h_rc[k] > 0)
			res[k].soft.assign(&h_soft[k * 444], &h_soft[k * 444] + (h_rc[k] == EDGE ? 444 : 148));
	}
	return res;
}

int detectAnyBurst(const signalVector &burst, unsigned tsc, float threshold, int sps, CorrType type, unsigned max_toa,
		   struct estim_burst_params *ebp)
{
	if (!ebp) return -SIGERR_INTERNAL;
	if (sps != 4 && sps != 1) return -SIGERR_UNSUPPORTED; // sigProcLib.cpp:1740
	// one sample per symbol: the vector is one slot (156 / 157 symbols, radioInterface.cpp:257-258); other lengths are refused
	const int blen1 = (int)burst.size();
	if (sps == 1 && (blen1 < 148 || blen1 > 160)) return -SIGERR_UNSUPPORTED;
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return -SIGERR_INTERNAL;
	float *db = upload_burst(burst, d_in);
	uint8_t *dt = (uint8_t *)d_a.get(16);
	float *dres = (float *)d_b.get(64);
	if (!db || !dt || !dres) return -SIGERR_INTERNAL;
	// small parameter block: type, tsc (bytes 0,1), max_toa (u16 at byte 2), tsc_out (byte 8), flags (byte 9)
	uint8_t par[16] = { 0 };
	par[0] = (uint8_t)type;
	par[1] = (uint8_t)(tsc > 255 ? 255 : tsc);
	const uint16_t mt = (uint16_t)(max_toa > 1024 ? 1025 : max_toa);
	memcpy(par + 2, &mt, 2);
	if (!ok(trxb200_copy_to_device(g_ctx, dt, par, 16), "copy")) return -SIGERR_INTERNAL;
	int32_t *drc = (int32_t *)dres;
	float *damp = dres + 2, *dtoa = dres + 4, *dci = dres + 5;
	const int rc = sps == 1 ? trxb200_detect_sps1_batch(g_ctx, db, 625, blen1, 1, dt, dt + 1, (const uint16_t *)(dt + 2), mt > 1024 ? 1024 : mt,
							    threshold, drc, damp, dtoa, dt + 8, dci, dt + 9)
				: trxb200_detect_batch(g_ctx, db, 625, 1, dt, dt + 1, (const uint16_t *)(dt + 2), mt > 1024 ? 1024 : mt, threshold, drc,
						       damp, dtoa, dt + 8, dci, dt + 9);
	if (!ok(rc, "detect_batch")) return -SIGERR_INTERNAL;
	float hres[8];
	uint8_t hpar[16];
	if (!ok(trxb200_copy_to_host(g_ctx, hres, dres, 32), "copy") || !ok(trxb200_copy_to_host(g_ctx, hpar, dt, 16), "copy"))
		return -SIGERR_INTERNAL;
	int32_t r;
	memcpy(&r, &hres[0], 4);
	ebp->amp = complex(hres[2], hres[3]);
	ebp->toa = hres[4];
	ebp->ci = hres[5];
	ebp->tsc = hpar[8];
	return r;
}

// sigProcLib.cpp:1805-1861.  SCH_DETECT_FULL (one burst, trxb200_detect_sch_batch) and SCH_DETECT_BUFFER (the 12-frame
// capture of the first acquisition, trxb200_detect_sch_buffer_batch) are built; the NARROW state reads past its 8-sample
// decimated vector in the reference and fails loudly here.
int detectSCHBurst(signalVector &burst, float threshold, int sps, sch_detect_type state, struct estim_burst_params *ebp)
{
	if (!ebp || sps != 4 || (state != sch_detect_type::SCH_DETECT_FULL && state != sch_detect_type::SCH_DETECT_BUFFER)) return -1;
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return -1;
	float *dres = (float *)d_b.get(64);
	if (!dres) return -1;
	int32_t *drc = (int32_t *)dres;
	float *damp = dres + 2, *dtoa = dres + 4, *dci = dres + 5;
	if (state == sch_detect_type::SCH_DETECT_BUFFER) {
		const int in_len = (12 * 8 * 625 / 4) * 4; // downsampleBurst(burst, len * 4, len), len = 15000 (:1827,1841)
		if (burst.size() < (size_t)in_len) return -1; // the reference would copy past the end of the vector
		float *db = (float *)d_in.get((size_t)in_len * 8);
		if (!db || trxb200_copy_to_device(g_ctx, db, burst.begin(), (size_t)in_len * 8) != TRXB200_OK) return -1;
		if (!ok(trxb200_detect_sch_buffer_batch(g_ctx, db, in_len, in_len, 1, threshold, drc, damp, dtoa, dci, nullptr), "detect_sch_buffer_batch"))
			return -1;
	} else {
	float *db = upload_burst(burst, d_in);
	if (!db) return -1;
	if (!ok(trxb200_detect_sch_batch(g_ctx, db, 625, 1, threshold, drc, damp, dtoa, dci, nullptr), "detect_sch_batch")) return -1;
	}
	float hres[8];
	if (!ok(trxb200_copy_to_host(g_ctx, hres, dres, 32), "copy")) return -1;
	int32_t r;
	memcpy(&r, &hres[0], 4);
	ebp->amp = complex(hres[2], hres[3]);
	ebp->toa = hres[4];
	if (r > 0) ebp->ci = hres[5]; // the reference leaves ci untouched when nothing is found
	return r;
}

SoftVector *demodAnyBurst(const signalVector &burst, CorrType type, int sps, struct estim_burst_params *ebp)
{
	if (!ebp || (sps != 4 && sps != 1)) return nullptr;
	const int blen1 = (int)burst.size();
	if (sps == 1 && (blen1 < 148 || blen1 > 160)) return nullptr;
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return nullptr;
	float *db = upload_burst(burst, d_in);
	float *dpar = (float *)d_b.get(64);
	const int nsoft = type == EDGE ? 444 : (sps == 1 ? blen1 : 156); // demodGmskBurst returns the 1-sps vector's length
	float *dsoft = (float *)d_out.get((size_t)nsoft * 4);
	if (!db || !dpar || !dsoft) return nullptr;
	float hpar[8] = { 0 };
	const int32_t rc = (int32_t)type;
	memcpy(&hpar[0], &rc, 4);
	hpar[2] = ebp->amp.real();
	hpar[3] = ebp->amp.imag();
	hpar[4] = ebp->toa;
	hpar[5] = ebp->ci;
	if (!ok(trxb200_copy_to_device(g_ctx, dpar, hpar, 32), "copy")) return nullptr;
	if (sps == 1) {
		if (!ok(trxb200_demod_sps1_batch(g_ctx, db, 625, blen1, 1, (const int32_t *)dpar, dpar + 2, dpar + 4, dpar + 5, dsoft, nsoft, blen1),
			"demod_sps1_batch"))
			return nullptr;
	} else if (!ok(trxb200_demod_batch(g_ctx, db, 625, 1, (const int32_t *)dpar, dpar + 2, dpar + 4, dpar + 5, dsoft, nsoft, 156), "demod_batch"))
		return nullptr;
	SoftVector *out = new SoftVector(nsoft);
	if (!ok(trxb200_copy_to_host(g_ctx, out->begin(), dsoft, (size_t)nsoft * 4), "copy")) { delete out; return nullptr; }
	if (type == EDGE) {
		float ci = 0.0f; // demodEdgeBurst overwrites ebp->ci (sigProcLib.cpp:2118)
		if (ok(trxb200_copy_to_host(g_ctx, &ci, dpar + 5, 4), "copy")) ebp->ci = ci;
	}
	return out;
}

/* ---------------- convolve.h ---------------- */
static int conv_one(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len, bool cplx, int base)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return -1;
	if (x_len < 1 || h_len < 1 || y_len < 1 || len < 1 || start + len > x_len || len > y_len || x_len < h_len) return -1;
	// the taps reach back to x[start - (h_len - 1)]: ship that span (head-room included) and index from its start
	const int first = start - (h_len - 1);
	const int span = len + h_len - 1;
	float *dx = (float *)d_in.get((size_t)span * 8), *dh = (float *)d_a.get((size_t)h_len * 8), *dy = (float *)d_out.get((size_t)len * 8);
	if (!dx || !dh || !dy) return -1;
	if (trxb200_copy_to_device(g_ctx, dx, x + 2 * (long)first, (size_t)span * 8) != TRXB200_OK) return -1;
	if (trxb200_copy_to_device(g_ctx, dh, h, (size_t)h_len * 8) != TRXB200_OK) return -1;
	const int rc = cplx ? trxb200_convolve_complex_batch(g_ctx, dx, span, span, dh, h_len, dy, len, len, h_len - 1, len, 1, base)
			    : trxb200_convolve_real_batch(g_ctx, dx, span, span, dh, h_len, dy, len, len, h_len - 1, len, 1, base);
	if (rc != TRXB200_OK) return -1;
	if (trxb200_copy_to_host(g_ctx, y, dy, (size_t)len * 8) != TRXB200_OK) return -1;
	return len;
}

extern "C" {
void *convolve_h_alloc(size_t num) { void *p = nullptr; return posix_memalign(&p, 16, num * 2 * sizeof(float)) ? nullptr : p; }
// The reference's arch library is usable on its own (tests/Transceiver52M/convolve_test.c calls convolve_init() and
// nothing else): its init entry points open the GPU context like sigProcLibSetup() does.
void convolve_init(void) { sigProcLibSetup(); }
int convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return conv_one(x, x_len, h, h_len, y, y_len, start, len, false, 0);
}
int convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return conv_one(x, x_len, h, h_len, y, y_len, start, len, true, 0);
}
int base_convolve_real(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return conv_one(x, x_len, h, h_len, y, y_len, start, len, false, 1);
}
int base_convolve_complex(const float *x, int x_len, const float *h, int h_len, float *y, int y_len, int start, int len)
{
	return conv_one(x, x_len, h, h_len, y, y_len, start, len, true, 1);
}
}

/* ---------------- convert.h ---------------- */
static void convert_fs(short *out, const float *in, float scale, int len, int mode)
{
	if (len <= 0) return;
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return;
	float *di = (float *)d_in.get((size_t)len * 4);
	int16_t *d_o = (int16_t *)d_out.get((size_t)len * 2);
	if (!di || !d_o) return;
	if (!ok(trxb200_copy_to_device(g_ctx, di, in, (size_t)len * 4), "copy")) return;
	if (!ok(trxb200_convert_float_short_mode(g_ctx, d_o, di, scale, (size_t)len, mode), "convert_float_short")) return;
	ok(trxb200_copy_to_host(g_ctx, out, d_o, (size_t)len * 2), "copy");
}
static void convert_sf(float *out, const short *in, int len)
{
	if (len <= 0) return;
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return;
	int16_t *di = (int16_t *)d_in.get((size_t)len * 2);
	float *d_o = (float *)d_out.get((size_t)len * 4);
	if (!di || !d_o) return;
	if (!ok(trxb200_copy_to_device(g_ctx, di, in, (size_t)len * 2), "copy")) return;
	if (!ok(trxb200_convert_short_float(g_ctx, d_o, di, (size_t)len), "convert_short_float")) return;
	ok(trxb200_copy_to_host(g_ctx, out, d_o, (size_t)len * 4), "copy");
}
extern "C" {
void convert_init(void) { sigProcLibSetup(); }
void convert_float_short(short *out, const float *in, float scale, int len) { convert_fs(out, in, scale, len, 1); }
void base_convert_float_short(short *out, const float *in, float scale, int len) { convert_fs(out, in, scale, len, 2); }
void convert_short_float(float *out, const short *in, int len) { convert_sf(out, in, len); }
void base_convert_short_float(float *out, const short *in, int len) { convert_sf(out, in, len); }
}

/* ---------------- Resampler ---------------- */
Resampler::Resampler(size_t p_, size_t q_, size_t fl) : p(p_), q(q_), filt_len(fl) {}
Resampler::~Resampler() { if (h) trxb200_resampler_destroy(h); }
bool Resampler::init(float bw)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return false;
	if (h) { trxb200_resampler_destroy(h); h = nullptr; }
	return trxb200_resampler_create(g_ctx, (int)p, (int)q, (int)filt_len, bw, &h) == TRXB200_OK;
}
size_t Resampler::len() { return filt_len; }
int Resampler::rotate(const float *in, size_t in_len, float *out, size_t out_len)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !h) return -1;
	const size_t hist = filt_len; // samples the filter reads before `in`
	float *di = (float *)d_in.get((in_len + hist) * 8), *dout = (float *)d_out.get(out_len * 8);
	if (!di || !dout) return -1;
	if (trxb200_copy_to_device(g_ctx, di, in - 2 * hist, (in_len + hist) * 8) != TRXB200_OK) return -1;
	if (trxb200_resampler_rotate(h, di + 2 * hist, (int)in_len, (int)(in_len + hist), dout, (int)out_len, (int)out_len, 1) != TRXB200_OK)
		return -1;
	if (trxb200_copy_to_host(g_ctx, out, dout, out_len * 8) != TRXB200_OK) return -1;
	return (int)out_len;
}

/* ---------------- Channelizer / Synthesis ---------------- */
ChannelizerBase::ChannelizerBase(size_t m_, size_t bl, size_t hl, bool synth) : m(m_), hLen(hl), blockLen(bl), synthesis(synth) {}
ChannelizerBase::~ChannelizerBase() { if (fb) trxb200_filterbank_destroy(fb); }
bool ChannelizerBase::checkLen(size_t innerLen, size_t outerLen) { return innerLen == blockLen && outerLen == blockLen * m; }
bool ChannelizerBase::init()
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return false;
	if (fb) { trxb200_filterbank_destroy(fb); fb = nullptr; }
	const int rc = synthesis ? trxb200_synthesis_create(g_ctx, (int)m, (int)blockLen, (int)hLen, &fb)
				 : trxb200_channelizer_create(g_ctx, (int)m, (int)blockLen, (int)hLen, &fb);
	if (rc != TRXB200_OK) return false;
	chanBuf.assign(m, std::vector<float>(2 * blockLen, 0.0f));
	return true;
}

Channelizer::Channelizer(size_t m_, size_t bl, size_t hl) : ChannelizerBase(m_, bl, hl, false) {}
Channelizer::~Channelizer() {}
size_t Channelizer::inputLen() const { return blockLen * m; }
size_t Channelizer::outputLen() const { return blockLen; }
float *Channelizer::outputBuffer(size_t chan) const { return chan < m && !chanBuf.empty() ? const_cast<float *>(chanBuf[chan].data()) : nullptr; }
bool Channelizer::rotate(const float *in, size_t iLen)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !fb || !checkLen(blockLen, iLen)) return false;
	const size_t total = blockLen * m;
	float *di = (float *)d_in.get(total * 8), *dout = (float *)d_out.get(total * 8);
	if (!di || !dout) return false;
	if (trxb200_copy_to_device(g_ctx, di, in, total * 8) != TRXB200_OK) return false;
	if (trxb200_channelizer_rotate(fb, di, dout, 1) != TRXB200_OK) return false;
	std::vector<float> host(2 * total);
	if (trxb200_copy_to_host(g_ctx, host.data(), dout, total * 8) != TRXB200_OK) return false;
	for (size_t c = 0; c < m; c++) memcpy(chanBuf[c].data(), &host[2 * c * blockLen], blockLen * 8);
	return true;
}

Synthesis::Synthesis(size_t m_, size_t bl, size_t hl) : ChannelizerBase(m_, bl, hl, true) {}
Synthesis::~Synthesis() {}
size_t Synthesis::inputLen() const { return blockLen; }
size_t Synthesis::outputLen() const { return blockLen * m; }
float *Synthesis::inputBuffer(size_t chan) const { return chan < m && !chanBuf.empty() ? const_cast<float *>(chanBuf[chan].data()) : nullptr; }
bool Synthesis::resetBuffer(size_t chan)
{
	if (chan >= m || chanBuf.empty()) return false;
	std::fill(chanBuf[chan].begin(), chanBuf[chan].end(), 0.0f);
	return true;
}
bool Synthesis::rotate(float *out, size_t oLen)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !fb || !checkLen(blockLen, oLen)) return false;
	const size_t total = blockLen * m;
	std::vector<float> host(2 * total);
	for (size_t c = 0; c < m; c++) memcpy(&host[2 * c * blockLen], chanBuf[c].data(), blockLen * 8);
	float *di = (float *)d_in.get(total * 8), *dout = (float *)d_out.get(total * 8);
	if (!di || !dout) return false;
	if (trxb200_copy_to_device(g_ctx, di, host.data(), total * 8) != TRXB200_OK) return false;
	if (trxb200_synthesis_rotate(fb, di, dout, 1) != TRXB200_OK) return false;
	return trxb200_copy_to_host(g_ctx, out, dout, total * 8) == TRXB200_OK;
}

/* ---------------- grgsm_vitac ---------------- */
gr_complex d_acc_training_seq[N_ACCESS_BITS];
gr_complex d_sch_training_seq[N_SYNC_BITS];
gr_complex d_norm_training_seq[TRAIN_SEQ_NUM][N_TRAIN_BITS];

namespace {
constexpr int kVitPad = 40, kVitRow = kVitPad + 1024 + kVitPad;
int g_vit_head = 0, g_vit_avail = 625; // samples the caller owns before / from `input` (vitac_input_headroom)

// the device row: kVitPad zeros | caller samples | zeros; input[0] sits at sample kVitPad
bool upload_vitac_row(const gr_complex *input, float *drow)
{
	if (trxb200_memset_device(g_ctx, drow, 0, (size_t)kVitRow * 8) != TRXB200_OK) return false;
	const int head = g_vit_head < kVitPad ? g_vit_head : kVitPad;
	const int avail = g_vit_avail < kVitRow - kVitPad ? g_vit_avail : kVitRow - kVitPad;
	return trxb200_copy_to_device(g_ctx, drow + 2 * (kVitPad - head), input - head, (size_t)(head + avail) * 8) == TRXB200_OK;
}

// one GPU call: CIR search (+ detection with the start clamped to [lo, hi])
bool run_vitac(const gr_complex *input, int is_ab, int tsc, int max_delay, int lo, int hi, gr_complex *cir, float *corr_max, int *start,
	       sbit_t *bits)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return false;
	const int N = is_ab == 1 ? 88 : 148; // is_ab: 0 normal, 1 access, 2 SCH burst
	float *drow = (float *)d_in.get((size_t)kVitRow * 8);
	uint8_t *dt = (uint8_t *)d_a.get(16);
	int8_t *dbits = (int8_t *)d_c.get(160);
	float *dres = (float *)d_b.get(8 + 20 * 8);
	if (!drow || !dt || !dbits || !dres) return false;
	if (!upload_vitac_row(input, drow)) return false;
	const uint8_t t8 = (uint8_t)tsc;
	if (trxb200_copy_to_device(g_ctx, dt, &t8, 1) != TRXB200_OK) return false;
	lo = lo < -kVitPad ? -kVitPad : lo;
	hi = hi > kVitRow - kVitPad - 4 * N ? kVitRow - kVitPad - 4 * N : hi;
	if (lo > hi) lo = hi;
	if (!ok(trxb200_vitac_batch(g_ctx, drow, kVitRow, kVitPad, 1, is_ab, dt, max_delay, lo, hi, dbits, (int32_t *)dres, dres + 1, dres + 2),
		"vitac_batch"))
		return false;
	float hres[2 + 40];
	if (trxb200_copy_to_host(g_ctx, hres, dres, sizeof(hres)) != TRXB200_OK) return false;
	if (start) memcpy(start, &hres[0], 4);
	if (corr_max) *corr_max = hres[1];
	if (cir) memcpy((void *)cir, &hres[2], 20 * 8);
	if (bits && trxb200_copy_to_host(g_ctx, bits, dbits, (size_t)N) != TRXB200_OK) return false;
	return true;
}
} // namespace

void vitac_input_headroom(int samples_before, int samples_from_input)
{
	std::lock_guard<std::mutex> lk(g_mu);
	g_vit_head = samples_before < 0 ? 0 : samples_before;
	g_vit_avail = samples_from_input < 1 ? 1 : samples_from_input;
}

// the reference builds its symbol tables here (grgsm_vitac.cpp:51-80); the device copies live in the GPU context
// (sigProcLibSetup), the exported host arrays are filled from it
void initvita()
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) return;
	float buf[2 * 64];
	for (int t = 0; t < TRAIN_SEQ_NUM; t++)
		if (trxb200_get_table(g_ctx, "vitac_norm", t, buf, 2 * N_TRAIN_BITS) == 2 * N_TRAIN_BITS)
			memcpy((void *)d_norm_training_seq[t], buf, sizeof(d_norm_training_seq[t]));
	if (trxb200_get_table(g_ctx, "vitac_access", 0, buf, 2 * N_ACCESS_BITS) == 2 * N_ACCESS_BITS)
		memcpy((void *)d_acc_training_seq, buf, sizeof(d_acc_training_seq));
	if (trxb200_get_table(g_ctx, "vitac_sch", 0, buf, 2 * N_SYNC_BITS) == 2 * N_SYNC_BITS)
		memcpy((void *)d_sch_training_seq, buf, sizeof(d_sch_training_seq));
}

int get_norm_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, float *corr_max, int bcc)
{
	int start = 0;
	if (!run_vitac(input, 0, bcc, 0, -kVitPad, 1 << 20, chan_imp_resp, corr_max, &start, nullptr)) return 0;
	return start;
}

int get_access_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, float *corr_max, int max_delay)
{
	int start = 0;
	if (!run_vitac(input, 1, 0, max_delay, -kVitPad, 1 << 20, chan_imp_resp, corr_max, &start, nullptr)) return 0;
	return start;
}

// grgsm_vitac.cpp:283-296; the reference's function keeps corr_max to itself
int get_sch_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp)
{
	int start = 0;
	float cmax = 0.0f;
	if (!run_vitac(input, 2, 0, 0, -kVitPad, 1 << 20, chan_imp_resp, &cmax, &start, nullptr)) return 0;
	return start;
}

// grgsm_vitac.cpp:298-309: the first SCH acquisition over a capture of `len` samples starting at `input`
int get_sch_buffer_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, unsigned int len, float *corr_max)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !input || !chan_imp_resp || len < 64 * 8 + 20) return 0;
	float *drow = (float *)d_in.get((size_t)len * 8);
	float *dres = (float *)d_b.get(16 + 20 * 8);
	if (!drow || !dres) return 0;
	if (trxb200_copy_to_device(g_ctx, drow, input, (size_t)len * 8) != TRXB200_OK) return 0;
	int32_t *dstart = (int32_t *)dres;
	float *dcmax = dres + 1, *dcir = dres + 4;
	if (!ok(trxb200_vitac_sch_buffer_batch(g_ctx, drow, (int)len, 0, (int)len, 1, nullptr, dstart, dcmax, dcir), "vitac_sch_buffer_batch"))
		return 0;
	float h[4 + 40];
	if (trxb200_copy_to_host(g_ctx, h, dres, sizeof(h)) != TRXB200_OK) return 0;
	int32_t st;
	memcpy(&st, &h[0], 4);
	if (corr_max) *corr_max = h[1];
	memcpy((void *)chan_imp_resp, &h[4], 20 * 8);
	return st;
}

// detect_burst_nb / detect_burst_ab (grgsm_vitac.cpp:105-123): matched filter + Viterbi with the CALLER's channel estimate
// and start, whatever pointer arithmetic the caller did between the estimate and this call (ms_rx_lower.cpp:177 passes
// &ss[start] with start 0)
static bool run_vitac_detect(const gr_complex *input, int is_ab, const gr_complex *cir, int burst_start, int ss, sbit_t *bits)
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx || !cir) return false;
	const int N = is_ab ? 88 : 148;
	float *drow = (float *)d_in.get((size_t)kVitRow * 8);
	float *dcir = (float *)d_b.get(8 + 20 * 8);
	int32_t *dstart = (int32_t *)d_a.get(16);
	int8_t *dbits = (int8_t *)d_c.get(160);
	if (!drow || !dcir || !dstart || !dbits) return false;
	if (!upload_vitac_row(input, drow)) return false;
	if (trxb200_copy_to_device(g_ctx, dcir, cir, 20 * 8) != TRXB200_OK) return false;
	int st = burst_start < -kVitPad ? -kVitPad : burst_start;
	if (st > kVitRow - kVitPad - 4 * N) st = kVitRow - kVitPad - 4 * N;
	const int32_t st32 = st;
	if (trxb200_copy_to_device(g_ctx, dstart, &st32, 4) != TRXB200_OK) return false;
	if (!ok(trxb200_vitac_detect_ss_batch(g_ctx, drow, kVitRow, kVitPad, 1, is_ab, dcir, dstart, st, st, ss, dbits), "vitac_detect_batch"))
		return false;
	return trxb200_copy_to_host(g_ctx, bits, dbits, (size_t)N) == TRXB200_OK;
}

void detect_burst_nb(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary, int ss)
{
	if (!run_vitac_detect(input, 0, chan_imp_resp, burst_start, ss, output_binary))
		memset(output_binary, 0, 148);
}
void detect_burst_ab(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary, int ss)
{
	if (!run_vitac_detect(input, 1, chan_imp_resp, burst_start, ss, output_binary))
		memset(output_binary, 0, 88);
}
void detect_burst_nb(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary)
{
	detect_burst_nb(input, chan_imp_resp, burst_start, output_binary, 3);
}
void detect_burst_ab(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary)
{
	detect_burst_ab(input, chan_imp_resp, burst_start, output_binary, 3);
}
