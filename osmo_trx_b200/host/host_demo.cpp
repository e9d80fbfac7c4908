// host_demo.cpp — drives the per-burst C++ mirror the way Transceiver::pullRadioVector / addRadioVector and
// burst-gen.cpp drive the reference (setup once, then per-burst calls); used by tests/test_host_mirror.py.
//   host_demo dd   in.bin out.bin   detectAnyBurst + demodAnyBurst per burst
//   host_demo mod  in.bin out.bin   modulateBurst / modulateEdgeBurst per bit vector
//   host_demo conv in.bin out.bin   convolve_real / convolve_complex / base_* single calls
//   host_demo vit  in.bin out.bin   get_norm_chan_imp_resp + detect_burst_nb per burst
//   host_demo misc in.bin out.bin   a list of single calls: Resampler / Channelizer / Synthesis objects, scaleVector,
//                                   delayVector, energyDetect, vectorSlicer, the filler-burst generators, modulateBurst at
//                                   1 sps / with emptyPulse, convert_*
#include "include/sigProcLib.h"
#include "include/convolve.h"
#include "include/convert.h"
#include "include/Resampler.h"
#include "include/Channelizer.h"
#include "include/Synthesis.h"
#include "include/grgsm_vitac.h"
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

static bool rd(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n; }
static void wr(FILE *f, const void *p, size_t n) { fwrite(p, 1, n, f); }

int main(int argc, char **argv)
{
	if (argc != 4) { fprintf(stderr, "usage: host_demo dd|dd1|sch|acq|mod|conv|vit in out\n"); return 2; }
	const std::string mode = argv[1];
	FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
	if (!fi || !fo) { perror("open"); return 2; }
	if (!sigProcLibSetup()) return 3;
	initvita();
	int32_t n = 0;
	if (!rd(fi, &n, 4)) return 2;
	if (mode == "dd" || mode == "dd1") {
		// dd: 625-sample bursts at 4 samples per symbol; dd1: a fourth header word gives the length of a burst at ONE sample per symbol
		const int sps = mode == "dd1" ? 1 : 4;
		for (int k = 0; k < n; k++) {
			int32_t hdr[4] = { 0, 0, 0, 625 };
			if (!rd(fi, hdr, sps == 1 ? 16 : 12) || hdr[3] < 1 || hdr[3] > 625) return 2;
			signalVector burst(hdr[3]);
			if (!rd(fi, burst.begin(), (size_t)hdr[3] * 8)) return 2;
			estim_burst_params ebp;
			ebp = estim_burst_params{ complex(0, 0), 0.0f, 0, 0.0f };
			const int rc = detectAnyBurst(burst, (unsigned)hdr[1], BURST_THRESH, sps, (CorrType)hdr[0], (unsigned)hdr[2], &ebp);
			float rec[6] = { (float)rc, ebp.amp.real(), ebp.amp.imag(), ebp.toa, (float)ebp.tsc, ebp.ci };
			std::vector<float> soft(444, 0.0f);
			int32_t nsoft = 0;
			if (rc > 0) {
				std::unique_ptr<SoftVector> sv(demodAnyBurst(burst, (CorrType)rc, sps, &ebp));
				if (!sv) return 4;
				nsoft = (int32_t)sv->size();
				memcpy(soft.data(), sv->begin(), sv->size() * 4);
				rec[5] = ebp.ci; // EDGE demodulation refines C/I
			}
			wr(fo, rec, sizeof(rec));
			wr(fo, &nsoft, 4);
			wr(fo, soft.data(), 444 * 4);
		}
	} else if (mode == "sch") {
		// detectSCHBurst(burst, BURST_THRESH, 4, SCH_DETECT_FULL) per burst; the BUFFER state needs its 12-frame capture and
		// the NARROW state is refused
		for (int k = 0; k < n; k++) {
			signalVector burst(625);
			if (!rd(fi, burst.begin(), 625 * 8)) return 2;
			estim_burst_params ebp;
			ebp = estim_burst_params{ complex(0, 0), 0.0f, 0, 0.0f };
			const int rc = detectSCHBurst(burst, BURST_THRESH, 4, sch_detect_type::SCH_DETECT_FULL, &ebp);
			if (k == 0 && detectSCHBurst(burst, BURST_THRESH, 4, sch_detect_type::SCH_DETECT_BUFFER, &ebp) >= 0) return 5;
			if (k == 0 && detectSCHBurst(burst, BURST_THRESH, 4, sch_detect_type::SCH_DETECT_NARROW, &ebp) >= 0) return 5;
			float rec[5] = { (float)rc, ebp.amp.real(), ebp.amp.imag(), ebp.toa, ebp.ci };
			wr(fo, rec, sizeof(rec));
		}
	} else if (mode == "acq") {
		// first SCH acquisition over 12-frame captures (ms_rx_lower.cpp:160-219): detectSCHBurst(SCH_DETECT_BUFFER), and
		// get_sch_buffer_chan_imp_resp + detect_burst_nb at the position found
		const int L = 60000;
		initvita();
		for (int k = 0; k < n; k++) {
			signalVector cap(L);
			if (!rd(fi, cap.begin(), (size_t)L * 8)) return 2;
			estim_burst_params ebp;
			ebp = estim_burst_params{ complex(0, 0), 0.0f, 0, 0.0f };
			const int rc = detectSCHBurst(cap, BURST_THRESH, 4, sch_detect_type::SCH_DETECT_BUFFER, &ebp);
			float rec[5] = { (float)rc, ebp.amp.real(), ebp.amp.imag(), ebp.toa, ebp.ci };
			wr(fo, rec, sizeof(rec));
			gr_complex cir[CHAN_IMP_RESP_LENGTH * 4];
			float cmax = 0.0f;
			const gr_complex *ss = reinterpret_cast<const gr_complex *>(cap.begin());
			int32_t start = get_sch_buffer_chan_imp_resp(ss, cir, L, &cmax);
			sbit_t bits[148];
			const int sd = start < 0 ? 0 : (start > L - 592 ? L - 592 : start);
			vitac_input_headroom(0, 592 + 64 > L - sd ? L - sd : 592 + 64);
			detect_burst_nb(&ss[sd], cir, 0, bits);
			wr(fo, &start, 4);
			wr(fo, &cmax, 4);
			wr(fo, cir, sizeof(cir));
			wr(fo, bits, 148);
		}
	} else if (mode == "mod") {
		for (int k = 0; k < n; k++) {
			int32_t hdr[2]; // edge?, nbits
			if (!rd(fi, hdr, 8)) return 2;
			BitVector bits((size_t)hdr[1]);
			if (!rd(fi, bits.begin(), (size_t)hdr[1])) return 2;
			std::unique_ptr<signalVector> w(hdr[0] ? modulateEdgeBurst(bits, 4) : modulateBurst(bits, 8, 4));
			if (!w || w->size() != 625) return 4;
			wr(fo, w->begin(), 625 * 8);
		}
	} else if (mode == "conv") {
		for (int k = 0; k < n; k++) {
			int32_t hdr[7]; // x_len, h_len, start, len, complex taps?, base?, head-room
			if (!rd(fi, hdr, 28)) return 2;
			std::vector<float> x(2 * (size_t)(hdr[0] + hdr[6])), h(2 * (size_t)hdr[1]), y(2 * (size_t)hdr[3], 0.0f);
			if (!rd(fi, x.data(), x.size() * 4) || !rd(fi, h.data(), h.size() * 4)) return 2;
			const float *x0 = x.data() + 2 * hdr[6];
			int rc;
			if (hdr[4]) rc = hdr[5] ? base_convolve_complex(x0, hdr[0], h.data(), hdr[1], y.data(), hdr[3], hdr[2], hdr[3])
						 : convolve_complex(x0, hdr[0], h.data(), hdr[1], y.data(), hdr[3], hdr[2], hdr[3]);
			else rc = hdr[5] ? base_convolve_real(x0, hdr[0], h.data(), hdr[1], y.data(), hdr[3], hdr[2], hdr[3])
					 : convolve_real(x0, hdr[0], h.data(), hdr[1], y.data(), hdr[3], hdr[2], hdr[3]);
			int32_t r32 = rc;
			wr(fo, &r32, 4);
			wr(fo, y.data(), y.size() * 4);
		}
	} else if (mode == "vit") {
		for (int k = 0; k < n; k++) {
			int32_t tsc;
			std::vector<gr_complex> buf(40 + 625 + 40, gr_complex(0, 0)); // ms_upper.cpp:164-171 pads 40 zeros
			if (!rd(fi, &tsc, 4) || !rd(fi, buf.data() + 40, 625 * 8)) return 2;
			gr_complex cir[20];
			float cmax = 0.0f;
			int start = get_norm_chan_imp_resp(buf.data() + 40, cir, &cmax, tsc);
			start = start < 39 ? start : 39; // ms_upper.cpp:224-225
			start = start > -39 ? start : -39;
			sbit_t bits[148];
			detect_burst_nb(buf.data() + 40, cir, start, bits);
			int32_t s32 = start;
			wr(fo, &s32, 4);
			wr(fo, &cmax, 4);
			wr(fo, bits, 148);
		}
	} else if (mode == "misc") {
		// n operations, each: int32 op, a, b, c, d + payload; every result is written as int32 count + payload
		auto put_cf = [&](const float *p, size_t ncf) { int32_t c = (int32_t)ncf; wr(fo, &c, 4); wr(fo, p, ncf * 8); };
		for (int k = 0; k < n; k++) {
			int32_t h[5];
			if (!rd(fi, h, 20)) return 2;
			switch (h[0]) {
			case 1: { // Resampler(p = a, q = b).rotate(in_len = c -> out_len = d); 16 samples of history precede `in`
				std::vector<float> x(2 * (size_t)(16 + h[3])), y(2 * (size_t)h[4]);
				if (!rd(fi, x.data(), x.size() * 4)) return 2;
				Resampler r((size_t)h[1], (size_t)h[2]);
				if (!r.init()) return 4;
				if (r.rotate(x.data() + 32, (size_t)h[3], y.data(), (size_t)h[4]) != h[4]) return 4;
				put_cf(y.data(), (size_t)h[4]);
				break;
			}
			case 2: { // Channelizer(m = a, blockLen = b): c blocks through rotate(), outputBuffer(chan) after each
				const size_t m = (size_t)h[1], bl = (size_t)h[2];
				Channelizer ch(m, bl);
				if (!ch.init() || ch.inputLen() != m * bl || ch.outputLen() != bl) return 4;
				std::vector<float> x(2 * m * bl);
				for (int blk = 0; blk < h[3]; blk++) {
					if (!rd(fi, x.data(), x.size() * 4)) return 2;
					if (!ch.rotate(x.data(), m * bl)) return 4;
					for (size_t c = 0; c < m; c++) put_cf(ch.outputBuffer(c), bl);
				}
				if (ch.rotate(x.data(), m * bl - 1)) return 5; // length mismatch is refused (Channelizer.cpp:79-82)
				break;
			}
			case 3: { // Synthesis(m = a, blockLen = b): c blocks; inputBuffer(chan) filled, rotate() out
				const size_t m = (size_t)h[1], bl = (size_t)h[2];
				Synthesis sy(m, bl);
				if (!sy.init() || sy.inputLen() != bl || sy.outputLen() != m * bl) return 4;
				std::vector<float> y(2 * m * bl);
				for (int blk = 0; blk < h[3]; blk++) {
					for (size_t c = 0; c < m; c++)
						if (!rd(fi, sy.inputBuffer(c), bl * 8)) return 2;
					if (blk == 1) sy.resetBuffer(0); // an inactive channel (radioInterfaceMulti.cpp:329-333)
					if (!sy.rotate(y.data(), m * bl)) return 4;
					put_cf(y.data(), m * bl);
				}
				break;
			}
			case 4: { // scaleVector(x[len = a], scale)
				float sc[2];
				signalVector x((size_t)h[1]);
				if (!rd(fi, sc, 8) || !rd(fi, x.begin(), (size_t)h[1] * 8)) return 2;
				scaleVector(x, complex(sc[0], sc[1]));
				put_cf((const float *)x.begin(), x.size());
				break;
			}
			case 5: { // delayVector(in[len = a], out = (b ? caller's vector : NULL), delay)
				float d;
				signalVector x((size_t)h[1]), own((size_t)h[1]);
				if (!rd(fi, &d, 4) || !rd(fi, x.begin(), (size_t)h[1] * 8)) return 2;
				signalVector *r = delayVector(&x, h[2] ? &own : nullptr, d);
				if (!r || (h[2] && r != &own)) return 4;
				put_cf((const float *)r->begin(), r->size());
				if (!h[2]) delete r;
				break;
			}
			case 6: { // energyDetect(x[len = a], window = b)
				signalVector x((size_t)h[1]);
				if (!rd(fi, x.begin(), (size_t)h[1] * 8)) return 2;
				const float e = energyDetect(x, (unsigned)h[2]);
				int32_t one = 1;
				wr(fo, &one, 4);
				wr(fo, &e, 4);
				wr(fo, &e, 4);
				break;
			}
			case 7: { // vectorSlicer(dst, src, len = a)
				std::vector<float> x((size_t)h[1]), y((size_t)h[1] + (h[1] & 1));
				if (!rd(fi, x.data(), x.size() * 4)) return 2;
				vectorSlicer(y.data(), x.data(), x.size());
				int32_t c = (int32_t)(y.size() / 2);
				wr(fo, &c, 4);
				wr(fo, y.data(), y.size() * 4);
				break;
			}
			case 8: { // filler-burst generators: kind a (0 normal, 1 access, 2 EDGE, 3 dummy, 4 empty), arg b, tn c, sps d
				std::unique_ptr<signalVector> w;
				if (h[1] == 0) w.reset(genRandNormalBurst(h[2], h[4], h[3]));
				else if (h[1] == 1) w.reset(genRandAccessBurst(h[2], h[4], h[3]));
				else if (h[1] == 2) w.reset(generateEdgeBurst(h[2]));
				else if (h[1] == 3) w.reset(generateDummyBurst(h[4], h[3]));
				else w.reset(generateEmptyBurst(h[4], h[3]));
				if (!w) { int32_t c = -1; wr(fo, &c, 4); break; }
				put_cf((const float *)w->begin(), w->size());
				break;
			}
			case 9: { // modulateBurst(bits[a], guard = b, sps = c, emptyPulse = d & 1); d & 2: modulateEdgeBurst
				BitVector bits((size_t)h[1]);
				if (!rd(fi, bits.begin(), (size_t)h[1])) return 2;
				std::unique_ptr<signalVector> w((h[4] & 2) ? modulateEdgeBurst(bits, h[3], h[4] & 1)
								       : modulateBurst(bits, h[2], h[3], h[4] & 1));
				if (!w) { int32_t c = -1; wr(fo, &c, 4); break; }
				put_cf((const float *)w->begin(), w->size());
				break;
			}
			case 10: { // convert: a values, mode b (0 convert_float_short, 1 base_convert_float_short, 2 convert_short_float), scale
				float sc;
				if (!rd(fi, &sc, 4)) return 2;
				const size_t len = (size_t)h[1];
				if (h[2] < 2) {
					std::vector<float> x(len);
					std::vector<short> y(len + (len & 1) + 2, 0);
					if (!rd(fi, x.data(), len * 4)) return 2;
					if (h[2] == 0) convert_float_short(y.data(), x.data(), sc, (int)len);
					else base_convert_float_short(y.data(), x.data(), sc, (int)len);
					int32_t c = (int32_t)len;
					wr(fo, &c, 4);
					wr(fo, y.data(), len * 2);
				} else {
					std::vector<short> x(len);
					std::vector<float> y(len);
					if (!rd(fi, x.data(), len * 2)) return 2;
					convert_short_float(y.data(), x.data(), (int)len);
					int32_t c = (int32_t)len;
					wr(fo, &c, 4);
					wr(fo, y.data(), len * 4);
				}
				break;
			}
			default:
				return 2;
			}
		}
	} else {
		return 2;
	}
	fclose(fi);
	fclose(fo);
	sigProcLibDestroy();
	return 0;
}
