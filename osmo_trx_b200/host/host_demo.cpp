// host_demo.cpp — drives the per-burst C++ mirror the way Transceiver::pullRadioVector / addRadioVector and
// burst-gen.cpp drive the reference (setup once, then per-burst calls); used by tests/test_host_mirror.py.
//   host_demo dd   in.bin out.bin   detectAnyBurst + demodAnyBurst per burst
//   host_demo mod  in.bin out.bin   modulateBurst / modulateEdgeBurst per bit vector
//   host_demo conv in.bin out.bin   convolve_real / convolve_complex / base_* single calls
//   host_demo vit  in.bin out.bin   get_norm_chan_imp_resp + detect_burst_nb per burst
#include "include/sigProcLib.h"
#include "include/convolve.h"
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
	if (argc != 4) { fprintf(stderr, "usage: host_demo dd|sch|mod|conv|vit in out\n"); return 2; }
	const std::string mode = argv[1];
	FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
	if (!fi || !fo) { perror("open"); return 2; }
	if (!sigProcLibSetup()) return 3;
	initvita();
	int32_t n = 0;
	if (!rd(fi, &n, 4)) return 2;
	if (mode == "dd") {
		for (int k = 0; k < n; k++) {
			int32_t hdr[3];
			signalVector burst(625);
			if (!rd(fi, hdr, 12) || !rd(fi, burst.begin(), 625 * 8)) return 2;
			estim_burst_params ebp;
			ebp = estim_burst_params{ complex(0, 0), 0.0f, 0, 0.0f };
			const int rc = detectAnyBurst(burst, (unsigned)hdr[1], BURST_THRESH, 4, (CorrType)hdr[0], (unsigned)hdr[2], &ebp);
			float rec[6] = { (float)rc, ebp.amp.real(), ebp.amp.imag(), ebp.toa, (float)ebp.tsc, ebp.ci };
			std::vector<float> soft(444, 0.0f);
			int32_t nsoft = 0;
			if (rc > 0) {
				std::unique_ptr<SoftVector> sv(demodAnyBurst(burst, (CorrType)rc, 4, &ebp));
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
		// detectSCHBurst(burst, BURST_THRESH, 4, SCH_DETECT_FULL) per burst; the other states must be refused
		for (int k = 0; k < n; k++) {
			signalVector burst(625);
			if (!rd(fi, burst.begin(), 625 * 8)) return 2;
			estim_burst_params ebp;
			ebp = estim_burst_params{ complex(0, 0), 0.0f, 0, 0.0f };
			const int rc = detectSCHBurst(burst, BURST_THRESH, 4, sch_detect_type::SCH_DETECT_FULL, &ebp);
			if (k == 0 && detectSCHBurst(burst, BURST_THRESH, 4, sch_detect_type::SCH_DETECT_BUFFER, &ebp) >= 0) return 5;
			float rec[5] = { (float)rc, ebp.amp.real(), ebp.amp.imag(), ebp.toa, ebp.ci };
			wr(fo, rec, sizeof(rec));
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
	} else {
		return 2;
	}
	fclose(fi);
	fclose(fo);
	sigProcLibDestroy();
	return 0;
}
