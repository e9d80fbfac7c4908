// tables.cpp — see tables.hpp.  Host-only; runs once per context.
//
// Precision notes (what has to be kept to stay bit-compatible with the reference tables):
//  * trig for the GMSK rotators, windows and sinc table is evaluated in double and narrowed
//    (sigProcLib.cpp:199-214,984-986,1030-1033); the 8-PSK rotators take a *float* phase and
//    therefore the float libm overloads (sigProcLib.cpp:681-682,754-755);
//  * the sinc lookup floors an index computed in double from a float argument (sigProcLib.cpp:990-998);
//  * FIR sums made during setup follow the SSE3 lane order of arch/x86/convolve_sse_3.c.
#include "tables.hpp"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <algorithm>

namespace trxb200 {
namespace {

const float kPiF = (float)M_PI;

inline cf mul(cf a, cf b) { return { a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r }; }
inline float pw(cf a) { return a.i * a.i + a.r * a.r; }

// 3GPP TS 45.002 bit patterns as listed in GSM/GSMCommon.cpp:35-68
const char *const kTsc[8] = {
	"00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110",
	"01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
	"10100111110110001010011111", "11101111000100101110111100" };
const char *const kEdgeTsc[8] = {
	"111111001111111001111001001001111111111111001111111111001111111001111001001001",
	"111111001111001001111001001001111001001001001111111111001111001001111001001001",
	"111001111111111111001001001111001001001111001111111001111111111111001001001111",
	"111001111111111001001001001111001001111001111111111001111111111001001001001111",
	"111111111001001111001111001001001111111001111111111111111001001111001111001001",
	"111001111111001001001111001111001001111111111111111001111111001001001111001111",
	"001111001111111001001001001001111001001111111111001111001111111001001001001001",
	"001001001111001001001001111111111001111111001111001001001111001001001001111111" };
const char *const kDummyTsc = "01110001011100010111000101";
const char *const kRachSync[3] = { "01001011011111111001100110101010001111000",
				   "01010100111110001000011000101111001001101",
				   "11101111001001110101011000001101101110111" };
const char *const kSchSync = "1011100101100010000001000000111100101101010001010111011000011011";
// grgsm_vitac/constants.h:113-125 (row 1 deviates from kTsc[1] at bit 20 in the reference; kept)
const char *const kVitacTrain[9] = {
	"00100101110000100010010111", "00101101110111100010010111", "01000011101110100100001110",
	"01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
	"10100111110110001010011111", "11101111000100101110111100", "01110001011100010111000101" };

std::vector<uint8_t> bits_of(const char *s)
{
	std::vector<uint8_t> v;
	for (; *s; ++s)
		v.push_back(*s == '1');
	return v;
}

float sinc_lookup(const HostTables &t, float x)
{
	float ax = std::fabs(x);
	if ((double)ax >= 8 * M_PI)
		return 0.0f;
	float fidx = (float)((double)ax / (8 * M_PI) * kSincSize);
	return t.sinc[(int)std::floor(fidx)];
}

// SSE3 lane-ordered sums (arch/x86/convolve_sse_3.c), enough variants for the setup-time shapes.
float lane_sum(const float *p, int n)
{
	float L[4];
	for (int j = 0; j < 4; j++) {
		switch (n) {
		case 4: L[j] = p[j]; break;
		case 8: L[j] = p[j] + p[4 + j]; break;
		case 16: L[j] = (p[j] + p[4 + j]) + (p[8 + j] + p[12 + j]); break;
		case 20: L[j] = ((p[j] + p[4 + j]) + p[8 + j]) + (p[12 + j] + p[16 + j]); break;
		default: {
			float a = 0.0f;
			for (int g = 0; g < n / 4; g++)
				a = a + p[4 * g + j];
			L[j] = a;
		}
		}
	}
	return (L[0] + L[1]) + (L[2] + L[3]);
}

} // namespace

void host_conv_real(const cf *x, int nx, const float *h, int nh, int start, int len, cf *y)
{
	float pr[64], pi[64];
	for (int i = 0; i < len; i++) {
		for (int k = 0; k < nh; k++) {
			int idx = i + start - (nh - 1) + k;
			cf v = (idx >= 0 && idx < nx) ? x[idx] : cf{ 0, 0 };
			pr[k] = v.r * h[k];
			pi[k] = v.i * h[k];
		}
		y[i] = { lane_sum(pr, nh), lane_sum(pi, nh) };
	}
}

void host_conv_cplx(const cf *x, int nx, const cf *h, int nh, int start, int len, cf *y)
{
	// sse_conv_cmplx_8n order (all setup-time complex filters have 8n taps)
	for (int i = 0; i < len; i++) {
		float Lr[4], Li[4];
		for (int j = 0; j < 4; j++) {
			float ar = 0, ai = 0, br = 0, bi = 0;
			for (int g = 0; g < nh / 8; g++) {
				for (int half = 0; half < 2; half++) {
					int k = 8 * g + 4 * half + j;
					int idx = i + start - (nh - 1) + k;
					cf v = (idx >= 0 && idx < nx) ? x[idx] : cf{ 0, 0 };
					float re = h[k].r * v.r - h[k].i * v.i;
					float im = h[k].r * v.i + h[k].i * v.r;
					if (half == 0) { ar = ar + re; ai = ai + im; }
					else { br = br + re; bi = bi + im; }
				}
			}
			Lr[j] = ar + br;
			Li[j] = ai + bi;
		}
		y[i] = { (Lr[0] + Lr[1]) + (Lr[2] + Lr[3]), (Li[0] + Li[1]) + (Li[2] + Li[3]) };
	}
}

namespace {

// 1 sps helpers used to derive correlation gains (modulateBurstBasic / rotateBurst at sps 1)
std::vector<cf> rotate_only(const HostTables &t, const uint8_t *b, int n)
{
	std::vector<cf> v(n);
	for (int i = 0; i < n; i++) {
		cf s{ (float)(2.0 * (b[i] & 1) - 1.0), 0.0f };
		cf r = mul(t.rot1[i], s);
		v[i] = { 0.0f + r.r * 1.0f, 0.0f + r.i * 1.0f };
	}
	return v;
}

std::vector<cf> shape_1sps(const HostTables &t, const uint8_t *b, int n)
{
	std::vector<cf> v(n), y(n);
	for (int i = 0; i < n; i++) {
		float s = (float)(2.0 * (b[i] & 1) - 1.0);
		v[i] = { t.rot1[i].r * s, t.rot1[i].i * s };
	}
	host_conv_real(v.data(), n, t.pulse1_c0, 4, 0, n, y.data());
	return y;
}

cf interp_at(const HostTables &t, const cf *x, int n, float ix)
{
	int lo = (int)(std::floor(ix) - 10), hi = (int)(std::floor(ix) + 11);
	if (lo < 0) lo = 0;
	if ((unsigned)hi > (unsigned)(n - 1)) hi = n - 1;
	cf acc{ 0, 0 };
	for (int i = lo; i < hi; i++) {
		float w = sinc_lookup(t, kPiF * (i - ix));
		acc.r += x[i].r * w;
		acc.i += x[i].i * w;
	}
	return acc;
}

// early-late peak refinement used at setup to measure sequence gain/toa (sigProcLib.cpp:1141-1186)
cf refine_peak(const HostTables &t, const cf *x, int n, float *where)
{
	float best = 0.0f, at = -1;
	for (int i = 0; i < n; i++)
		if (pw(x[i]) > best) { best = pw(x[i]); at = (float)i; }
	float early = at - 1, late = at + 1, step = 0.5f;
	while (step > 1.0 / 1024.0) {
		float pe = pw(interp_at(t, x, n, early)), pl = pw(interp_at(t, x, n, late));
		if (pe < pl) early += step;
		else if (pe > pl) early -= step;
		else break;
		step /= 2.0;
		late = early + 2.0;
	}
	*where = early + 1.0;
	return interp_at(t, x, n, *where);
}

void make_seq(const HostTables &t, CorrSeq &cs, const std::vector<uint8_t> &full, int ref_off, int ref_len,
	      bool midamble, double toa_bias)
{
	std::vector<cf> ref = rotate_only(t, full.data() + ref_off, ref_len);
	std::vector<cf> rx = shape_1sps(t, full.data(), (int)full.size());
	if (midamble) {
		for (auto &v : ref) v = mul(v, cf{ -1.0f, 0.0f });
		for (auto &v : rx) v = mul(v, cf{ 0.0f, 1.0f });
	}
	for (auto &v : ref) v.i = -v.i;
	cs.len = ref_len;
	std::copy(ref.begin(), ref.end(), cs.seq);
	std::vector<cf> ac(rx.size());
	host_conv_cplx(rx.data(), (int)rx.size(), ref.data(), ref_len, ref_len / 2, (int)rx.size(), ac.data());
	float where;
	cs.gain = refine_peak(t, ac.data(), (int)ac.size(), &where);
	cs.toa = where - toa_bias;
}

void vitac_map(const uint8_t *in, int n, cf *out, cf first)
{
	out[0] = first;
	int prev = 2 * in[0] - 1;
	for (int i = 1; i < n; i++) {
		int cur = 2 * in[i] - 1;
		cf e{ (float)(cur * prev), 0.0f };
		out[i] = mul(mul(cf{ 0.0f, 1.0f }, e), out[i - 1]);
		prev = cur;
	}
	for (int i = 0; i < n; i++)
		out[i].i = -out[i].i;
}

float bh_window(int n, int N, float a0, float a1, float a2, float a3)
{
	return (float)(a0 - a1 * cos(2 * M_PI * n / (N - 1)) + a2 * cos(4 * M_PI * n / (N - 1)) -
		       a3 * cos(6 * M_PI * n / (N - 1)));
}

} // namespace

void build_resampler_taps(int p, int q, int filt_len, float bw, std::vector<float> &parts)
{
	const int L = p * filt_len;
	std::vector<float> proto(L);
	const float a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
	const float cutoff = p > q ? (float)p : (float)q;
	const float mid = (L - 1) / 2.0;
	float sum = 0.0f;
	for (int i = 0; i < L; i++) {
		float x = ((float)i - mid) / cutoff * bw;
		float s = (x == 0.0) ? (float)0.9999999999 : (float)(sin(M_PI * x) / (M_PI * x));
		double w = a0 - a1 * cos(2 * M_PI * i / (L - 1)) + a2 * cos(4 * M_PI * i / (L - 1)) -
			   a3 * cos(6 * M_PI * i / (L - 1));
		proto[i] = (float)(s * w);
		sum += proto[i];
	}
	const float scale = p / sum;
	parts.assign((size_t)p * filt_len, 0.0f);
	for (int i = 0; i < filt_len; i++)
		for (int n = 0; n < p; n++)
			parts[(size_t)n * filt_len + (filt_len - 1 - i)] = proto[i * p + n] * scale;
}

void build_channelizer_taps(int m, int h_len, std::vector<float> &parts)
{
	const size_t L = (size_t)m * h_len;
	std::vector<float> proto(L);
	const float a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
	const float mid = (float)(L - 1.0) / 2.0;
	float sum = 0.0f;
	for (size_t i = 0; i < L; i++) {
		float x = ((float)i - mid) / (float)m;
		float s = (x == 0.0f) ? 0.999999999999f : (float)(sin(M_PI * x) / (M_PI * x));
		double w = a0 - a1 * cos(2 * M_PI * i / (L - 1)) + a2 * cos(4 * M_PI * i / (L - 1)) -
			   a3 * cos(6 * M_PI * i / (L - 1));
		proto[i] = (float)(s * w);
		sum += proto[i];
	}
	const float scale = (float)m / sum;
	parts.assign(L, 0.0f);
	for (int i = 0; i < h_len; i++)
		for (int n = 0; n < m; n++)
			parts[(size_t)n * h_len + (h_len - 1 - i)] = proto[(size_t)i * m + n] * scale;
}

void build_host_tables(HostTables &t)
{
	// sinc table
	for (int i = 0; i < kSincSize; i++) {
		double x = (double)i / kSincSize * 8 * M_PI;
		double y = sin(x) / x;
		t.sinc[i] = std::isnan(y) ? 1.0 : y;
	}
	t.sinc[kSincSize] = 0.0f;

	// GMSK rotators: phase accumulated in double
	double ph = 0.0;
	for (int i = 0; i < 625; i++, ph += M_PI / 2.0 / 4.0) {
		t.rot4[i] = { (float)cos(ph), (float)sin(ph) };
		t.rrot4[i] = { (float)cos(-ph), (float)sin(-ph) };
	}
	ph = 0.0;
	for (int i = 0; i < 157; i++, ph += M_PI / 2.0) {
		t.rot1[i] = { (float)cos(ph), (float)sin(ph) };
		t.rrot1[i] = { (float)cos(-ph), (float)sin(-ph) };
	}

	// Laurent pulses, EDGE equaliser (literal constants of the reference)
	const double c0[16] = { 0.0, 4.46348606e-03, 2.84385729e-02, 1.03184855e-01, 2.56065552e-01, 4.76375085e-01,
				7.05961177e-01, 8.71291644e-01, 9.29453645e-01, 8.71291644e-01, 7.05961177e-01,
				4.76375085e-01, 2.56065552e-01, 1.03184855e-01, 2.84385729e-02, 4.46348606e-03 };
	const double c1[8] = { 0.0, 8.16373112e-03, 2.84385729e-02, 5.64158904e-02, 7.05463553e-02, 5.64158904e-02,
			       2.84385729e-02, 8.16373112e-03 };
	const double eq[5] = { 0.15884, -0.43176, 1.00000, -0.42608, 0.14882 };
	for (int i = 0; i < 16; i++) t.pulse4_c0[i] = (float)c0[i];
	for (int i = 0; i < 8; i++) t.pulse4_c1[i] = (float)c1[i];
	for (int i = 0; i < 5; i++) t.c0_inv[i] = (float)eq[i];
	{
		const float center = (float)(4 - 1.0) / 2.0;
		float e = 0.0f;
		for (int i = 0; i < 4; i++) {
			float a = ((float)i - center) / 1.0f;
			t.pulse1_c0[i] = (float)(0.96 * exp(-1.1380 * a * a - 0.527 * a * a * a * a));
			e += 0.0f * 0.0f + t.pulse1_c0[i] * t.pulse1_c0[i];
		}
		const float norm = std::sqrt(e / 1);
		for (int i = 0; i < 4; i++) t.pulse1_c0[i] /= norm;
	}

	// correlation sequences (all at 1 sps)
	for (int i = 0; i < 3; i++)
		make_seq(t, t.rach[i], bits_of(kRachSync[i]), 0, 40, false, 20.5);
	make_seq(t, t.sch, bits_of(kSchSync), 0, 64, false, 32.5);
	make_seq(t, t.dummy, bits_of(kDummyTsc), 5, 16, true, 13.5);
	const cf psk8[8] = { { -0.70710678f, 0.70710678f }, { 0.0f, -1.0f }, { 0.0f, 1.0f }, { 0.70710678f, -0.70710678f },
			     { -1.0f, 0.0f }, { -0.70710678f, -0.70710678f }, { 0.70710678f, 0.70710678f }, { 1.0f, 0.0f } };
	std::memcpy(t.psk8, psk8, sizeof(psk8));
	for (int i = 0; i < 156; i++) {
		float phase = i * 3.0f * M_PI / 8.0f;
		t.edge_mod_rot[i] = { std::cos(phase), std::sin(phase) };
	}
	for (int k = 0; k < 8; k++) {
		make_seq(t, t.midamble[k], bits_of(kTsc[k]), 5, 16, true, 13.5);
		std::vector<uint8_t> eb = bits_of(kEdgeTsc[k]);
		CorrSeq &cs = t.edge_midamble[k];
		cs.len = 16;
		for (int s = 0; s < 16; s++) {
			const uint8_t *b = eb.data() + 15 + 3 * s;
			cf sym = t.psk8[(b[0] & 1) | ((b[1] & 1) << 1) | ((b[2] & 1) << 2)];
			cf v = mul(sym, t.edge_mod_rot[s]);
			cs.seq[s] = { v.r, -v.i };
		}
		const float div = 1.18;
		cs.gain = { (float)-19.6432 / div, (float)19.5006 / div };
		cs.toa = 0;
	}

	// fractional delay bank
	for (int f = 0; f < kDelayFilts; f++) {
		float sum = 0.0f;
		for (int n = 0; n < kDelayTaps; n++) {
			float k = (float)n;
			float s = sinc_lookup(t, (float)(kPiF * (k - (float)kDelayTaps / 2.0 - (float)f / kDelayFilts)));
			float v = s * bh_window(n, kDelayTaps, 0.35875, 0.48829, 0.14128, 0.01168);
			t.delay[f][kDelayTaps - 1 - n] = v;
			sum += v;
		}
		for (int n = 0; n < kDelayTaps; n++)
			t.delay[f][n] /= sum;
	}

	std::vector<float> dn;
	build_resampler_taps(1, 4, kDecTaps, 1.0f, dn);
	std::copy(dn.begin(), dn.end(), t.dnsamp);

	// interpolation weights on the 1/512 grid
	t.interp_w.assign((size_t)kInterpGrid * kInterpSpan, 0.0f);
	for (int F = 0; F < kInterpGrid; F++)
		for (int d = -10; d <= 10; d++) {
			float diff = (float)d - (float)F / (float)kInterpGrid; // exact in float
			t.interp_w[(size_t)F * kInterpSpan + (d + 10)] = sinc_lookup(t, kPiF * diff);
		}

	// the same weights indexed by the distance a = |512*(d-10) - F| on the 1/512 grid (the lookup depends on |x| only);
	// every (F, d) pair must land on a consistent entry
	t.sinc512.assign((size_t)11 * kInterpGrid, 0.0f);
	{
		std::vector<char> seen(t.sinc512.size(), 0);
		for (int F = 0; F < kInterpGrid; F++)
			for (int d = 0; d <= 20; d++) {
				int a = 512 * (d - 10) - F;
				if (a < 0) a = -a;
				const float w = t.interp_w[(size_t)F * kInterpSpan + d];
				if (seen[a] && t.sinc512[a] != w) {
					fprintf(stderr, "trxb200: sinc512 inconsistency at F=%d d=%d\n", F, d);
					abort();
				}
				seen[a] = 1;
				t.sinc512[a] = w;
			}
	}

	// composite (fractional delay (*) decimator) filters, truncated from below for the leading outputs
	t.comp.assign((size_t)kCompFilts * 16 * kCompStride, 0.0f);
	for (int f = 0; f < kCompFilts; f++) {
		double h[kDelayTaps];
		for (int j = 0; j < kDelayTaps; j++)
			h[j] = f < kDelayFilts ? (double)t.delay[f][j] : (j == 9 ? 1.0 : 0.0);
		for (int kmin = 0; kmin < 16; kmin++) {
			double acc[kCompTaps] = { 0 };
			for (int k = kmin; k < kDecTaps; k++)
				for (int j = 0; j < kDelayTaps; j++)
					acc[k + j] += (double)t.dnsamp[k] * h[j];
			float *dst = &t.comp[((size_t)f * 16 + kmin) * kCompStride];
			for (int u = 0; u < kCompTaps; u++) dst[u] = (float)acc[u];
		}
	}

	// EDGE demod constants
	for (int i = 0; i < 16; i++) {
		float phase = (float)(i % 16) * 3.0f * M_PI / 8.0f;
		t.edge_derot[i] = { cosf(phase), -sinf(phase) };
	}
	{
		const float step = 2.0f * kPiF / 8.0f;
		for (int k = -4; k <= 4; k++) {
			float phase = step * (float)k;
			t.edge_ideal[k + 4] = { std::cos(phase), std::sin(phase) };
		}
	}
	t.edge_rot1 = { (float)cos(-M_PI / 8.0), (float)sin(-M_PI / 8.0) };
	t.edge_rot2 = { (float)cos(-M_PI / 4.0), (float)sin(-M_PI / 4.0) };

	// vitac
	{
		std::vector<uint8_t> s = bits_of(kSchSync);
		vitac_map(s.data(), 64, t.vitac_sch, cf{ 0.0f, -1.0f });
		std::vector<uint8_t> a = bits_of(kRachSync[0]);
		vitac_map(a.data(), 41, t.vitac_access, cf{ 0.0f, -1.0f });
		for (int k = 0; k < 9; k++) {
			std::vector<uint8_t> b = bits_of(kVitacTrain[k]);
			vitac_map(b.data(), 26, t.vitac_norm[k], cf{ b[0] == 0 ? 1.0f : -1.0f, 0.0f });
		}
	}
}

} // namespace trxb200
