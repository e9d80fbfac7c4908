// tables.hpp — host-side construction of every constant table the kernels use.
//
// Replaces sigProcLibSetup() (sigProcLib.cpp:2139-2172) and initvita() (grgsm_vitac.cpp:51-80):
// the tables are computed once on the host with the reference's formulas and precision
// (double libm -> float casts where the reference does so), then uploaded to the device.
// Derived tables that only exist in this implementation (interpolation weights indexed by the
// 1/512-symbol TOA grid, composite delay*decimation filters) are built from those.
#pragma once
#include <cstdint>
#include <vector>

namespace trxb200 {

struct cf { float r, i; };

struct CorrSeq {
	cf seq[64];   // conjugated, unshaped reference (generateMidamble etc.)
	int len = 0;  // 16 (TSC/EDGE/dummy), 40 (RACH), 64 (SCH)
	cf gain{0, 0};
	float toa = 0;
};

constexpr int kSincSize = 1024;	  // TABLESIZE sigProcLib.cpp:45
constexpr int kDelayFilts = 64;	  // DELAYFILTS sigProcLib.cpp:46
constexpr int kDelayTaps = 20;
constexpr int kDecTaps = 16;
constexpr int kInterpSpan = 21;	  // taps floor(ix)-10 .. floor(ix)+10 (sigProcLib.cpp:1102-1105)
constexpr int kInterpGrid = 512;  // TOA bisection resolution 1/512 symbol (sigProcLib.cpp:1162-1173)
constexpr int kCompTaps = 35;	  // 20 + 16 - 1
constexpr int kCompStride = 36;
constexpr int kCompFilts = kDelayFilts + 1; // + identity (no fractional filter, sigProcLib.cpp:1056)

struct HostTables {
	float sinc[kSincSize + 1];
	cf rot4[625], rrot4[625], rot1[157], rrot1[157];
	float pulse4_c0[16], pulse4_c1[8], pulse1_c0[4], c0_inv[5];
	float delay[kDelayFilts][kDelayTaps]; // stored reversed, as convolve consumes them
	float dnsamp[kDecTaps];		      // Resampler(1,4) partition, reversed
	CorrSeq midamble[8], edge_midamble[8], rach[3], sch, dummy;
	cf psk8[8];
	// derived
	std::vector<float> interp_w;  // [kInterpGrid][kInterpSpan]: sinc(pi_f*(d - F/512)), d = -10..10
	std::vector<float> sinc512;   // [11*512]: sinc(pi_f * a/512) as interpolatePoint sees it, a = |512*(d-10) - F| (same values as interp_w)
	std::vector<float> comp;      // [kCompFilts][16 kmin][kCompStride]: truncated composite filters
	cf edge_derot[16];	      // (cosf, -sinf) of (i%16)*3pi/8  (sigProcLib.cpp:703-704)
	cf edge_ideal[9];	      // (cos, sin)(k*pi/4), k=-4..4 as computeEdgeCI evaluates them (:2082-2083)
	cf edge_rot1, edge_rot2;      // rotateBurst2(-pi/8), (-pi/4) (:582-588,1975,1994)
	cf edge_mod_rot[156];	      // e^{j i 3pi/8} as shapeEdgeBurst evaluates it (:754-755)
	// grgsm_vitac reference symbol sequences (conjugated), grgsm_vitac.cpp:46-80
	cf vitac_norm[9][26], vitac_access[41], vitac_sch[64];
};

// Build everything. Deterministic; no device interaction.
void build_host_tables(HostTables &t);

// Resampler / Channelizer prototype filters (Resampler.cpp:47-96, ChannelizerBase.cpp:68-132).
// parts: [p][filt_len] real taps stored reversed.
void build_resampler_taps(int p, int q, int filt_len, float bw, std::vector<float> &parts);
void build_channelizer_taps(int m, int h_len, std::vector<float> &parts);

// Host reference FIR used only during table construction (tiny vectors).
void host_conv_real(const cf *x, int nx, const float *h, int nh, int start, int len, cf *y);
void host_conv_cplx(const cf *x, int nx, const cf *h, int nh, int start, int len, cf *y);

} // namespace trxb200
