// device_tables.cuh — constant-memory tables shared by the kernels, and exact-arithmetic helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace trxb200 {

// sequence ids
constexpr int SEQ_MIDAMBLE = 0;	 // 0..7
constexpr int SEQ_EDGE = 8;	 // 8..15
constexpr int SEQ_RACH = 16;	 // 16..18
constexpr int SEQ_SCH = 19;
constexpr int SEQ_DUMMY = 20;
constexpr int SEQ_COUNT = 21;
constexpr int SEQ_STORE = 17 * 16 + 3 * 40 + 64;

struct SeqInfo {
	int off, len;	 // into c_seq
	float inv_gr, inv_gi; // gain.inv() (Complex.h:144-150): amp = xcorr * inv(gain)  (sigProcLib.cpp:1701)
	float ci_den;	 // (N-1) * |gain|   (sigProcLib.cpp:1629)
	float toa;	 // sequence time offset (sigProcLib.cpp:1704)
	int pm1;	 // 1: every even tap is (+-1, tiny) and every odd tap (tiny, +-1) - the rotated GMSK sequences (corr_long_items_pm1)
	int pad_;
};

struct ConstTables {
	float2 seq[SEQ_STORE];
	SeqInfo info[SEQ_COUNT];
	float dnsamp[16];
	float pulse_c0[16], pulse_c1[8];
	float c0_inv[5];
	float2 rrot1[160];
	float2 rot4[628];
	float2 psk8[8];
	float2 edge_mod_rot[156];
	float2 edge_derot[16];
	float2 edge_ideal[9];
	float2 edge_rot1, edge_rot2;
	float delay[64][20];
	float2 vitac_norm[9][26];
	float2 vitac_access[41];
	float2 vitac_sch[64];
	float comp0[65][2][36]; // untruncated composite demod filters, shifted by e: comp0[f][e][u] = comp[f][0][u - e]
	float2 rot1[160];	// e^{+j*pi*n/2}, the 1-sps GMSK rotator (GMSKRotation1, sigProcLib.cpp:191-216)
	float pulse1_c0[4];	// GSMPulse1->c0 (:523-532)
};

// layout of the modulator table block kept in global memory (indexed per lane, so not __constant__)
constexpr int kModRot4 = 0, kModC0 = 1250, kModC1 = 1266, kModEdgeRot = 1274, kModPsk8 = 1586, kModTabFloats = 1602;

// Pointers to the larger tables kept in global memory (L1/L2 resident).
struct GlobalTables {
	const float *interp_w; // [512][21]
	const float *comp;     // [65][16][36]
};

// ---- exact float32 helpers: no contraction, reference operation order ----
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
// Complex::norm2() = i*i + r*r  (Complex.h:112)
__device__ __forceinline__ float norm2(float2 v) { return fa(fm(v.y, v.y), fm(v.x, v.x)); }
// Complex * Complex (Complex.h:73)
__device__ __forceinline__ float2 cmul_exact(float2 a, float2 b)
{
	return make_float2(fs(fm(a.x, b.x), fm(a.y, b.y)), fa(fm(a.x, b.y), fm(a.y, b.x)));
}


// `(int32) = double` the way x86-64 cvttsd2si does it (the reference's implicit conversions in proto_trxd.c compile
// to it): truncation, 0x80000000 for NaN and out-of-range values
__device__ __forceinline__ int dbl_to_i32_x86(double v)
{
	if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
	return (int)v; // in range: cvt.rzi
}

// vectorSlicer (sigProcLib.cpp:546-556) + trxd_fill_burst_normalized255 (proto_trxd.c:62-67):
// (uint8_t)round(clamp(0.5 * (s + 1), 0, 1) * 255.0); round() is half away from zero and the argument is >= 0 or NaN
__device__ __forceinline__ unsigned soft_to_u8(float s)
{
	float v = fm(0.5f, fa(s, 1.0f));
	if (v > 1.0f) v = 1.0f;
	else if (v < 0.0f) v = 0.0f;
	// The reference evaluates round(v * 255.0) in double, where the product is exact.  The same integer in FP32: a
	// candidate from the rounded product, then the two neighbouring half-integer boundaries are tested with a fused
	// multiply-add, whose sign is that of the exact v * 255 - boundary (k +- 0.5 are exact floats).
	int k = __float2int_rd(fmaf(v, 255.0f, 0.5f));
	if (fmaf(v, 255.0f, -((float)k + 0.5f)) >= 0.0f) k++;
	else if (fmaf(v, 255.0f, -((float)k - 0.5f)) < 0.0f) k--;
	return (v == v) ? ((unsigned)k & 0xffu) : 0u; // NaN: cvttsd2si gives 0x80000000, whose low byte is 0
}

} // namespace trxb200
