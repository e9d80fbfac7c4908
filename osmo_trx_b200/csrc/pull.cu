// pull.cu — the receive chain around the hot path, so that a slot enters the GPU
// as the radio delivers it (int16 I/Q) and leaves it as the TRXD uplink datagram the BTS consumes:
//
//   ingest_kernel   RadioInterface::pullBuffer's convert_short_float (radioInterface.cpp:345-349,
//                   arch/x86/convert.c:37-79) fused with pullRadioVector's pre-detection power measurement
//                   energyDetect(burst, 20*sps) (Transceiver.cpp:723-731, sigProcLib.cpp:1573-1585) and the
//                   slot-type gate (OFF: no processing :713-716; IDLE: measured but never detected :754).
//   pack_kernel     what follows demodAnyBurst: RSSI (:742-751), vectorSlicer (:803, sigProcLib.cpp:546-556),
//                   idle handling (:810-814) and trxd_send_burst_ind_v0/_v1's header + soft bits normalised
//                   to 0..255 (proto_trxd.c:27-117), written as the exact datagram bytes.
// Detection and demodulation in between are the kernels of detect.cu / demod.cu, unchanged.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

// One warp per slot: 625 short2 in (4-byte coalesced loads), 625 float2 out (8-byte coalesced stores).
// (float)int16 is exact.  energyDetect's sum is sequential in float (80 terms at sample stride 4): the
// terms sit in lanes 0,4,..,28 of ten registers and are folded in order with shuffles.
__global__ void __launch_bounds__(256)
ingest_kernel(IngestParams p)
{
	const int lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < p.n; b += gridDim.x * wpb) {
		const int type = p.type[b];
		const bool off = (type == 0);
		if (lane == 0) {
			// detection sees OFF and IDLE slots as "no attempt" (type 0): pullRadioVector never calls
			// detectAnyBurst for them (Transceiver.cpp:713-716,754-755)
			p.type_out[b] = (type == 6) ? (uint8_t)0 : (uint8_t)type;
			if (off) p.energy[b] = 0.0f;
		}
		if (off) continue;
		const short2 *x = reinterpret_cast<const short2 *>(p.iq) + (size_t)b * p.stride_in;
		float2 *o = reinterpret_cast<float2 *>(p.out) + (size_t)b * p.stride_out;
		short2 s[20];
#pragma unroll
		for (int it = 0; it < 20; it++) {
			const int i = lane + 32 * it;
			s[it] = make_short2(0, 0);
			if (i < 625) s[it] = __ldg(&x[i]);
		}
		float pw[10];
#pragma unroll
		for (int it = 0; it < 20; it++) {
			const int i = lane + 32 * it;
			const float2 v = make_float2((float)s[it].x, (float)s[it].y);
			if (i < 625) o[i] = v;
			if (it < 10) pw[it] = norm2(v);
		}
		float e = 0.0f;
#pragma unroll
		for (int it = 0; it < 10; it++)
#pragma unroll
			for (int j = 0; j < 8; j++)
				e = fa(e, __shfl_sync(0xffffffffu, pw[it], 4 * j));
		if (lane == 0) p.energy[b] = e / 80.0f;
	}
}

namespace {

// `(int32) = double` the way x86-64 cvttsd2si does it (the reference's implicit conversions in proto_trxd.c
// compile to it): truncation, 0x80000000 for NaN and out-of-range values
__device__ __forceinline__ int dbl_to_i32_x86(double v)
{
	if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
	return (int)v; // in range: cvt.rzi
}

constexpr int kPktMax = 11 + 444 + 2;

} // namespace

// One warp per slot.  The datagram is assembled in shared memory and written with 4-byte stores where the
// row allows it.  Rows of slots that emit nothing (OFF, v0 idle, truncated) get pkt_len 0 and are not touched.
__global__ void __launch_bounds__(256)
pack_kernel(PackParams p)
{
	__shared__ __align__(16) uint8_t spk_all[8][464];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int wpb = blockDim.x >> 5;
	uint8_t *spk = spk_all[warp];
	const int hdr = p.version == 1 ? 11 : 8;
	for (int b = blockIdx.x * wpb + warp; b < p.n; b += gridDim.x * wpb) {
		const int type = p.type[b];
		const int rc = p.rc[b];
		const bool idle = !(rc > 0);
		const bool psk8 = (rc == 5);
		const int nbits = idle ? 0 : (psk8 ? 444 : 148);
		int len = 0;
		bool trunc = false;
		if (type != 0 && !(p.version == 0 && idle)) {
			len = hdr + nbits + ((p.version == 0) ? 2 : 0);
			if (p.version == 1 && idle) len = hdr;
			if (len > p.pkt_stride || (!idle && nbits > p.soft_stride)) { len = 0; trunc = true; }
		}
		__syncwarp();
		if (len > 0) {
			if (lane == 0) {
				// trxd_fill_common :27-33
				const uint32_t fn = p.fn[b];
				spk[0] = (uint8_t)((p.tn[b] & 7) | ((p.version & 15) << 4));
				spk[1] = (uint8_t)(fn >> 24); spk[2] = (uint8_t)(fn >> 16); spk[3] = (uint8_t)(fn >> 8); spk[4] = (uint8_t)fn;
				// Transceiver.cpp:742,751 then trxd_fill_v0_specific :35-44
				const float avg = __fsqrt_rn(p.energy[b]);
				const double rssi = 20.0 * log10(p.full_scale / (double)avg) + p.rssi_offset;
				spk[5] = (uint8_t)((unsigned)dbl_to_i32_x86(rssi) & 0xffu);
				const double toa = idle ? 0.0 : (double)p.toa[b];
				const unsigned toa_i = (unsigned)dbl_to_i32_x86(toa * 256.0 + 0.5);
				spk[6] = (uint8_t)(toa_i >> 8); spk[7] = (uint8_t)toa_i;
				if (p.version == 1) {
					// trxd_fill_v1_specific :46-60 (ci * 10 is a float product)
					const float ci = idle ? 0.0f : p.ci[b];
					const unsigned ci_cb = (unsigned)dbl_to_i32_x86((double)fm(ci, 10.0f) + 0.5);
					const int tsc = idle ? 0 : p.tsc_out[b];
					spk[8] = (uint8_t)((tsc & 7) | ((psk8 ? 4 : 0) << 3) | ((idle ? 1 : 0) << 7));
					spk[9] = (uint8_t)(ci_cb >> 8); spk[10] = (uint8_t)ci_cb;
				}
				if (p.version == 0) { spk[hdr + nbits] = 0; spk[hdr + nbits + 1] = 0; }
			}
			// vectorSlicer + trxd_fill_burst_normalized255 :62-67: (uint8_t)round(clamp(0.5*(s+1),0,1) * 255.0)
			const float *srow = p.soft + (size_t)b * p.soft_stride;
			for (int i = lane; i < nbits; i += 32) {
				float v = fm(0.5f, fa(srow[i], 1.0f));
				if (v > 1.0f) v = 1.0f;
				else if (v < 0.0f) v = 0.0f;
				const double x = (double)v * 255.0; // exact (24-bit x 8-bit significands)
				// round(): half away from zero; x >= 0 or NaN here
				spk[hdr + i] = (uint8_t)((unsigned)dbl_to_i32_x86(floor(x + 0.5)) & 0xffu);
			}
			__syncwarp();
			uint8_t *row = p.pkt + (size_t)b * p.pkt_stride;
			if ((reinterpret_cast<uintptr_t>(row) & 3u) == 0) {
				const uint32_t *s4 = reinterpret_cast<const uint32_t *>(spk);
				const int nw = len >> 2;
				for (int j = lane; j < nw; j += 32) reinterpret_cast<uint32_t *>(row)[j] = s4[j];
				if (lane < (len & 3)) row[4 * nw + lane] = spk[4 * nw + lane];
			} else {
				for (int j = lane; j < len; j += 32) row[j] = spk[j];
			}
		}
		if (lane == 0) {
			p.pkt_len[b] = (uint16_t)len;
			if (p.flags && trunc) p.flags[b] |= 8;
		}
	}
}

} // namespace trxb200
