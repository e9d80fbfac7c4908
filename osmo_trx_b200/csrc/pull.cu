// pull.cu — the receive chain around the hot path, so that a slot enters the GPU
// as the radio delivers it (int16 I/Q) and leaves it as the TRXD uplink datagram the BTS consumes:
//
//   extract_kernel  the slot-type gate (OFF: no processing, Transceiver.cpp:713-716; IDLE: measured but never
//                   detected :754) and RadioInterface::pullBuffer's convert_short_float (radioInterface.cpp:345-349,
//                   arch/x86/convert.c:37-79) for the correlator windows only.
//   demod_kernel<true> (demod.cu) reads the int16 slot itself: conversion on the way to the FIR, plus pullRadioVector's
//                   pre-detection power measurement energyDetect(burst, 20*sps) (Transceiver.cpp:723-731,
//                   sigProcLib.cpp:1573-1585) from the staged slot.
//                   It writes the soft bits straight into the datagram rows: vectorSlicer (:803,
//                   sigProcLib.cpp:546-556) + soft bits normalised to 0..255 (proto_trxd.c:62-67).
//   header_kernel   the slot power (ordered sum of the terms demod left), RSSI (:742-751), idle handling
//                   (:810-814) and trxd_send_burst_ind_v0/_v1's header (proto_trxd.c:27-60).
// Detection in between is the float path's corr/peak kernels, unchanged, on the extracted windows.
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {

// extract_kernel: the fused pull chain does not convert whole slots.  Detection reads only the correlator windows
// (152 of 625 samples for a normal burst), so those are converted into a compact float buffer the detection kernels
// address as if it were the full row (pointer shifted by s_min samples, row stride W); demodulation then reads the
// int16 slot itself (demod_kernel<true>), which also measures the slot's power.  One warp per slot.
__global__ void __launch_bounds__(256)
extract_kernel(ExtractParams p)
{
	const int lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < p.n; b += gridDim.x * wpb) {
		const int type = p.type[b];
		// detection sees OFF and IDLE slots as "no attempt" (type 0): pullRadioVector never calls detectAnyBurst
		// for them (Transceiver.cpp:713-716,754-755)
		if (lane == 0) p.type_out[b] = (type == 6) ? (uint8_t)0 : (uint8_t)type;
		if (type == 0 || type == 6) continue;
		const short2 *x = reinterpret_cast<const short2 *>(p.iq) + (size_t)b * p.stride_in;
		float2 *o = reinterpret_cast<float2 *>(p.win) + (size_t)b * p.W;
		for (int i0 = 0; i0 < p.W; i0 += 128) {
			short2 v[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int i = i0 + lane + 32 * k, ns = p.s_min + i;
				v[k] = make_short2(0, 0);
				if (i < p.W && ns >= 0 && ns <= 624) v[k] = __ldg(&x[ns]);
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int i = i0 + lane + 32 * k;
				if (i < p.W) o[i] = make_float2((float)v[k].x, (float)v[k].y);
			}
		}
	}
}

// ---- burst-type scheduler: Transceiver::expectedCorrType (Transceiver.cpp:513-601) and the
// search window pullRadioVector picks from it (:757-758), one thread per slot.  With it the host ships raw slots
// plus (FN, TN, channel) and the per-channel timeslot configuration instead of a CorrType per slot. ----
__constant__ unsigned char c_sd4[102], c_sd8[102]; // SDCCH/4, SDCCH/8 sub-slot per 102-multiframe position

__global__ void __launch_bounds__(256)
sched_kernel(SchedParams p)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= p.n) return;
	enum { CC_FILL, CC_I, CC_II, CC_III, CC_IV, CC_V, CC_VI, CC_VII, CC_VIII, CC_IX, CC_X, CC_XI, CC_XII, CC_XIII, CC_NONE, CC_LOOPBACK };
	const unsigned t = p.tn[i] & 7u, f = p.fn[i];
	const unsigned ch = p.chan ? p.chan[i] : 0u;
	int r = 0; // OFF
	if (ch < (unsigned)p.n_chan) {
		const unsigned ho = p.handover[t];
		const int rach = p.ext_rach ? 2 : 3;
		// half-rate sub-slot of the 26-multiframe position: 0,1 alternating, 0,0 at 12-13, then 1,0 alternating, 1 at 25
		const unsigned f26 = f % 26u;
		const unsigned hsub = f26 < 12u ? (f26 & 1u) : (f26 == 12u ? 0u : (f26 < 25u ? ((f26 & 1u) ^ 1u) : 1u));
		switch (p.chan_type[ch * 8u + t]) {
		case CC_FILL: r = 6; break;
		case CC_I: r = (ho & 1u) ? 3 : 1; break;
		case CC_II: r = hsub == 1u ? 6 : ((ho & 1u) ? 3 : 1); break;
		case CC_III: r = ((ho >> hsub) & 1u) ? 3 : 1; break;
		case CC_IV:
		case CC_VI: r = rach; break;
		case CC_V: {
			const unsigned m = f % 51u;
			if ((m >= 14u && m <= 36u) || m == 4u || m == 5u || m == 45u || m == 46u) r = rach;
			else r = ((ho >> c_sd4[f % 102u]) & 1u) ? 3 : 1;
			break;
		}
		case CC_VII: {
			const unsigned m = f % 51u;
			if (m >= 12u && m <= 14u) r = 6;
			else r = ((ho >> c_sd8[f % 102u]) & 1u) ? 3 : 1;
			break;
		}
		case CC_XIII: {
			const unsigned m = f % 52u;
			if (m == 12u || m == 38u) r = 3; // PTCCH/U: always the 8-bit access burst
			else if (m == 25u || m == 51u) r = 6;
			else r = p.egprs ? 5 : 1;
			break;
		}
		case CC_LOOPBACK: r = (f % 51u >= 48u) ? 6 : 1; break;
		default: r = 0; break; // NONE and the combinations the reference does not schedule
		}
	}
	p.type_out[i] = (uint8_t)r;
	if (p.max_toa_out) p.max_toa_out[i] = (uint16_t)((r == 3 || r == 2) ? p.max_toa_ab : p.max_toa_nb);
}

// header_kernel: what surrounds the soft bytes demod_kernel<true> has already written into the datagram rows.  Lanes =
// slots.  The slot power is the sequential float sum of the 80 terms demod_kernel<true> left in p.pw (a warp's 32 rows
// are brought into shared memory coalesced and summed one row per lane); then RSSI (:742-751, in double as the
// reference), idle handling (:810-814) and trxd_send_burst_ind_v0/_v1's header (proto_trxd.c:27-60).  Rows of slots
// that emit nothing (OFF, v0 idle, truncated) get pkt_len 0 and are not touched.
constexpr int kHdrWarps = 4;
__global__ void __launch_bounds__(kHdrWarps * 32)
header_kernel(HeaderParams p)
{
	__shared__ float pwt[kHdrWarps][32][81]; // odd pitch: a lane walking its row stays on its own bank
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int hdr = p.version == 1 ? 11 : 8;
	const int ntiles = (p.n + 31) >> 5;
	for (int tile = blockIdx.x * kHdrWarps + warp; tile < ntiles; tile += gridDim.x * kHdrWarps) {
		const int b0 = tile * 32;
		const int nb = min(32, p.n - b0);
		__syncwarp();
		// the tile's 32 rows by asynchronous 4-byte copies: all 96 per lane are in flight together (one round trip)
		for (int j = 0; j < nb; j++) {
			const float *row = p.pw + (size_t)(b0 + j) * 80;
#pragma unroll
			for (int k = 0; k < 3; k++)
				if (lane + 32 * k < 80)
					asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&pwt[warp][j][lane + 32 * k])),
						     "l"(row + lane + 32 * k)
						     : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		const int b = b0 + lane;
		if (b >= p.n) continue;
		const int type = p.type[b];
		const int rc = p.rc[b];
		float e = 0.0f;
		if (type != 0) {
#pragma unroll 8
			for (int i = 0; i < 80; i++) e = fa(e, pwt[warp][lane][i]);
			e = e / 80.0f;
		}
		p.energy[b] = e;
		const bool idle = !(rc > 0);
		const bool psk8 = (rc == 5);
		const int nbits = idle ? 0 : (psk8 ? 444 : 148);
		int len = 0;
		bool trunc = false;
		if (type != 0 && !(p.version == 0 && idle)) {
			len = hdr + nbits + ((p.version == 0) ? 2 : 0);
			if (p.version == 1 && idle) len = hdr;
			if (len > p.pkt_stride) { len = 0; trunc = true; }
		}
		if (len > 0) {
			uint8_t h[11];
			// trxd_fill_common :27-33
			const uint32_t fn = p.fn[b];
			h[0] = (uint8_t)((p.tn[b] & 7) | ((p.version & 15) << 4));
			h[1] = (uint8_t)(fn >> 24); h[2] = (uint8_t)(fn >> 16); h[3] = (uint8_t)(fn >> 8); h[4] = (uint8_t)fn;
			// Transceiver.cpp:742,751 then trxd_fill_v0_specific :35-44
			const float avg = __fsqrt_rn(e);
			const double rssi = 20.0 * log10(p.full_scale / (double)avg) + p.rssi_offset;
			h[5] = (uint8_t)((unsigned)dbl_to_i32_x86(rssi) & 0xffu);
			const double toa = idle ? 0.0 : (double)p.toa[b];
			const unsigned toa_i = (unsigned)dbl_to_i32_x86(toa * 256.0 + 0.5);
			h[6] = (uint8_t)(toa_i >> 8); h[7] = (uint8_t)toa_i;
			h[8] = h[9] = h[10] = 0;
			if (p.version == 1) {
				// trxd_fill_v1_specific :46-60 (ci * 10 is a float product)
				const float ci = idle ? 0.0f : p.ci[b];
				const unsigned ci_cb = (unsigned)dbl_to_i32_x86((double)fm(ci, 10.0f) + 0.5);
				const int tsc = idle ? 0 : p.tsc_out[b];
				h[8] = (uint8_t)((tsc & 7) | ((psk8 ? 4 : 0) << 3) | ((idle ? 1 : 0) << 7));
				h[9] = (uint8_t)(ci_cb >> 8); h[10] = (uint8_t)ci_cb;
			}
			uint8_t *row = p.pkt + (size_t)b * p.pkt_stride;
			if ((reinterpret_cast<uintptr_t>(row) & 3u) == 0) {
				reinterpret_cast<uint32_t *>(row)[0] = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)h[3] << 24);
				reinterpret_cast<uint32_t *>(row)[1] = (uint32_t)h[4] | ((uint32_t)h[5] << 8) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 24);
			} else {
#pragma unroll
				for (int k = 0; k < 8; k++) row[k] = h[k];
			}
			if (p.version == 1) { row[8] = h[8]; row[9] = h[9]; row[10] = h[10]; } // byte 11 is the first soft byte
		}
		p.pkt_len[b] = (uint16_t)len;
		if (p.flags && trunc) p.flags[b] |= 8;
	}
}

} // namespace trxb200
