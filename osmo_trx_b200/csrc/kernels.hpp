// kernels.hpp — launch parameter blocks shared between the C ABI (capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

namespace trxb200 {

// detection = rounds of (corr_kernel, peak_kernel); all pointers are already offset to the chunk
struct CorrParams {
	const float *bursts; // complex[n][stride]
	int stride, n;
	const uint8_t *type, *tsc;
	const uint16_t *max_toa;
	const int32_t *rc; // state left by earlier rounds (read when round > 0)
	int round;	   // attempt index (EDGE -> TSC fall-through, EXT_RACH sequences)
	int max_toa_bound;
	int lmax, ndmax; // row lengths of the intermediates: max correlation length / decimated window
	float2 *corr;	 // [n][lmax]
	float *pwr;	 // [n][ndmax] |decimated sample|^2 (computeCI)
	float negzero;	 // -0.0f at run time (see mul2 in detect.cu)
	// pull chain (corr_nb_kernel<true>): the windows are read from the int16 slots instead of `bursts`
	const int16_t *iq; // [n][iq_stride] complex int16, or null
	int iq_stride;
	int sch;	   // every burst is a SCH_DETECT_FULL search (type / tsc / max_toa are not read)
	const uint4 *seq_pm1 = nullptr; // [SEQ_STORE] per tap of the sequences flagged SeqInfo::pm1: sign masks and the tiny component (corr_long_items_pm1)
	int sps1_len = 0;  // > 0: the rows are bursts of this many samples at ONE sample per symbol - no decimation, the correlator reads
			   // the burst itself (detectBurst, sigProcLib.cpp:1659-1662); corr_long_kernel only
};

struct PeakParams {
	int n;
	const uint8_t *type, *tsc;
	const uint16_t *max_toa;
	int round, last_round;
	int max_toa_bound;
	float thresh;
	int lmax, ndmax;
	const float2 *corr;
	const float *pwr;
	const float *sinc512; // [5632] table-sinc on the 1/512 TOA grid
	int32_t *rc;
	float *amp, *toa, *ci;
	uint8_t *tsc_out, *flags;
	float negzero;
	int sch; // every burst is a SCH_DETECT_FULL search (type / tsc / max_toa are not read)
	int dec_size = 156; // samples of the decimated burst computeCI may read (15000 for the SCH buffer search)
};

struct DemodParams {
	const float *bursts;
	int stride, n;
	int32_t *rc;	   // read; rewritten to -SIGERR_CLIP when fix_clip and the burst clips undetected
	const float *amp, *toa;
	float *ci;
	uint8_t *flags;	   // may be null
	float *soft;
	int soft_stride, n_gmsk_soft;
	const float *comp; // [65][16][36]
	const float *dnsamp_g; // [16] decimator taps in global memory (per-lane indexed)
	const float2 *edge_tab; // [16] derotation (cosf,-sinf)((i%16)*3pi/8) then [9] ideal 8-PSK points k=-4..4
	int fix_clip;
	const uint8_t *type; // burst types as detection saw them (clip report only for types detectAnyBurst handles); may be null
	// pull path (demod_kernel<true>): the slots as the radio delivered them, and the per-slot power measurement
	const int16_t *iq;	 // [n][iq_stride] complex int16 (I,Q); replaces `bursts`
	int iq_stride;
	const uint8_t *type_raw; // slot types as scheduled by the caller (0 = off: no measurement)
	float *pw;		 // [n][80] |x[4i]|^2, the terms of energyDetect(slot, 80); summed in order by header_kernel
	uint8_t *pkt;		 // [n][pkt_stride] TRXD datagram rows: the soft bytes are written here (header_kernel adds the header)
	int pkt_stride, pkt_hdr, pkt_v0; // header length (8 / 11), v0 = two trailing zero bytes
	int sps1_len = 0; // demod1_kernel: `bursts` holds the delayed bursts of this many samples at one sample per symbol
};

// receive chain around the hot path (pull.cu)
// pull path, detection input: per slot the part of the slot the correlators read, as complex float
struct ExtractParams {
	const int16_t *iq; // [n][stride_in] complex int16
	int stride_in, n;
	const uint8_t *type; // CorrType per slot as scheduled by the caller
	uint8_t *type_out;   // type as detection must see it (OFF / IDLE -> 0)
	float *win;	     // [n][W] complex float: samples s_min .. s_min + W - 1 of each slot
	int W, s_min;
};

// burst-type scheduler (pull.cu)
struct SchedParams {
	int n, n_chan;
	const uint32_t *fn;
	const uint8_t *tn;
	const uint16_t *chan;	   // channel index per slot, or null (channel 0)
	const uint8_t *chan_type;  // [n_chan][8] ChannelCombination per timeslot
	const uint8_t *handover;   // [8] per timeslot: bit s = handover expected on sub-slot s
	int ext_rach, egprs, max_toa_nb, max_toa_ab;
	uint8_t *type_out;
	uint16_t *max_toa_out;	   // may be null
};

struct HeaderParams {
	int n, version;
	const uint8_t *type; // caller's slot types (OFF emits nothing)
	const int32_t *rc;
	const float *toa, *ci;
	const float *pw;     // [n][80] terms of energyDetect, written by demod_kernel<true>
	float *energy;	     // out: energyDetect(slot, 80)
	const uint8_t *tsc_out;
	const uint32_t *fn;
	const uint8_t *tn;
	double full_scale, rssi_offset;
	uint8_t *pkt;
	int pkt_stride;
	uint16_t *pkt_len;
	uint8_t *flags; // may be null
};

} // namespace trxb200
