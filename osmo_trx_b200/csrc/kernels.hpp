// kernels.hpp — launch parameter blocks shared between the C ABI (capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

namespace trxb200 {

struct DetectParams {
	const float *bursts; // complex[n][stride]
	int stride, n;
	const uint8_t *type, *tsc;
	const uint16_t *max_toa;
	int max_toa_bound;
	float thresh;
	int32_t *rc;
	float *amp, *toa, *ci;
	uint8_t *tsc_out, *flags;
	const float *interp_w;
	int lmax, ndmax; // shared-memory sizing: max correlation length / decimated window
	int scan_clip;	 // 1: run maxAmplitude over the whole burst here; 0: deferred to the demod kernel
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
};

} // namespace trxb200
