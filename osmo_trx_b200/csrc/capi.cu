// capi.cu — the C ABI declared in include/trxb200.h: context, table upload, kernel launches and
// the host-buffer pipeline.  This translation unit includes the kernel sources so that all kernels
// share one __constant__ table block (no relocatable device code needed).
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <new>

#include "../../include/trxb200.h"
#include "tables.hpp"
#include "device_tables.cuh"
#include "kernels.hpp"

namespace trxb200 {
__constant__ ConstTables c_tab;
}

#include "detect.cu"
#include "demod.cu"
#include "detect_lane.cu"
#include "nbfused.cu"
#include "modulate.cu"
#include "convolve.cu"
#include "misc.cu"
#include "vitac.cu"
#include "vitac_lane.cu"
#include "filterbank.cu"
#include "resampler_pq.cu"
#include "pull.cu"

using namespace trxb200;

struct HostStage; // pinned staging for the *_host entry points

// intermediates of the detection kernels (corr_kernel -> peak_kernel), sized for one chunk of bursts
struct DetectScratch {
	float2 *corr = nullptr; // [cap][lmax]
	float *pwr = nullptr;	// [cap][ndmax]
	size_t corr_bytes = 0, pwr_bytes = 0;
	int *list = nullptr; // detect_lane_kernel's retry lists: two counters (32 bytes apart), then two lists of list_cap bursts
	size_t list_cap = 0;
};

// device scratch of the pull path (int16 slots -> TRXD datagrams) for one chunk of slots
struct PullScratch {
	int cap = 0, soft_stride = 0;
	float *bursts = nullptr; // [cap][win] complex float: the correlator windows of the slots (extract_kernel)
	int win = 0;
	float *pw = nullptr;	 // [cap][80] terms of energyDetect (demod_kernel<true> -> header_kernel)
	uint8_t *type2 = nullptr, *tsc_out = nullptr;
	float *amp = nullptr, *toa = nullptr, *ci = nullptr;
	DetectScratch ws;
};

struct PullStage; // pinned-host pipeline state of trxb200_pull_host

struct trxb200_ctx {
	int device = -1;
	int sm_count = 0;
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	HostTables *ht = nullptr;
	float *d_sinc512 = nullptr;
	float *d_comp = nullptr;
	uint4 *d_seq_pm1 = nullptr; // per tap of every sequence: sign masks of the +-1 component, the tiny component (corr_long_items_pm1)
	float2 *d_edge_tab = nullptr; // derotation + ideal-symbol tables for the EDGE demodulator
	void *d_tmap = nullptr;	      // TMA descriptors of detect_lane_kernel (four slots)
	float *d_mod_tab = nullptr;   // modulator tables in global memory (per-lane indexed): rot4 | c0 | c1 | edge_rot | psk8
	uint64_t launches = 0;
	int max_seq_len = 40; // longest sync sequence detect batches may need (sizes on-chip buffers)
	int max_attempts = 3; // detection rounds scheduled per batch (EXT_RACH needs 3, EDGE 2, others 1)
	DetectScratch ws;     // used by the *_batch entry points (one stream at a time)
	// detect -> demod pipelining inside trxb200_detect_demod_batch: the demodulation of chunk i runs on a side
	// stream while chunk i+1 is being detected (FP32-bound correlator / peak search beside the HBM-bound demod)
	cudaStream_t side_stream = nullptr;
	cudaEvent_t order_ev = nullptr; // orders a newly selected stream behind the work of the previous one (trxb200_set_stream)
	cudaEvent_t pipe_ev[4] = {};
	int pipe_ev_next = 0;
	struct Tune { // launch geometry; environment overrides (TRXB200_*) are read once in trxb200_init
		// overlap: measured on B200 (tools/sweep_overlap.py, profiles/r1t_overlap_sweep.txt): every overlapped geometry is
		// slower than the serial order (1.75-2.4 ms vs 1.71 ms per 2^20 bursts) - corr_nb_kernel and demod_kernel each
		// need the whole register file of an SM to hide their latencies - so the pipeline is off unless asked for
		int overlap = 0, chunk_cap = 131072;
		int detect_chunk = 1 << 30; // bursts per corr/peak launch pair.  Measured (profiles/r2l_detect_chunk_sweep.txt): one pair for the whole batch beats 262,144-burst chunks (corr 0.454 -> 0.403 ms, peak 0.220 -> 0.178 ms per 2^20 bursts); keeping the intermediates L2 resident with small chunks does not pay for the extra launches and tails
		int corr_long_wpb = 10; // warps per corr_long_kernel CTA (two CTAs per SM)
		int corr_pm1 = 1; // corr_long_kernel: sign flips instead of products for the +-1 components of the rotated GMSK sequences
		int resamp_pq = 1; // 1: resampler_pq_kernel for the 65/48 and 48/65 ratios of the multi-ARFCN interface
		int resamp_up = 0; // 1: resampler_up_kernel (three outputs per thread from a register window) for interpolating ratios; measured 4.4 ms vs 3.2 ms of resampler16_kernel on the cfg-5 stream (profiles/r2k_*), so off by default
		int fused = 0; // 1: nb_fused_kernel (one persistent warp-specialised kernel) for detect+demod in the normal-burst geometry; measured 2.75 ms vs 1.76 ms per 2^20 bursts for the three-kernel path (profiles/r2f_*), so off by default
		int host_chunk = 16384; // slots per stage of the pinned-host pipelines (H2D | kernels | D2H on three streams)
		int pull_chunk = 262144; // slots per pass of the pull chain (scratch: correlator windows + soft bits, about 2 KB per slot)
		int corr_bps = 0, peak_bps = 0, peak_warps = 16, demod_bps = 2; // 0 = derive from the on-chip footprint
		int detect_tma16 = 1; // detect_lane_kernel<int16>: TMA tiles over the slots taken four at a time (0: 4-byte cp.async chunks)
		int detect_tma = 1; // detect_lane_kernel: window chunks as TMA tiles (two per chunk) instead of one bulk copy per row
		int detect_lists = 1; // detect_lane_kernel: rounds after the first walk a list of the bursts left for them
		int detect_lane = 1; // 1: detect_lane_kernel (lane = burst, one launch) for the normal-burst geometry; 0: corr_nb_kernel + peak_kernel
		int vitac_lane = 1; // 1: vitac_lane_kernel (lane = burst), 0: vitac_kernel (warp = burst pair)
		int corr_wpb = 18; // corr_nb_kernel as one CTA of 18 warps per SM (96 registers) instead of two of 8 (118): 0.383 -> 0.373 ms per 2^20 bursts
		int demod_wpb = 8; // warps per demod CTA (two CTAs per SM); 17 = one CTA of 17 warps.  Measured per 2^20 bursts (profiles/r2o_demod_warps.txt):
				   // 10 warps 1.52 ms, 12: 1.28, 14: 1.24, 16: 1.10, 17: 1.13 - latency bound up to 12 warps, HBM bound from 16
		int ov_corr_bps = 1, ov_peak_bps = 1, ov_peak_warps = 8, ov_demod_bps = 1; // while overlapping: leave room for the other kernel
	} tune;
	std::string err;
	// once-per-context device setup (function attributes and __constant__ tables are per device)
	bool cfg_fused = false, cfg_detlane = false, cfg_rspq[2] = { false, false };
	bool cfg_detect = false, cfg_demod = false, cfg_ch64 = false, cfg_sy64 = false, sched_tables = false;
	HostStage *stage = nullptr;
	PullScratch pull;	      // trxb200_pull_batch
	PullStage *pull_stage = nullptr; // trxb200_pull_host
	// optional per-kernel timing (trxb200_profile_begin/end): CUDA events around the detect/demod kernels
	bool prof = false;
	struct ProfRec { const char *name; cudaEvent_t a, b; };
	std::vector<ProfRec> prof_recs;
};

// Host-buffer pipelines (trxb200_detect_demod_host / trxb200_pull_host): per chunk ONE host->device copy of the samples,
// ONE of the small per-slot inputs (packed struct-of-arrays in a pinned block), and on the way back ONE copy of the bulk
// result rows and ONE of the small per-slot results (packed, scattered to the caller's arrays by the host once landed).
// Caller memory that is not page-locked is staged through internal pinned buffers, so the copies stay asynchronous.
struct HostStage {
	static constexpr int kSlots = 3;
	int chunk = 0, stride = 0, soft_stride = 0;
	cudaStream_t streams[kSlots] = {};
	float *d_bursts[kSlots] = {}, *d_soft[kSlots] = {};
	uint8_t *d_in[kSlots] = {}, *d_out[kSlots] = {}; // packed small arrays (device)
	uint8_t *h_in[kSlots] = {}, *h_out[kSlots] = {}; // their pinned mirrors
	float *h_bursts[kSlots] = {}, *h_soft[kSlots] = {}; // pinned staging, allocated only for pageable callers
	int pend_lo[kSlots] = {}, pend_m[kSlots] = {};
	DetectScratch ws[kSlots];
};

namespace {

int fail(trxb200_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
	if (ctx) {
		ctx->err = what;
		if (e != cudaSuccess) {
			ctx->err += ": ";
			ctx->err += cudaGetErrorString(e);
		}
	}
	return code;
}

#define CK(call)                                                                  \
	do {                                                                      \
		cudaError_t e_ = (call);                                          \
		if (e_ != cudaSuccess) return fail(ctx, TRXB200_ECUDA, #call, e_); \
	} while (0)

// Every entry point that launches, allocates or copies runs on the context's device whatever the caller's current
// device is, and leaves the caller's device as it found it.
struct DevGuard {
	int prev = -1, dev;
	explicit DevGuard(int d) : dev(d)
	{
		if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
		if (dev >= 0 && prev != dev) cudaSetDevice(dev);
	}
	~DevGuard()
	{
		if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev);
	}
	DevGuard(const DevGuard &) = delete;
	DevGuard &operator=(const DevGuard &) = delete;
};

int post_launch(trxb200_ctx *ctx, const char *name)
{
	ctx->launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return fail(ctx, TRXB200_ECUDA, name, e);
	return TRXB200_OK;
}

// event pair around a kernel launch when profiling is on (events are recorded on the launching stream)
void prof_pre(trxb200_ctx *ctx, cudaStream_t st)
{
	if (!ctx->prof) return;
	trxb200_ctx::ProfRec r{ nullptr, nullptr, nullptr };
	cudaEventCreate(&r.a);
	cudaEventCreate(&r.b);
	cudaEventRecord(r.a, st);
	ctx->prof_recs.push_back(r);
}
void prof_post(trxb200_ctx *ctx, cudaStream_t st, const char *name)
{
	if (!ctx->prof || ctx->prof_recs.empty()) return;
	ctx->prof_recs.back().name = name;
	cudaEventRecord(ctx->prof_recs.back().b, st);
}

void fill_const_tables(const HostTables &t, ConstTables &c)
{
	std::memset(&c, 0, sizeof(c));
	int off = 0;
	auto put = [&](int id, const CorrSeq &s) {
		c.info[id].off = off;
		c.info[id].len = s.len;
		for (int i = 0; i < s.len; i++) c.seq[off + i] = make_float2(s.seq[i].r, s.seq[i].i);
		off += s.len;
		const float n = s.gain.i * s.gain.i + s.gain.r * s.gain.r; // Complex::inv() Complex.h:144-150
		c.info[id].inv_gr = s.gain.r / n;
		c.info[id].inv_gi = -s.gain.i / n;
		c.info[id].ci_den = (float)(s.len - 1) * std::sqrt(n);
		c.info[id].toa = s.toa;
		// the rotated GMSK sequences: even taps (+-1, tiny), odd taps (tiny, +-1), exactly
		bool pm1 = (s.len % 8) == 0;
		for (int i = 0; i < s.len && pm1; i++) {
			const float dom = (i & 1) ? s.seq[i].i : s.seq[i].r, oth = (i & 1) ? s.seq[i].r : s.seq[i].i;
			pm1 = std::fabs(dom) == 1.0f && std::fabs(oth) < 1e-6f;
		}
		c.info[id].pm1 = pm1 ? 1 : 0;
	};
	for (int k = 0; k < 8; k++) put(SEQ_MIDAMBLE + k, t.midamble[k]);
	for (int k = 0; k < 8; k++) put(SEQ_EDGE + k, t.edge_midamble[k]);
	for (int k = 0; k < 3; k++) put(SEQ_RACH + k, t.rach[k]);
	put(SEQ_SCH, t.sch);
	put(SEQ_DUMMY, t.dummy);
	std::memcpy(c.dnsamp, t.dnsamp, sizeof(c.dnsamp));
	std::memcpy(c.pulse_c0, t.pulse4_c0, sizeof(c.pulse_c0));
	std::memcpy(c.pulse_c1, t.pulse4_c1, sizeof(c.pulse_c1));
	std::memcpy(c.c0_inv, t.c0_inv, sizeof(c.c0_inv));
	for (int i = 0; i < 157; i++) c.rrot1[i] = make_float2(t.rrot1[i].r, t.rrot1[i].i);
	for (int i = 0; i < 625; i++) c.rot4[i] = make_float2(t.rot4[i].r, t.rot4[i].i);
	for (int i = 0; i < 157; i++) c.rot1[i] = make_float2(t.rot1[i].r, t.rot1[i].i);
	std::memcpy(c.pulse1_c0, t.pulse1_c0, sizeof(c.pulse1_c0));
	for (int i = 0; i < 8; i++) c.psk8[i] = make_float2(t.psk8[i].r, t.psk8[i].i);
	for (int i = 0; i < 156; i++) c.edge_mod_rot[i] = make_float2(t.edge_mod_rot[i].r, t.edge_mod_rot[i].i);
	for (int i = 0; i < 16; i++) c.edge_derot[i] = make_float2(t.edge_derot[i].r, t.edge_derot[i].i);
	for (int i = 0; i < 9; i++) c.edge_ideal[i] = make_float2(t.edge_ideal[i].r, t.edge_ideal[i].i);
	c.edge_rot1 = make_float2(t.edge_rot1.r, t.edge_rot1.i);
	c.edge_rot2 = make_float2(t.edge_rot2.r, t.edge_rot2.i);
	std::memcpy(c.delay, t.delay, sizeof(c.delay));
	for (int k = 0; k < 9; k++)
		for (int i = 0; i < 26; i++) c.vitac_norm[k][i] = make_float2(t.vitac_norm[k][i].r, t.vitac_norm[k][i].i);
	for (int i = 0; i < 41; i++) c.vitac_access[i] = make_float2(t.vitac_access[i].r, t.vitac_access[i].i);
	for (int i = 0; i < 64; i++) c.vitac_sch[i] = make_float2(t.vitac_sch[i].r, t.vitac_sch[i].i);
	for (int f = 0; f < kCompFilts; f++)
		for (int e = 0; e < 2; e++)
			for (int u = 0; u < 36; u++) {
				const int tt = u - e;
				c.comp0[f][e][u] = (tt >= 0 && tt < kCompTaps) ? t.comp[((size_t)f * 16 + 0) * kCompStride + tt] : 0.0f;
			}
}

int grid_for(trxb200_ctx *ctx, long work_items, int per_block, int blocks_per_sm)
{
	long want = (work_items + per_block - 1) / per_block;
	long cap = (long)ctx->sm_count * blocks_per_sm;
	if (want < 1) want = 1;
	return (int)std::min(want, cap);
}

} // namespace

extern "C" {

int trxb200_abi_version(void) { return TRXB200_ABI_VERSION; }

int trxb200_init(int device, trxb200_ctx **out)
{
	if (!out)
		return TRXB200_EINVAL;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
		cudaGetLastError();
		return TRXB200_ENODEV;
	}
	trxb200_ctx *ctx = new (std::nothrow) trxb200_ctx();
	if (!ctx)
		return TRXB200_ENOMEM;
	ctx->device = device;
	cudaDeviceProp prop;
	if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
		delete ctx;
		return TRXB200_ENODEV;
	}
	if (prop.major < 10) { // sm_100a code only; no fallback path exists
		delete ctx;
		return TRXB200_ENODEV;
	}
	ctx->sm_count = prop.multiProcessorCount;
	ctx->ht = new HostTables();
	build_host_tables(*ctx->ht);
	ConstTables *c = new ConstTables();
	fill_const_tables(*ctx->ht, *c);
	cudaError_t e = cudaMemcpyToSymbol(c_tab, c, sizeof(ConstTables));
	// per tap of the +-1 sequences (corr_long_items_pm1): x * (+-1) is a sign flip (.x, .y = XOR masks for re, im), the other
	// component keeps its multiplication (.z, .w = the operand pair of that product, as floats)
	std::vector<uint4> pm(SEQ_STORE, make_uint4(0, 0, 0, 0));
	for (int id = 0; id < SEQ_COUNT; id++) {
		if (!c->info[id].pm1) continue;
		for (int i = 0; i < c->info[id].len; i++) {
			const float2 h = c->seq[c->info[id].off + i];
			auto sgn = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); return u & 0x80000000u; };
			auto bits = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; };
			uint4 t;
			if ((i & 1) == 0) { // (x.re * hr, x.im * hr) = +-x;  (x.re * hi, x.im * -hi) stays a product
				t.x = sgn(h.x); t.y = sgn(h.x); t.z = bits(h.y); t.w = bits(-h.y);
			} else {	    // (x.re * hr, x.im * hr) stays a product;  (x.re * hi, x.im * -hi) = (+-x.re, -+x.im)
				t.x = sgn(h.y); t.y = sgn(-h.y); t.z = bits(h.x); t.w = bits(h.x);
			}
			pm[c->info[id].off + i] = t;
		}
	}
	delete c;
	if (e == cudaSuccess) e = cudaMalloc(&ctx->d_seq_pm1, pm.size() * sizeof(uint4));
	if (e == cudaSuccess) e = cudaMemcpy(ctx->d_seq_pm1, pm.data(), pm.size() * sizeof(uint4), cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
	// interpolation weights for peak_kernel: tap-major, columns bit-reversed (detect.cu kSinc512)
	std::vector<float> wtab((size_t)21 * 512);
	for (int d = 0; d < 21; d++)
		for (int F = 0; F < 512; F++) {
			int r = 0;
			for (int bit = 0; bit < 9; bit++) r |= ((F >> bit) & 1) << (8 - bit);
			wtab[(size_t)d * 512 + r] = ctx->ht->sinc512[(size_t)std::abs(512 * (d - 10) - F)];
		}
	if (e == cudaSuccess) e = cudaMalloc(&ctx->d_sinc512, wtab.size() * sizeof(float));
	if (e == cudaSuccess) e = cudaMalloc(&ctx->d_comp, (ctx->ht->comp.size() + 16) * sizeof(float)); // + decimator taps
	if (e == cudaSuccess)
		e = cudaMemcpy(ctx->d_sinc512, wtab.data(), wtab.size() * sizeof(float), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(ctx->d_comp, ctx->ht->comp.data(), ctx->ht->comp.size() * sizeof(float), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(ctx->d_comp + ctx->ht->comp.size(), ctx->ht->dnsamp, 16 * sizeof(float), cudaMemcpyHostToDevice);
	if (e == cudaSuccess) {
		std::vector<float2> et(25);
		for (int i = 0; i < 16; i++) et[i] = make_float2(ctx->ht->edge_derot[i].r, ctx->ht->edge_derot[i].i);
		for (int i = 0; i < 9; i++) et[16 + i] = make_float2(ctx->ht->edge_ideal[i].r, ctx->ht->edge_ideal[i].i);
		e = cudaMalloc(&ctx->d_edge_tab, et.size() * sizeof(float2));
		if (e == cudaSuccess) e = cudaMemcpy(ctx->d_edge_tab, et.data(), et.size() * sizeof(float2), cudaMemcpyHostToDevice);
	}
	if (e == cudaSuccess) {
		std::vector<float> mt(kModTabFloats, 0.0f);
		const HostTables &t = *ctx->ht;
		for (int i = 0; i < 625; i++) { mt[kModRot4 + 2 * i] = t.rot4[i].r; mt[kModRot4 + 2 * i + 1] = t.rot4[i].i; }
		for (int i = 0; i < 16; i++) mt[kModC0 + i] = t.pulse4_c0[i];
		for (int i = 0; i < 8; i++) mt[kModC1 + i] = t.pulse4_c1[i];
		for (int i = 0; i < 156; i++) { mt[kModEdgeRot + 2 * i] = t.edge_mod_rot[i].r; mt[kModEdgeRot + 2 * i + 1] = t.edge_mod_rot[i].i; }
		for (int i = 0; i < 8; i++) { mt[kModPsk8 + 2 * i] = t.psk8[i].r; mt[kModPsk8 + 2 * i + 1] = t.psk8[i].i; }
		e = cudaMalloc(&ctx->d_mod_tab, mt.size() * sizeof(float));
		if (e == cudaSuccess) e = cudaMemcpy(ctx->d_mod_tab, mt.data(), mt.size() * sizeof(float), cudaMemcpyHostToDevice);
	}
	if (e != cudaSuccess) {
		fprintf(stderr, "trxb200_init: %s\n", cudaGetErrorString(e));
		trxb200_destroy(ctx);
		return TRXB200_ECUDA;
	}
	if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess) {
		trxb200_destroy(ctx);
		return TRXB200_ECUDA;
	}
	for (auto &ev : ctx->pipe_ev)
		if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
			trxb200_destroy(ctx);
			return TRXB200_ECUDA;
		}
	if (cudaEventCreateWithFlags(&ctx->order_ev, cudaEventDisableTiming) != cudaSuccess) {
		trxb200_destroy(ctx);
		return TRXB200_ECUDA;
	}
	{
		auto env_int = [](const char *name, int &v) {
			const char *e = getenv(name);
			if (e && *e) v = atoi(e);
		};
		trxb200_ctx::Tune &t = ctx->tune;
		env_int("TRXB200_OVERLAP", t.overlap);
		env_int("TRXB200_FUSED", t.fused);
		env_int("TRXB200_RESAMP_UP", t.resamp_up);
		env_int("TRXB200_RESAMP_PQ", t.resamp_pq);
		env_int("TRXB200_CORR_PM1", t.corr_pm1);
		env_int("TRXB200_CORR_LONG_WPB", t.corr_long_wpb);
		env_int("TRXB200_CHUNK", t.chunk_cap);
		env_int("TRXB200_DETECT_CHUNK", t.detect_chunk);
		env_int("TRXB200_DEMOD_WPB", t.demod_wpb);
		env_int("TRXB200_CORR_WPB", t.corr_wpb);
		env_int("TRXB200_VITAC_LANE", t.vitac_lane);
		env_int("TRXB200_DETECT_LANE", t.detect_lane);
		env_int("TRXB200_DETECT_LISTS", t.detect_lists);
		env_int("TRXB200_DETECT_TMA16", t.detect_tma16);
		env_int("TRXB200_DETECT_TMA", t.detect_tma);
		env_int("TRXB200_PULL_CHUNK", t.pull_chunk);
		env_int("TRXB200_HOST_CHUNK", t.host_chunk);
		env_int("TRXB200_CORR_BPS", t.corr_bps);
		env_int("TRXB200_PEAK_BPS", t.peak_bps);
		env_int("TRXB200_PEAK_WARPS", t.peak_warps);
		env_int("TRXB200_DEMOD_BPS", t.demod_bps);
		env_int("TRXB200_OV_CORR_BPS", t.ov_corr_bps);
		env_int("TRXB200_OV_PEAK_BPS", t.ov_peak_bps);
		env_int("TRXB200_OV_PEAK_WARPS", t.ov_peak_warps);
		env_int("TRXB200_OV_DEMOD_BPS", t.ov_demod_bps);
		if (t.chunk_cap < 4096) t.chunk_cap = 4096;
		if (t.pull_chunk < 1024) t.pull_chunk = 1024;
		if (t.host_chunk < 1024) t.host_chunk = 1024;
	}
	ctx->stream = ctx->own_stream;
	*out = ctx;
	return TRXB200_OK;
}

static void stage_free(HostStage *s);
static void pull_scratch_free(PullScratch &w);
static void pull_stage_free(PullStage *s);

void trxb200_destroy(trxb200_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	if (ctx->stage) stage_free(ctx->stage);
	if (ctx->pull_stage) pull_stage_free(ctx->pull_stage);
	pull_scratch_free(ctx->pull);
	if (ctx->d_sinc512) cudaFree(ctx->d_sinc512);
	cudaFree(ctx->ws.corr);
	cudaFree(ctx->ws.pwr);
	cudaFree(ctx->ws.list);
	if (ctx->d_comp) cudaFree(ctx->d_comp);
	if (ctx->d_seq_pm1) cudaFree(ctx->d_seq_pm1);
	if (ctx->d_edge_tab) cudaFree(ctx->d_edge_tab);
	if (ctx->d_tmap) cudaFree(ctx->d_tmap);
	if (ctx->d_mod_tab) cudaFree(ctx->d_mod_tab);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
	for (auto ev : ctx->pipe_ev)
		if (ev) cudaEventDestroy(ev);
	if (ctx->order_ev) cudaEventDestroy(ctx->order_ev);
	delete ctx->ht;
	delete ctx;
}

const char *trxb200_last_error(trxb200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// The context owns ONE set of device scratch (correlation intermediates, pull scratch, filterbank history): work
// enqueued through it must not overlap itself.  Switching streams therefore orders the new stream behind everything the
// context enqueued on the previous one (event record + wait, no host synchronisation).
static int switch_stream(trxb200_ctx *ctx, cudaStream_t ns)
{
	if (ns == ctx->stream) return TRXB200_OK;
	DevGuard dg(ctx->device);
	CK(cudaEventRecord(ctx->order_ev, ctx->stream));
	CK(cudaStreamWaitEvent(ns, ctx->order_ev, 0));
	ctx->stream = ns;
	return TRXB200_OK;
}
int trxb200_set_stream(trxb200_ctx *ctx, void *s)
{
	if (!ctx) return TRXB200_EINVAL;
	return switch_stream(ctx, (cudaStream_t)s); // NULL is the CUDA default stream, exactly as passed
}
int trxb200_use_own_stream(trxb200_ctx *ctx)
{
	if (!ctx) return TRXB200_EINVAL;
	return switch_stream(ctx, ctx->own_stream);
}
void *trxb200_get_stream(trxb200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int trxb200_sync(trxb200_ctx *ctx)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	CK(cudaStreamSynchronize(ctx->stream));
	return TRXB200_OK;
}
int trxb200_device(trxb200_ctx *ctx) { return ctx ? ctx->device : -1; }
int trxb200_sm_count(trxb200_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t trxb200_launch_count(trxb200_ctx *ctx) { return ctx ? ctx->launches : 0; }

/* ---------------- device memory helpers for bindings that have no CUDA runtime of their own ---------------- */
int trxb200_dev_alloc(trxb200_ctx *ctx, size_t bytes, void **out)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !out) return TRXB200_EINVAL;
	*out = nullptr;
	CK(cudaSetDevice(ctx->device));
	cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
	if (e != cudaSuccess) return fail(ctx, e == cudaErrorMemoryAllocation ? TRXB200_ENOMEM : TRXB200_ECUDA, "dev_alloc", e);
	return TRXB200_OK;
}
int trxb200_dev_free(trxb200_ctx *ctx, void *p)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	if (p) CK(cudaFree(p));
	return TRXB200_OK;
}
int trxb200_copy_to_device(trxb200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || (bytes && (!dst || !src))) return TRXB200_EINVAL;
	if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return TRXB200_OK;
}
int trxb200_copy_to_host(trxb200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || (bytes && (!dst || !src))) return TRXB200_EINVAL;
	if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return TRXB200_OK;
}
int trxb200_memset_device(trxb200_ctx *ctx, void *dst, int value, size_t bytes)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || (bytes && !dst)) return TRXB200_EINVAL;
	if (bytes) CK(cudaMemsetAsync(dst, value, bytes, ctx->stream));
	return TRXB200_OK;
}

int trxb200_profile_begin(trxb200_ctx *ctx)
{
	if (!ctx) return TRXB200_EINVAL;
	for (auto &r : ctx->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
	ctx->prof_recs.clear();
	ctx->prof = true;
	return TRXB200_OK;
}

int trxb200_profile_end(trxb200_ctx *ctx, char *out, int cap)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !out || cap < 1) return TRXB200_EINVAL;
	ctx->prof = false;
	CK(cudaDeviceSynchronize());
	struct Acc { const char *name; double ms; int count; };
	std::vector<Acc> acc;
	for (auto &r : ctx->prof_recs) {
		float ms = 0.0f;
		if (r.name && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
			bool found = false;
			for (auto &a : acc)
				if (!strcmp(a.name, r.name)) { a.ms += ms; a.count++; found = true; }
			if (!found) acc.push_back(Acc{ r.name, ms, 1 });
		}
		cudaEventDestroy(r.a);
		cudaEventDestroy(r.b);
	}
	ctx->prof_recs.clear();
	std::string s;
	for (auto &a : acc) {
		char buf[160];
		snprintf(buf, sizeof(buf), "%s%s:%.6f:%d", s.empty() ? "" : ";", a.name, a.ms, a.count);
		s += buf;
	}
	if ((int)s.size() + 1 > cap) return TRXB200_EINVAL;
	std::memcpy(out, s.c_str(), s.size() + 1);
	return (int)acc.size();
}

int trxb200_detect_config(trxb200_ctx *ctx, int max_seq_len, int max_attempts)
{
	if (!ctx) return TRXB200_EINVAL;
	if (max_seq_len != 16 && max_seq_len != 40) return fail(ctx, TRXB200_EINVAL, "detect_config: max_seq_len must be 16 or 40");
	if (max_attempts < 1 || max_attempts > 3) return fail(ctx, TRXB200_EINVAL, "detect_config: max_attempts must be 1..3");
	ctx->max_seq_len = max_seq_len;
	ctx->max_attempts = max_attempts;
	return TRXB200_OK;
}

int trxb200_get_table(trxb200_ctx *ctx, const char *name, int idx, float *out, int max_floats)
{
	if (!ctx || !name || !out) return TRXB200_EINVAL;
	const HostTables &t = *ctx->ht;
	std::vector<float> v;
	std::string s(name);
	auto real_as_cx = [&](const float *p, int n) { for (int i = 0; i < n; i++) { v.push_back(p[i]); v.push_back(0.0f); } };
	auto cx = [&](const cf *p, int n) { for (int i = 0; i < n; i++) { v.push_back(p[i].r); v.push_back(p[i].i); } };
	auto seq = [&](const CorrSeq &c, bool meta) {
		if (meta) { v = { c.gain.r, c.gain.i, c.toa }; }
		else cx(c.seq, c.len);
	};
	const bool meta = s.size() > 5 && s.compare(s.size() - 5, 5, "_meta") == 0;
	const std::string base = meta ? s.substr(0, s.size() - 5) : s;
	if (base == "sinc") v.assign(t.sinc, t.sinc + kSincSize + 1);
	else if (base == "rot4") cx(t.rot4, 625);
	else if (base == "rrot4") cx(t.rrot4, 625);
	else if (base == "rot1") cx(t.rot1, 157);
	else if (base == "rrot1") cx(t.rrot1, 157);
	else if (base == "delay") { if (idx < 0 || idx >= kDelayFilts) return TRXB200_EINVAL; real_as_cx(t.delay[idx], kDelayTaps); }
	else if (base == "pulse4_c0") real_as_cx(t.pulse4_c0, 16);
	else if (base == "pulse4_c1") real_as_cx(t.pulse4_c1, 8);
	else if (base == "pulse4_c0inv") real_as_cx(t.c0_inv, 5);
	else if (base == "pulse1_c0") real_as_cx(t.pulse1_c0, 4);
	else if (base == "dnsamp") real_as_cx(t.dnsamp, 16);
	else if (base == "psk8") cx(t.psk8, 8);
	else if (base == "midamble") { if (idx < 0 || idx > 7) return TRXB200_EINVAL; seq(t.midamble[idx], meta); }
	else if (base == "edge_midamble") { if (idx < 0 || idx > 7) return TRXB200_EINVAL; seq(t.edge_midamble[idx], meta); }
	else if (base == "rach") { if (idx < 0 || idx > 2) return TRXB200_EINVAL; seq(t.rach[idx], meta); }
	else if (base == "sch") seq(t.sch, meta);
	else if (base == "dummy") seq(t.dummy, meta);
	else if (base == "vitac_norm") { if (idx < 0 || idx > 8) return TRXB200_EINVAL; cx(t.vitac_norm[idx], 26); }
	else if (base == "vitac_access") cx(t.vitac_access, 41);
	else if (base == "vitac_sch") cx(t.vitac_sch, 64);
	else if (base == "interp_w") v = t.interp_w;
	else if (base == "sinc512") v = t.sinc512;
	else if (base == "comp") { if (idx < 0 || idx >= kCompFilts * 16) return TRXB200_EINVAL; v.assign(&t.comp[(size_t)idx * kCompStride], &t.comp[(size_t)idx * kCompStride] + kCompStride); }
	else return TRXB200_EINVAL;
	if ((int)v.size() > max_floats) return TRXB200_EINVAL;
	std::memcpy(out, v.data(), v.size() * sizeof(float));
	return (int)v.size();
}

/* ---------------- modulators ---------------- */
int trxb200_modulate_gmsk_batch(trxb200_ctx *ctx, const uint8_t *bits, int nbits, int bits_stride, int n, float *out,
				int out_stride)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !bits || !out || nbits < 2 || nbits > 156 || bits_stride < nbits || out_stride < 625 || n < 0)
		return fail(ctx, TRXB200_EINVAL, "modulate_gmsk: bad argument");
	if (n == 0) return TRXB200_OK;
	prof_pre(ctx, ctx->stream);
	modulate_gmsk_kernel<<<grid_for(ctx, n, kModWarps, 8), kModWarps * 32, 0, ctx->stream>>>(bits, nbits, bits_stride, n, out, out_stride, ctx->d_mod_tab, -0.0f);
	prof_post(ctx, ctx->stream, "modulate_gmsk_kernel");
	return post_launch(ctx, "modulate_gmsk_kernel");
}

int trxb200_modulate_edge_batch(trxb200_ctx *ctx, const uint8_t *bits, int nbits, int bits_stride, int n, float *out,
				int out_stride)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !bits || !out || nbits < 3 || (nbits % 3) || bits_stride < nbits || out_stride < 625 || n < 0)
		return fail(ctx, TRXB200_EINVAL, "modulate_edge: bad argument");
	if (n == 0) return TRXB200_OK;
	modulate_edge_kernel<<<grid_for(ctx, n, kModWarps, 8), kModWarps * 32, 0, ctx->stream>>>(bits, nbits, bits_stride, n, out, out_stride, ctx->d_mod_tab, -0.0f);
	return post_launch(ctx, "modulate_edge_kernel");
}

// modulateBurst(bits, guard, sps, emptyPulse) outside the 4-sps Laurent case and modulateEdgeBurst(bits, sps, true)
// (sigProcLib.cpp:558-580,672-689,938-979): mode 0 modulateBurstBasic (sps 1), 1 rotateBurst (sps 1 or 4), 2 rotateEdgeBurst
int trxb200_modulate_basic_batch(trxb200_ctx *ctx, const uint8_t *bits, int nbits, int bits_stride, int n, int guard, int sps,
				 int mode, float *out, int out_stride)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !bits || !out || n < 0 || nbits < 1 || bits_stride < nbits || guard < 0 || (sps != 1 && sps != 4) || mode < 0 || mode > 2)
		return fail(ctx, TRXB200_EINVAL, "modulate_basic: bad argument");
	if (mode == 0 && sps != 1) return fail(ctx, TRXB200_EINVAL, "modulate_basic: the single-pulse shaper exists at 1 sps only");
	if (mode == 2 && (nbits % 3 || guard)) return fail(ctx, TRXB200_EINVAL, "modulate_basic: 8-PSK bits come in threes, no guard period");
	const int nsym = mode == 2 ? nbits / 3 : nbits;
	const int olen = sps * (nsym + guard);
	// the rotators cover 157 symbols (1 sps) / 625 samples (4 sps); the 8-PSK rotator 156 symbols
	if (olen > out_stride || (mode != 2 && olen > (sps == 1 ? 157 : 625)) || (mode == 2 && nsym > 156))
		return fail(ctx, TRXB200_EINVAL, "modulate_basic: burst longer than the rotation tables / output row");
	if (n == 0) return TRXB200_OK;
	modulate_basic_kernel<<<grid_for(ctx, (long)n * 32, 256, 8), 256, 0, ctx->stream>>>(bits, nbits, bits_stride, n, guard, sps, mode, out, out_stride);
	return post_launch(ctx, "modulate_basic_kernel");
}

/* ---------------- detection / demodulation ---------------- */
static int gcd_i(long a, long b) { while (b) { long t = a % b; a = b; b = t; } return (int)a; }

// after_chunk (optional) is called once the detection of bursts [lo, lo + m) has been enqueued on st: the fused
// entry point uses it to start their demodulation on the side stream.  overlapped = the launch geometry leaves
// room on every SM for the demod kernel running beside detection.
struct ChunkHook {
	virtual int operator()(long lo, int m) = 0;
};
static bool nb_geometry(const trxb200_ctx *ctx, int bound) { return ctx->max_seq_len == 16 && bound <= 4; }

static int launch_detect(trxb200_ctx *ctx, cudaStream_t st, DetectScratch &ws, const float *bursts, int stride, int n,
			 const uint8_t *type, const uint8_t *tsc, const uint16_t *max_toa, int bound, float thresh, int32_t *rc,
			 float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags, int scan_clip,
			 ChunkHook *after_chunk = nullptr, bool overlapped = false, const int16_t *iq = nullptr, int iq_stride = 0,
			 bool sch = false, int sps1_len = 0)
{
	const trxb200_ctx::Tune &tn = ctx->tune;
	// sch: detectSCHBurst's full search (156 correlation outputs, 64-symbol sequence), one attempt, no per-burst arrays
	// 16-symbol sync sequences at max_toa <= 4: the register-blocked corr_nb_kernel with its fixed row lengths
	// sps1_len: bursts of that many samples at one sample per symbol - no decimator, hence the long-window correlator
	const bool nb = !sch && !sps1_len && nb_geometry(ctx, bound);
	const int lmax = sch ? 156 : nb ? 20 : (16 + bound + 1) & ~1;	    // row pitch of the correlation vectors (even: 16-byte row loads in peak_kernel)
	const int ndmax = sch ? 64 + 156 - 1 : nb ? 35 : ctx->max_seq_len + 16 + bound - 1; // decimated samples a correlation window needs
	// ---- launch geometry ----
	if (iq && !nb) return fail(ctx, TRXB200_EINVAL, "detect: int16 rows are read by corr_nb_kernel only");
	int cw = 8; // warps per corr block
	const bool cwide = nb && !iq && !overlapped && tn.corr_wpb == 18; // corr_nb_kernel as one CTA of 18 warps per SM
	if (cwide) cw = 18;
	if (!nb) {
		cw = std::max(1, std::min(tn.corr_long_wpb, 10));
		while (cw > 1 && corr_lg_warp_bytes(ndmax) * cw + corr_lg_hdr_bytes() > 110 * 1024) cw--;
	}
	const size_t csmem = nb ? corr_nb_hdr_bytes() + corr_nb_warp_bytes() * cw : corr_lg_warp_bytes(ndmax) * cw + corr_lg_hdr_bytes();
	const int cgroup = nb ? kNbGroup : 1;
	int pw = std::max(1, std::min(32, overlapped ? tn.ov_peak_warps : tn.peak_warps)); // warps per peak block
	// as many warps as the shared memory of one SM holds (long correlation vectors: 7 warps at lmax 80, not a power of two)
	pw = std::max(1, std::min(pw, (int)((225 * 1024 - peak_hdr_bytes()) / peak_warp_bytes(lmax))));
	const size_t psmem = peak_hdr_bytes() + peak_warp_bytes(lmax) * pw;
	if (csmem > 227 * 1024 || psmem > 227 * 1024)
		return fail(ctx, TRXB200_EINVAL, "detect: max_toa_bound too large for on-chip buffers");
	if (!ctx->cfg_detect) {
		CK(cudaFuncSetAttribute(corr_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
		CK(cudaFuncSetAttribute(corr_nb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
		CK(cudaFuncSetAttribute(corr_nb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
		CK(cudaFuncSetAttribute(corr_nb_kernel<false, 18, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
		CK(cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
		ctx->cfg_detect = true;
	}
	int cbps = (int)std::max<size_t>(1, std::min<size_t>(2, (225 * 1024) / (csmem + 1024)));
	int pbps = (int)std::max<size_t>(1, std::min<size_t>(2, (225 * 1024) / (psmem + 1024)));
	const int want_c = overlapped ? tn.ov_corr_bps : tn.corr_bps, want_p = overlapped ? tn.ov_peak_bps : tn.peak_bps;
	if (want_c > 0) cbps = std::min(cbps, want_c);
	if (want_p > 0) pbps = std::min(pbps, want_p);
	const long corr_sweep = (long)ctx->sm_count * cbps * cw * cgroup; // bursts one full wave of corr warps covers
	const long peak_sweep = (long)ctx->sm_count * pbps * pw * 32;
	// chunk: a whole number of sweeps of both kernels (no tail quantisation), small enough that the
	// intermediates (lmax*8 + ndmax*4 bytes per burst) stay L2 resident
	long chunk = corr_sweep / gcd_i(corr_sweep, peak_sweep) * peak_sweep;
	const long cap = after_chunk ? tn.chunk_cap : tn.detect_chunk;
	if (chunk > cap) chunk = std::max<long>(1, cap / peak_sweep) * peak_sweep;
	else chunk *= std::max<long>(1, cap / chunk);
	if (chunk > n) chunk = n;
	if (nb && tn.detect_lane && !overlapped && !after_chunk) {
		// lane = burst: decimate + correlate + peak logic per thread, one launch per round, no intermediates
		if (!ctx->cfg_detlane) {
			CK(cudaFuncSetAttribute(detect_lane_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_lane_smem()));
			CK(cudaFuncSetAttribute(detect_lane_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_lane_smem()));
			CK(cudaFuncSetAttribute(detect_lane_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_lane_smem()));
			CK(cudaFuncSetAttribute(detect_lane_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_lane_smem()));
			ctx->cfg_detlane = true;
		}
		const int nrounds = ctx->max_attempts;
		// later rounds walk the list of bursts the round before left for them (detect_lane.cu); lists ping-pong
		const bool lists = nrounds > 1 && tn.detect_lists;
		if (lists) {
			if (ws.list_cap < (size_t)n) {
				CK(cudaStreamSynchronize(st));
				cudaFree(ws.list);
				ws.list = nullptr; ws.list_cap = 0;
				const size_t cap = ((size_t)n + 1023) & ~(size_t)1023;
				CK(cudaMalloc(&ws.list, (16 + 2 * cap) * sizeof(int)));
				ws.list_cap = cap;
			}
			CK(cudaMemsetAsync(ws.list, 0, 16 * sizeof(int), st));
		}
		for (int r = 0; r < nrounds; r++) {
			DetLaneParams dp;
			dp.list_in = dp.list_n_in = nullptr;
			dp.list_out = dp.list_n_out = nullptr;
			if (lists) {
				// round r appends to list r & 1 and counts in ws.list[4 * r] (a counter of its own per round); lists start at ws.list + 16
				if (r > 0) { dp.list_in = ws.list + 16 + (size_t)((r - 1) & 1) * ws.list_cap; dp.list_n_in = ws.list + 4 * (r - 1); }
				if (r + 1 < nrounds) { dp.list_out = ws.list + 16 + (size_t)(r & 1) * ws.list_cap; dp.list_n_out = ws.list + 4 * r; }
			}
			CorrParams &c = dp.c;
			c.bursts = bursts; c.stride = stride; c.n = n; c.iq = iq; c.iq_stride = iq_stride; c.type = type; c.tsc = tsc; c.max_toa = max_toa;
			c.rc = rc; c.round = r; c.sch = 0; c.max_toa_bound = bound; c.lmax = lmax; c.ndmax = ndmax; c.corr = nullptr; c.pwr = nullptr;
			c.negzero = -0.0f;
			PeakParams &q = dp.q;
			q.n = n; q.type = type; q.tsc = tsc; q.max_toa = max_toa; q.round = r; q.last_round = (r == nrounds - 1); q.sch = 0;
			q.max_toa_bound = bound; q.thresh = thresh; q.lmax = lmax; q.ndmax = ndmax; q.corr = nullptr; q.pwr = nullptr;
			q.sinc512 = ctx->d_sinc512; q.negzero = -0.0f; q.rc = rc; q.amp = amp; q.toa = toa; q.ci = ci; q.tsc_out = tsc_out; q.flags = flags;
			dp.tma_on = 0;
			dp.tmap = nullptr;
			dp.tma_shift = 0;
			// float rows taken two at a time ([n / 2][4 * stride] floats, row pitch 16 * stride bytes) and int16 rows taken four at a
			// time ([n / 4][4 * iq_stride] I/Q pairs, row pitch 16 * iq_stride bytes) are legal TMA tensors.  An array that starts off
			// the 16-byte grid is described from the grid point in front of it (no box reaches the samples before the array).
			const bool tma_f = !iq && n >= 2 && (reinterpret_cast<uintptr_t>(bursts) & 7u) == 0;
			const bool tma_i = iq && n >= 4 && (reinterpret_cast<uintptr_t>(iq) & 3u) == 0 && tn.detect_tma16;
			if ((tma_f || tma_i) && !dp.list_in && tn.detect_tma) {
				typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
							     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
							     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
				static EncodeFn encode = nullptr;
				if (!encode) {
					void *fn = nullptr;
					cudaDriverEntryPointQueryResult qr;
					if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
						encode = (EncodeFn)fn;
					else
						cudaGetLastError();
				}
				if (encode && !ctx->d_tmap) {
					if (cudaMalloc(&ctx->d_tmap, 4 * sizeof(CUtensorMap)) != cudaSuccess) { ctx->d_tmap = nullptr; cudaGetLastError(); }
				}
				if (encode && ctx->d_tmap) {
					alignas(64) CUtensorMap tm;
					CUresult er;
					int shift;
					if (tma_f) {
						shift = (int)((reinterpret_cast<uintptr_t>(bursts) >> 3) & 1u);
						const cuuint64_t gdim[2] = { (cuuint64_t)4 * (cuuint64_t)stride, (cuuint64_t)(n / 2) };
						const cuuint64_t gstr[1] = { (cuuint64_t)16 * (cuuint64_t)stride };
						const cuuint32_t box[2] = { 2u * kDlBulkPitch, 16u }, est[2] = { 1u, 1u }; // kDlBulkPitch samples x 16 row pairs
						er = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(bursts - 2 * shift), gdim, gstr, box, est,
							    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
							    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
					} else {
						shift = (int)((reinterpret_cast<uintptr_t>(iq) >> 2) & 3u);
						const cuuint64_t gdim[2] = { (cuuint64_t)4 * (cuuint64_t)iq_stride, (cuuint64_t)(n / 4) };
						const cuuint64_t gstr[1] = { (cuuint64_t)16 * (cuuint64_t)iq_stride };
						const cuuint32_t box[2] = { (cuuint32_t)kDlBox16, 8u }, est[2] = { 1u, 1u }; // kDlBox16 I/Q pairs x 8 row quadruples
						er = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<int16_t *>(iq - 2 * shift), gdim, gstr, box, est,
							    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
							    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
					}
					if (er == CUDA_SUCCESS) {
						// the descriptor lives in global memory (one slot per round; stream order keeps a slot intact while a
						// kernel that reads it is still running)
						unsigned char *slot = reinterpret_cast<unsigned char *>(ctx->d_tmap) + (size_t)(r & 3) * sizeof(CUtensorMap);
						if (cudaMemcpyAsync(slot, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice, st) == cudaSuccess) {
							dp.tmap = slot;
							dp.tma_on = 1;
							dp.tma_shift = shift;
						} else {
							cudaGetLastError();
						}
					}
				}
			}
			const int ntiles = (n + 31) / 32;
			const int grid = std::max(1, std::min((ntiles + kDlWarps - 1) / kDlWarps, ctx->sm_count));
			prof_pre(ctx, st);
			if (dp.list_in) {
				if (iq) detect_lane_kernel<true, true><<<grid, kDlWarps * 32, det_lane_smem(), st>>>(dp);
				else detect_lane_kernel<false, true><<<grid, kDlWarps * 32, det_lane_smem(), st>>>(dp);
			} else {
				if (iq) detect_lane_kernel<true><<<grid, kDlWarps * 32, det_lane_smem(), st>>>(dp);
				else detect_lane_kernel<false><<<grid, kDlWarps * 32, det_lane_smem(), st>>>(dp);
			}
			prof_post(ctx, st, "detect_lane_kernel");
			const int e = post_launch(ctx, "detect_lane_kernel");
			if (e) return e;
		}
		if (scan_clip) {
			prof_pre(ctx, st);
			clip_kernel<<<std::max(1, std::min((n + 7) / 8, ctx->sm_count * 8)), 256, 0, st>>>(bursts, stride, n, rc, flags, type);
			prof_post(ctx, st, "clip_kernel");
			return post_launch(ctx, "clip_kernel");
		}
		return TRXB200_OK;
	}
	// ---- scratch ----
	// a short remainder (< chunk / 4) rides along with the last full chunk instead of paying two more launches
	const long rem = n % chunk;
	const long max_m = (rem > 0 && rem < chunk / 4 && n > chunk) ? chunk + rem : chunk;
	const size_t need_c = (size_t)max_m * lmax * sizeof(float2), need_p = (size_t)max_m * ndmax * sizeof(float);
	if (ws.corr_bytes < need_c) {
		CK(cudaStreamSynchronize(st));
		cudaFree(ws.corr);
		ws.corr = nullptr; ws.corr_bytes = 0;
		CK(cudaMalloc(&ws.corr, need_c));
		ws.corr_bytes = need_c;
	}
	if (ws.pwr_bytes < need_p) {
		CK(cudaStreamSynchronize(st));
		cudaFree(ws.pwr);
		ws.pwr = nullptr; ws.pwr_bytes = 0;
		CK(cudaMalloc(&ws.pwr, need_p));
		ws.pwr_bytes = need_p;
	}
	const int rounds = sch ? 1 : ctx->max_attempts;
	long step_m = 0;
	for (long lo = 0; lo < n; lo += step_m) {
		long mm = std::min<long>(chunk, n - lo);
		if (n - lo - mm > 0 && n - lo - mm < chunk / 4) mm = n - lo;
		step_m = mm;
		const int m = (int)mm;
		for (int r = 0; r < rounds; r++) {
			CorrParams c;
			c.bursts = bursts ? bursts + (size_t)lo * stride * 2 : nullptr; c.stride = stride; c.n = m;
			c.iq = iq ? iq + (size_t)lo * iq_stride * 2 : nullptr; c.iq_stride = iq_stride;
			c.type = sch ? nullptr : type + lo; c.tsc = sch ? nullptr : tsc + lo; c.max_toa = sch ? nullptr : max_toa + lo; c.rc = rc + lo; c.round = r;
			c.sch = sch ? 1 : 0;
			c.sps1_len = sps1_len;
			c.seq_pm1 = ctx->tune.corr_pm1 ? ctx->d_seq_pm1 : nullptr;
			c.max_toa_bound = bound; c.lmax = lmax; c.ndmax = ndmax; c.corr = ws.corr; c.pwr = ws.pwr; c.negzero = -0.0f;
			const int ngroups = (m + cgroup - 1) / cgroup;
			const int cgrid = std::max(1, std::min((ngroups + cw - 1) / cw, ctx->sm_count * cbps));
			prof_pre(ctx, st);
			if (nb && iq) corr_nb_kernel<true><<<cgrid, cw * 32, csmem, st>>>(c);
			else if (cwide) corr_nb_kernel<false, 18, 1><<<cgrid, cw * 32, csmem, st>>>(c);
			else if (nb) corr_nb_kernel<false><<<cgrid, cw * 32, csmem, st>>>(c);
			else corr_long_kernel<<<cgrid, cw * 32, csmem, st>>>(c);
			prof_post(ctx, st, "corr_kernel");
			int e = post_launch(ctx, "corr_kernel");
			if (e) return e;
			PeakParams q;
			q.n = m; q.type = sch ? nullptr : type + lo; q.tsc = sch ? nullptr : tsc + lo; q.max_toa = sch ? nullptr : max_toa + lo; q.round = r;
			q.last_round = (r == rounds - 1); q.sch = sch ? 1 : 0;
			q.max_toa_bound = bound; q.thresh = thresh; q.lmax = lmax; q.ndmax = ndmax; q.corr = ws.corr; q.pwr = ws.pwr;
			q.sinc512 = ctx->d_sinc512; q.negzero = -0.0f; q.rc = rc + lo; q.amp = amp + (size_t)lo * 2; q.toa = toa + lo; q.ci = ci + lo;
			q.tsc_out = tsc_out ? tsc_out + lo : nullptr; q.flags = flags ? flags + lo : nullptr;
			if (sps1_len) q.dec_size = sps1_len; // computeCI's bound is the correlator input's own length (:1617)
			const int ntiles = (m + 31) / 32;
			const int pgrid = std::max(1, std::min((ntiles + pw - 1) / pw, ctx->sm_count * pbps));
			prof_pre(ctx, st);
			peak_kernel<<<pgrid, pw * 32, psmem, st>>>(q);
			prof_post(ctx, st, "peak_kernel");
			e = post_launch(ctx, "peak_kernel");
			if (e) return e;
		}
		if (after_chunk) {
			const int e = (*after_chunk)(lo, m);
			if (e) return e;
		}
	}
	if (scan_clip) {
		prof_pre(ctx, st);
		clip_kernel<<<std::max(1, std::min((n + 7) / 8, ctx->sm_count * 8)), 256, 0, st>>>(bursts, stride, n, rc, flags, sch ? nullptr : type,
												      sps1_len ? sps1_len : 625);
		prof_post(ctx, st, "clip_kernel");
		return post_launch(ctx, "clip_kernel");
	}
	return TRXB200_OK;
}

static int launch_demod(trxb200_ctx *ctx, cudaStream_t st, const float *bursts, int stride, int n, int32_t *rc,
			const float *amp, const float *toa, float *ci, uint8_t *flags, float *soft, int soft_stride,
			int n_gmsk_soft, int fix_clip, const uint8_t *type = nullptr, int bps = 0, const int16_t *iq = nullptr,
			int iq_stride = 0, const uint8_t *type_raw = nullptr, float *pw = nullptr, uint8_t *pkt = nullptr,
			int pkt_stride = 0, int pkt_version = 1)
{
	DemodParams p;
	p.type = type;
	p.iq = iq; p.iq_stride = iq_stride; p.type_raw = type_raw; p.pw = pw;
	p.pkt = pkt; p.pkt_stride = pkt_stride; p.pkt_hdr = pkt_version == 1 ? 11 : 8; p.pkt_v0 = pkt_version == 0;
	p.bursts = bursts; p.stride = stride; p.n = n; p.rc = rc; p.amp = amp; p.toa = toa; p.ci = ci; p.flags = flags;
	p.soft = soft; p.soft_stride = soft_stride; p.n_gmsk_soft = n_gmsk_soft; p.comp = ctx->d_comp; p.dnsamp_g = ctx->d_comp + ctx->ht->comp.size(); p.edge_tab = ctx->d_edge_tab; p.fix_clip = fix_clip;
	const bool wide = !iq && bps <= 0 && ctx->tune.demod_wpb == 17; // one CTA of 17 warps per SM (float rows, not while overlapping)
	const int wpb = wide ? 17 : ((!iq && bps <= 0 && ctx->tune.demod_wpb >= 1 && ctx->tune.demod_wpb < 8) ? ctx->tune.demod_wpb : 8);
	const size_t smem = (size_t)wpb * kDemodWarpFloats * sizeof(float);
	if (!ctx->cfg_demod) {
		CK(cudaFuncSetAttribute(demod_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kDemodWarpFloats * sizeof(float))));
		CK(cudaFuncSetAttribute(demod_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kDemodWarpFloats * sizeof(float))));
		CK(cudaFuncSetAttribute(demod_kernel<false, 17, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(17 * kDemodWarpFloats * sizeof(float))));
		ctx->cfg_demod = true;
	}
	if (bps <= 0) bps = ctx->tune.demod_bps;
	int grid = std::min((n + wpb - 1) / wpb, ctx->sm_count * (wide ? 1 : std::max(1, std::min(2, bps))));
	if (grid < 1) grid = 1;
	prof_pre(ctx, st);
	if (iq) demod_kernel<true><<<grid, wpb * 32, smem, st>>>(p);
	else if (wide) demod_kernel<false, 17, 1><<<grid, wpb * 32, smem, st>>>(p);
	else demod_kernel<false><<<grid, wpb * 32, smem, st>>>(p);
	prof_post(ctx, st, "demod_kernel");
	return post_launch(ctx, "demod_kernel");
}

// detect + demodulate in one persistent kernel (nbfused.cu): normal-burst geometry, one detection attempt, float rows
static bool fused_applies(const trxb200_ctx *ctx, int bound) { return ctx->tune.fused && ctx->max_attempts == 1 && nb_geometry(ctx, bound); }

static int launch_nb_fused(trxb200_ctx *ctx, cudaStream_t st, const float *bursts, int stride, int n, const uint8_t *type,
			   const uint8_t *tsc, const uint16_t *max_toa, int bound, float thresh, int32_t *rc, float *amp, float *toa,
			   uint8_t *tsc_out, float *ci, uint8_t *flags, float *soft, int soft_stride, int n_gmsk_soft)
{
	FusedParams fp;
	CorrParams &c = fp.c;
	c.bursts = bursts; c.stride = stride; c.n = n; c.type = type; c.tsc = tsc; c.max_toa = max_toa; c.rc = rc; c.round = 0;
	c.max_toa_bound = bound; c.lmax = 20; c.ndmax = 35; c.corr = nullptr; c.pwr = nullptr; c.negzero = -0.0f; c.iq = nullptr;
	c.iq_stride = 0; c.sch = 0;
	PeakParams &q = fp.q;
	q.n = n; q.type = type; q.tsc = tsc; q.max_toa = max_toa; q.round = 0; q.last_round = 1; q.max_toa_bound = bound; q.thresh = thresh;
	q.lmax = 20; q.ndmax = 35; q.corr = nullptr; q.pwr = nullptr; q.sinc512 = ctx->d_sinc512; q.rc = rc; q.amp = amp; q.toa = toa; q.ci = ci;
	q.tsc_out = tsc_out; q.flags = flags; q.negzero = -0.0f; q.sch = 0;
	DemodParams &d = fp.d;
	d.type = type; d.iq = nullptr; d.iq_stride = 0; d.type_raw = nullptr; d.pw = nullptr; d.pkt = nullptr; d.pkt_stride = 0; d.pkt_hdr = 11;
	d.pkt_v0 = 0; d.bursts = bursts; d.stride = stride; d.n = n; d.rc = rc; d.amp = amp; d.toa = toa; d.ci = ci; d.flags = flags;
	d.soft = soft; d.soft_stride = soft_stride; d.n_gmsk_soft = n_gmsk_soft; d.comp = ctx->d_comp;
	d.dnsamp_g = ctx->d_comp + ctx->ht->comp.size(); d.edge_tab = ctx->d_edge_tab; d.fix_clip = 1;
	if (!ctx->cfg_fused) {
		CK(cudaFuncSetAttribute(nb_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFuSmemBytes));
		ctx->cfg_fused = true;
	}
	const int ntiles = (n + kFuTile - 1) / kFuTile;
	const int grid = std::max(1, std::min(ntiles, ctx->sm_count));
	prof_pre(ctx, st);
	nb_fused_kernel<<<grid, kFuThreads, kFuSmemBytes, st>>>(fp);
	prof_post(ctx, st, "nb_fused_kernel");
	return post_launch(ctx, "nb_fused_kernel");
}

static int check_dd(trxb200_ctx *ctx, const void *bursts, int stride, int n, int bound)
{
	if (!ctx) return TRXB200_EINVAL;
	if (n == 0) return TRXB200_OK; /* empty batch: nothing to validate, nothing to do */
	if (!bursts || stride < 625 || n < 0 || bound < 0 || bound > 1024)
		return fail(ctx, TRXB200_EINVAL, "detect/demod: bad argument");
	return TRXB200_OK;
}

int trxb200_detect_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, const uint8_t *type,
			 const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh, int32_t *rc,
			 float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_dd(ctx, bursts, stride, n, max_toa_bound);
	if (r) return r;
	if (!type || !tsc || !max_toa || !rc || !amp || !toa || !tsc_out || !ci)
		return fail(ctx, TRXB200_EINVAL, "detect: null output");
	if (n == 0) return TRXB200_OK;
	return launch_detect(ctx, ctx->stream, ctx->ws, bursts, stride, n, type, tsc, max_toa, max_toa_bound, thresh, rc, amp, toa,
			     tsc_out, ci, flags, 1);
}

int trxb200_detect_sch_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, float thresh, int32_t *rc, float *amp,
			     float *toa, float *ci, uint8_t *flags)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_dd(ctx, bursts, stride, n, 0);
	if (r) return r;
	if (!rc || !amp || !toa || !ci) return fail(ctx, TRXB200_EINVAL, "detect_sch: null output");
	return launch_detect(ctx, ctx->stream, ctx->ws, bursts, stride, n, nullptr, nullptr, nullptr, 0, thresh, rc, amp, toa, nullptr, ci,
			     flags, 0, nullptr, false, nullptr, 0, true);
}

int trxb200_detect_sch_buffer_batch(trxb200_ctx *ctx, const float *bufs, int stride, int in_len, int n, float thresh, int32_t *rc,
				    float *amp, float *toa, float *ci, uint8_t *flags)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	const int len = in_len / 4;
	// len: Resampler::rotate's MAX_OUTPUT_LEN (Resampler.cpp:35) bounds what the reference can decimate in one call
	if (!bufs || !rc || !amp || !toa || !ci || n < 0 || in_len < 4 * 64 || (in_len & 3) || in_len > stride || len > 4096 * 4)
		return fail(ctx, TRXB200_EINVAL, "detect_sch_buffer: bad argument");
	if (n == 0) return TRXB200_OK;
	cudaStream_t st = ctx->stream;
	const int tiles = (n + 31) / 32;
	const size_t dec_b = (size_t)n * (len + kSchBufDecPad) * sizeof(float2);
	const size_t tile_b = (size_t)tiles * (len + 2 * kPadRows) * kRowPitch * sizeof(float2);
	SchBufDetParams q;
	q.bursts = bufs; q.stride = stride; q.in_len = in_len; q.len = len; q.n = n; q.thresh = thresh; q.sinc512 = ctx->d_sinc512;
	q.rc = rc; q.amp = amp; q.toa = toa; q.ci = ci; q.flags = flags; q.negzero = -0.0f;
	CK(cudaMallocAsync(&q.dec, dec_b, st));
	CK(cudaMallocAsync(&q.ctile, tile_b, st));
	CK(cudaMemsetAsync(q.dec, 0, dec_b, st));
	CK(cudaMemsetAsync(q.ctile, 0, tile_b, st));
	const dim3 grid((len + 255) / 256, n);
	sch_buf_decim_kernel<<<grid, 256, 0, st>>>(q);
	int r = post_launch(ctx, "sch_buf_decim_kernel");
	if (!r) {
		sch_buf_corr_kernel<<<grid, 256, 0, st>>>(q);
		r = post_launch(ctx, "sch_buf_corr_kernel");
	}
	if (!r) {
		sch_buf_peak_kernel<<<tiles, 32, 0, st>>>(q);
		r = post_launch(ctx, "sch_buf_peak_kernel");
	}
	cudaFreeAsync(q.dec, st);
	cudaFreeAsync(q.ctile, st);
	return r;
}

int trxb200_demod_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, const int32_t *rc, const float *amp,
			const float *toa, float *ci, float *soft, int soft_stride, int n_gmsk_soft)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_dd(ctx, bursts, stride, n, 0);
	if (r) return r;
	if (!rc || !amp || !toa || !ci || !soft || n_gmsk_soft < 1 || n_gmsk_soft > 156 || soft_stride < n_gmsk_soft)
		return fail(ctx, TRXB200_EINVAL, "demod: bad argument");
	if (n == 0) return TRXB200_OK;
	return launch_demod(ctx, ctx->stream, bursts, stride, n, const_cast<int32_t *>(rc), amp, toa, ci, nullptr, soft,
			    soft_stride, n_gmsk_soft, 0);
}

/* ---- the 1-sample-per-symbol receive path (rx_sps = 1): detectAnyBurst / demodAnyBurst with sps == 1 ---- */
static int check_sps1(trxb200_ctx *ctx, const void *bursts, int stride, int blen, int n, int bound)
{
	if (!ctx) return TRXB200_EINVAL;
	if (n == 0) return TRXB200_OK;
	// a slot is 156 or 157 symbols (radioInterface.cpp:257-258 at one sample per symbol); 148 is the burst itself
	if (!bursts || blen < 148 || blen > 160 || stride < blen || n < 0 || bound < 0 || bound > 1024)
		return fail(ctx, TRXB200_EINVAL, "detect/demod (1 sps): bad argument");
	return TRXB200_OK;
}

int trxb200_detect_sps1_batch(trxb200_ctx *ctx, const float *bursts, int stride, int blen, int n, const uint8_t *type,
			      const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh, int32_t *rc, float *amp,
			      float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_sps1(ctx, bursts, stride, blen, n, max_toa_bound);
	if (r) return r;
	if (!type || !tsc || !max_toa || !rc || !amp || !toa || !tsc_out || !ci)
		return fail(ctx, TRXB200_EINVAL, "detect (1 sps): null output");
	return launch_detect(ctx, ctx->stream, ctx->ws, bursts, stride, n, type, tsc, max_toa, max_toa_bound, thresh, rc, amp, toa,
			     tsc_out, ci, flags, 1, nullptr, false, nullptr, 0, false, blen);
}

int trxb200_demod_sps1_batch(trxb200_ctx *ctx, const float *bursts, int stride, int blen, int n, const int32_t *rc, const float *amp,
			     const float *toa, float *ci, float *soft, int soft_stride, int n_gmsk_soft)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_sps1(ctx, bursts, stride, blen, n, 0);
	if (r) return r;
	if (!rc || !amp || !toa || !ci || !soft || n_gmsk_soft < 1 || n_gmsk_soft > blen || soft_stride < n_gmsk_soft)
		return fail(ctx, TRXB200_EINVAL, "demod (1 sps): bad argument");
	cudaStream_t st = ctx->stream;
	// delayVector(burst, -toa * sps) (:2038) into a scratch row per burst, then scale / derotate / slice
	float *d = nullptr;
	CK(cudaMallocAsync(&d, (size_t)n * blen * sizeof(float2), st));
	const long tiles = (long)n * ((blen + kCvTile - 1) / kCvTile);
	delay_vector_blk_kernel<<<grid_for(ctx, tiles * 32, 256, 8), 256, 0, st>>>(bursts, stride, blen, n, toa, d, blen, -0.0f, -1.0f);
	r = post_launch(ctx, "delay_vector_blk_kernel");
	if (!r) {
		DemodParams p = DemodParams();
		p.bursts = d; p.stride = blen; p.n = n; p.rc = const_cast<int32_t *>(rc); p.amp = amp; p.toa = toa; p.ci = ci;
		p.soft = soft; p.soft_stride = soft_stride; p.n_gmsk_soft = n_gmsk_soft; p.comp = ctx->d_comp;
		p.dnsamp_g = ctx->d_comp + ctx->ht->comp.size(); p.edge_tab = ctx->d_edge_tab; p.pkt_hdr = 11; p.sps1_len = blen;
		demod1_kernel<<<std::max(1, std::min((n + 7) / 8, ctx->sm_count * 8)), 256, 8 * kScratchFloats * sizeof(float), st>>>(p);
		r = post_launch(ctx, "demod1_kernel");
	}
	cudaFreeAsync(d, st);
	return r;
}

int trxb200_detect_demod_batch(trxb200_ctx *ctx, const float *bursts, int stride, int n, const uint8_t *type,
			       const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh,
			       int32_t *rc, float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags,
			       float *soft, int soft_stride, int n_gmsk_soft)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_dd(ctx, bursts, stride, n, max_toa_bound);
	if (r) return r;
	if (!type || !tsc || !max_toa || !rc || !amp || !toa || !tsc_out || !ci || !soft || n_gmsk_soft < 1 ||
	    n_gmsk_soft > 156 || soft_stride < n_gmsk_soft)
		return fail(ctx, TRXB200_EINVAL, "detect_demod: bad argument");
	if (n == 0) return TRXB200_OK;
	if (fused_applies(ctx, max_toa_bound))
		return launch_nb_fused(ctx, ctx->stream, bursts, stride, n, type, tsc, max_toa, max_toa_bound, thresh, rc, amp, toa, tsc_out, ci,
				       flags, soft, soft_stride, n_gmsk_soft);
	// Small batches, and the per-kernel profiling pass (which wants clean, serialised kernel times): one demod
	// launch after detection.  Otherwise the batch is pipelined chunk by chunk over two streams.
	if (!ctx->tune.overlap || ctx->prof || n <= ctx->tune.chunk_cap) {
		r = launch_detect(ctx, ctx->stream, ctx->ws, bursts, stride, n, type, tsc, max_toa, max_toa_bound, thresh, rc, amp,
				  toa, tsc_out, ci, flags, 0);
		if (r) return r;
		return launch_demod(ctx, ctx->stream, bursts, stride, n, rc, amp, toa, ci, flags, soft, soft_stride, n_gmsk_soft, 1,
				    type);
	}
	struct Hook : ChunkHook {
		trxb200_ctx *ctx;
		const float *bursts; int stride;
		int32_t *rc; float *amp, *toa, *ci; uint8_t *flags; float *soft; int soft_stride, n_gmsk_soft;
		const uint8_t *type;
		int operator()(long lo, int m) override
		{
			cudaEvent_t ev = ctx->pipe_ev[ctx->pipe_ev_next];
			ctx->pipe_ev_next = (ctx->pipe_ev_next + 1) & 3;
			CK(cudaEventRecord(ev, ctx->stream));
			CK(cudaStreamWaitEvent(ctx->side_stream, ev, 0));
			return launch_demod(ctx, ctx->side_stream, bursts + (size_t)lo * stride * 2, stride, m, rc + lo, amp + 2 * lo,
					    toa + lo, ci + lo, flags ? flags + lo : nullptr, soft + (size_t)lo * soft_stride, soft_stride,
					    n_gmsk_soft, 1, type + lo, ctx->tune.ov_demod_bps);
		}
	} hook;
	hook.ctx = ctx; hook.bursts = bursts; hook.stride = stride; hook.rc = rc; hook.amp = amp; hook.toa = toa; hook.ci = ci;
	hook.flags = flags; hook.soft = soft; hook.soft_stride = soft_stride; hook.n_gmsk_soft = n_gmsk_soft; hook.type = type;
	r = launch_detect(ctx, ctx->stream, ctx->ws, bursts, stride, n, type, tsc, max_toa, max_toa_bound, thresh, rc, amp, toa,
			  tsc_out, ci, flags, 0, &hook, true);
	if (r) return r;
	// join: everything enqueued on the caller's stream after this call sees the demodulated results
	cudaEvent_t ev = ctx->pipe_ev[ctx->pipe_ev_next];
	ctx->pipe_ev_next = (ctx->pipe_ev_next + 1) & 3;
	CK(cudaEventRecord(ev, ctx->side_stream));
	CK(cudaStreamWaitEvent(ctx->stream, ev, 0));
	return TRXB200_OK;
}

/* ---------------- host-buffer pipeline ---------------- */
static bool host_ptr_is_pinned(const void *p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
static inline int round16(int m) { return (m + 15) & ~15; }

static void stage_free(HostStage *s)
{
	for (int k = 0; k < HostStage::kSlots; k++) {
		cudaFree(s->d_bursts[k]); cudaFree(s->d_soft[k]); cudaFree(s->d_in[k]); cudaFree(s->d_out[k]);
		cudaFreeHost(s->h_in[k]); cudaFreeHost(s->h_out[k]); cudaFreeHost(s->h_bursts[k]); cudaFreeHost(s->h_soft[k]);
		cudaFree(s->ws[k].corr); cudaFree(s->ws[k].pwr); cudaFree(s->ws[k].list);
		if (s->streams[k]) cudaStreamDestroy(s->streams[k]);
	}
	delete s;
}

// packed per-burst arrays of the float pipeline, m16 = round16(bursts in the chunk):
//   in : max_toa u16 @0 | type u8 @2 m16 | tsc u8 @3 m16                                              ( 4 B per burst)
//   out: rc i32 @0 | amp f32x2 @4 m16 | toa f32 @12 m16 | ci f32 @16 m16 | tsc_out u8 @20 m16 | flags u8 @21 m16 (22 B per burst)
constexpr int kDdInBytes = 4, kDdOutBytes = 22;

static int stage_get(trxb200_ctx *ctx, int stride, int soft_stride, bool stage_in, bool stage_out, HostStage **out)
{
	const int chunk = round16(ctx->tune.host_chunk);
	HostStage *s = ctx->stage;
	if (s && (s->stride != stride || s->soft_stride != soft_stride || s->chunk != chunk)) {
		stage_free(s);
		ctx->stage = s = nullptr;
	}
	cudaError_t e = cudaSuccess;
	auto A = [&](auto **p_, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p_, bytes); };
	auto H = [&](auto **p_, size_t bytes) { if (e == cudaSuccess && !*p_) e = cudaMallocHost(p_, bytes); };
	if (!s) {
		// built locally and published only when complete: a failed allocation must not leave a half-built stage cached
		s = new HostStage();
		s->chunk = chunk; s->stride = stride; s->soft_stride = soft_stride;
		for (int k = 0; k < HostStage::kSlots && e == cudaSuccess; k++) {
			e = cudaStreamCreateWithFlags(&s->streams[k], cudaStreamNonBlocking);
			A(&s->d_bursts[k], (size_t)chunk * stride * 8);
			A(&s->d_soft[k], (size_t)chunk * soft_stride * 4);
			A(&s->d_in[k], (size_t)chunk * kDdInBytes); A(&s->d_out[k], (size_t)chunk * kDdOutBytes);
			H(&s->h_in[k], (size_t)chunk * kDdInBytes); H(&s->h_out[k], (size_t)chunk * kDdOutBytes);
		}
		if (e != cudaSuccess) {
			stage_free(s);
			return fail(ctx, e == cudaErrorMemoryAllocation ? TRXB200_ENOMEM : TRXB200_ECUDA, "detect_demod_host: staging buffers", e);
		}
		ctx->stage = s;
	}
	for (int k = 0; k < HostStage::kSlots; k++) {
		if (stage_in) H(&s->h_bursts[k], (size_t)chunk * stride * 8);
		if (stage_out) H(&s->h_soft[k], (size_t)chunk * soft_stride * 4);
	}
	if (e != cudaSuccess) return fail(ctx, TRXB200_ENOMEM, "detect_demod_host: pinned staging", e);
	*out = s;
	return TRXB200_OK;
}

int trxb200_detect_demod_host(trxb200_ctx *ctx, const float *bursts, int stride, int n, const uint8_t *type,
			      const uint8_t *tsc, const uint16_t *max_toa, int max_toa_bound, float thresh,
			      int32_t *rc, float *amp, float *toa, uint8_t *tsc_out, float *ci, uint8_t *flags,
			      float *soft, int soft_stride, int n_gmsk_soft)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (ctx && n == 0) return TRXB200_OK;
	int r = check_dd(ctx, bursts, stride, n, max_toa_bound);
	if (r) return r;
	if (!type || !tsc || !max_toa || !rc || !amp || !toa || !tsc_out || !ci || !soft || n_gmsk_soft < 1 ||
	    n_gmsk_soft > 156 || soft_stride < n_gmsk_soft)
		return fail(ctx, TRXB200_EINVAL, "detect_demod_host: bad argument");
	if (n == 0) return TRXB200_OK;
	const bool pin_in = host_ptr_is_pinned(bursts), pin_out = host_ptr_is_pinned(soft);
	HostStage *s = nullptr;
	r = stage_get(ctx, stride, soft_stride, !pin_in, !pin_out, &s);
	if (r) return r;
	// small results of the chunk a slot carried last: wait for its copies, scatter to the caller's arrays
	auto finish = [&](int slot) -> int {
		CK(cudaStreamSynchronize(s->streams[slot]));
		const int m = s->pend_m[slot], lo = s->pend_lo[slot];
		if (!m) return TRXB200_OK;
		const int m16 = round16(m);
		const uint8_t *o = s->h_out[slot];
		std::memcpy(rc + lo, o, (size_t)m * 4);
		std::memcpy(amp + (size_t)lo * 2, o + (size_t)4 * m16, (size_t)m * 8);
		std::memcpy(toa + lo, o + (size_t)12 * m16, (size_t)m * 4);
		std::memcpy(ci + lo, o + (size_t)16 * m16, (size_t)m * 4);
		std::memcpy(tsc_out + lo, o + (size_t)20 * m16, m);
		if (flags) std::memcpy(flags + lo, o + (size_t)21 * m16, m);
		if (!pin_out) std::memcpy(soft + (size_t)lo * soft_stride, s->h_soft[slot], (size_t)m * soft_stride * 4);
		s->pend_m[slot] = 0;
		return TRXB200_OK;
	};
	int slot = 0;
	for (int lo = 0; lo < n; lo += s->chunk, slot = (slot + 1) % HostStage::kSlots) {
		const int m = std::min(s->chunk, n - lo), m16 = round16(m);
		cudaStream_t st = s->streams[slot];
		r = finish(slot); // the slot's previous chunk (incl. its D2H) has fully landed
		if (r) return r;
		uint8_t *hi = s->h_in[slot];
		std::memcpy(hi, max_toa + lo, (size_t)m * 2);
		std::memcpy(hi + (size_t)2 * m16, type + lo, m);
		std::memcpy(hi + (size_t)3 * m16, tsc + lo, m);
		const float *src = bursts + (size_t)lo * stride * 2;
		if (!pin_in) {
			std::memcpy(s->h_bursts[slot], src, (size_t)m * stride * 8);
			src = s->h_bursts[slot];
		}
		CK(cudaMemcpyAsync(s->d_bursts[slot], src, (size_t)m * stride * 8, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(s->d_in[slot], hi, (size_t)m16 * kDdInBytes, cudaMemcpyHostToDevice, st));
		uint8_t *di = s->d_in[slot], *d_o = s->d_out[slot];
		const uint16_t *d_max_toa = reinterpret_cast<const uint16_t *>(di);
		const uint8_t *d_type = di + (size_t)2 * m16, *d_tsc = di + (size_t)3 * m16;
		int32_t *d_rc = reinterpret_cast<int32_t *>(d_o);
		float *d_amp = reinterpret_cast<float *>(d_o + (size_t)4 * m16), *d_toa = reinterpret_cast<float *>(d_o + (size_t)12 * m16);
		float *d_ci = reinterpret_cast<float *>(d_o + (size_t)16 * m16);
		uint8_t *d_tsc_out = d_o + (size_t)20 * m16, *d_flags = d_o + (size_t)21 * m16;
		// rows of undetected bursts are never written by the kernels: define them as zero for host callers
		CK(cudaMemsetAsync(s->d_soft[slot], 0, (size_t)m * soft_stride * 4, st));
		if (fused_applies(ctx, max_toa_bound)) {
			r = launch_nb_fused(ctx, st, s->d_bursts[slot], stride, m, d_type, d_tsc, d_max_toa, max_toa_bound, thresh, d_rc, d_amp,
					    d_toa, d_tsc_out, d_ci, d_flags, s->d_soft[slot], soft_stride, n_gmsk_soft);
			if (r) return r;
		} else {
			r = launch_detect(ctx, st, s->ws[slot], s->d_bursts[slot], stride, m, d_type, d_tsc, d_max_toa, max_toa_bound, thresh, d_rc,
					  d_amp, d_toa, d_tsc_out, d_ci, d_flags, 0);
			if (r) return r;
			r = launch_demod(ctx, st, s->d_bursts[slot], stride, m, d_rc, d_amp, d_toa, d_ci, d_flags, s->d_soft[slot], soft_stride,
					 n_gmsk_soft, 1, d_type);
			if (r) return r;
		}
		CK(cudaMemcpyAsync(pin_out ? soft + (size_t)lo * soft_stride : s->h_soft[slot], s->d_soft[slot], (size_t)m * soft_stride * 4,
				   cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(s->h_out[slot], d_o, (size_t)m16 * kDdOutBytes, cudaMemcpyDeviceToHost, st));
		s->pend_lo[slot] = lo; s->pend_m[slot] = m;
	}
	for (int k = 0; k < HostStage::kSlots; k++) {
		r = finish(k);
		if (r) return r;
	}
	return TRXB200_OK;
}

/* ---------------- pull path: int16 slots -> TRXD uplink datagrams ---------------- */
static void pull_scratch_free(PullScratch &w)
{
	cudaFree(w.bursts); cudaFree(w.pw); cudaFree(w.type2); cudaFree(w.tsc_out); cudaFree(w.amp); cudaFree(w.toa);
	cudaFree(w.ci); cudaFree(w.ws.corr); cudaFree(w.ws.pwr); cudaFree(w.ws.list);
	w = PullScratch();
}

// part of a slot the detection kernels can read: correlator windows of every burst type the configuration allows
// (make_attempt in detect.cu: 16-symbol sequences start at sample 209, access bursts at -15; all end by 345 + 4 T)
static void pull_window(const trxb200_ctx *ctx, int bound, int &s_min, int &W)
{
	s_min = ctx->max_seq_len >= 40 ? 0 : 208;
	const int s_max = std::min(624, 345 + 4 * bound + 4);
	W = ((s_max - s_min) + 3) & ~3;
}

static int pull_scratch_get(trxb200_ctx *ctx, cudaStream_t st, PullScratch &w, int cap, int soft_stride, int win)
{
	if (w.cap >= cap && w.soft_stride == soft_stride && w.win == win) return TRXB200_OK;
	CK(cudaStreamSynchronize(st));
	DetectScratch keep = w.ws;
	w.ws = DetectScratch();
	pull_scratch_free(w);
	w.ws = keep;
	CK(cudaMalloc(&w.bursts, (size_t)cap * win * 8));
	CK(cudaMalloc(&w.pw, (size_t)cap * 80 * 4));
	CK(cudaMalloc(&w.type2, cap)); CK(cudaMalloc(&w.tsc_out, cap));
	CK(cudaMalloc(&w.amp, (size_t)cap * 8)); CK(cudaMalloc(&w.toa, (size_t)cap * 4)); CK(cudaMalloc(&w.ci, (size_t)cap * 4));
	w.cap = cap; w.soft_stride = soft_stride; w.win = win;
	return TRXB200_OK;
}

static int pull_check(trxb200_ctx *ctx, const trxb200_pull_args *a)
{
	if (!ctx) return TRXB200_EINVAL;
	if (!a) return fail(ctx, TRXB200_EINVAL, "pull: null argument block");
	if (a->n == 0) return TRXB200_OK;
	const int hdr = a->trxd_version == 1 ? 11 : 10;
	if (a->n < 0 || !a->iq || a->stride < 625 || !a->type || !a->tsc || !a->max_toa || !a->fn || !a->tn || !a->rc || !a->energy ||
	    !a->pkt || !a->pkt_len || (a->trxd_version != 0 && a->trxd_version != 1) || a->pkt_stride < hdr + 148 ||
	    a->max_toa_bound < 0 || a->max_toa_bound > 1024)
		return fail(ctx, TRXB200_EINVAL, "pull: bad argument");
	return TRXB200_OK;
}

// soft-bit row the datagram rows can hold: 444 (8-PSK capable) or 148
static int pull_soft_stride(const trxb200_pull_args *a)
{
	const int room = a->pkt_stride - (a->trxd_version == 1 ? 11 : 10);
	return room >= 444 ? 444 : 148;
}

// m slots, all pointers on the device and already offset to the first slot of the chunk
static int pull_chunk(trxb200_ctx *ctx, cudaStream_t st, PullScratch &w, const trxb200_pull_args *a, int m, const int16_t *iq,
		      const uint8_t *type, const uint8_t *tsc, const uint16_t *max_toa, const uint32_t *fn, const uint8_t *tn,
		      int32_t *rc, float *energy, uint8_t *pkt, uint16_t *pkt_len, uint8_t *flags, float *amp, float *toa, float *ci,
		      uint8_t *tsc_out)
{
	int s_min, W;
	pull_window(ctx, a->max_toa_bound, s_min, W);
	// 16-symbol sequences at max_toa <= 4 (launch_detect's corr_nb_kernel case): the correlator reads the int16 slots
	// itself and no window is extracted; otherwise the windows are converted for corr_long_kernel
	const bool nb = (ctx->max_seq_len + 16 + a->max_toa_bound - 1 == 35) && (((16 + a->max_toa_bound + 1) & ~1) == 20);
	ExtractParams ip;
	ip.iq = iq; ip.stride_in = a->stride; ip.n = m; ip.type = type; ip.type_out = w.type2; ip.win = w.bursts; ip.W = nb ? 0 : W;
	ip.s_min = s_min;
	prof_pre(ctx, st);
	extract_kernel<<<std::max(1, std::min((m + 7) / 8, ctx->sm_count * 8)), 256, 0, st>>>(ip);
	prof_post(ctx, st, "extract_kernel");
	int r = post_launch(ctx, "extract_kernel");
	if (r) return r;
	if (!amp) amp = w.amp;
	if (!toa) toa = w.toa;
	if (!ci) ci = w.ci;
	if (!tsc_out) tsc_out = w.tsc_out;
	if (nb) {
		r = launch_detect(ctx, st, w.ws, nullptr, 0, m, w.type2, tsc, max_toa, a->max_toa_bound, a->thresh, rc, amp, toa, tsc_out,
				  ci, flags, 0, nullptr, false, iq, a->stride);
	} else {
		// the detection kernels index samples of the full slot: hand them the window buffer shifted back by s_min samples
		const float *det_rows = w.bursts - (ptrdiff_t)2 * s_min;
		r = launch_detect(ctx, st, w.ws, det_rows, W, m, w.type2, tsc, max_toa, a->max_toa_bound, a->thresh, rc, amp, toa,
				  tsc_out, ci, flags, 0);
	}
	if (r) return r;
	r = launch_demod(ctx, st, nullptr, 0, m, rc, amp, toa, ci, flags, nullptr, 0, 148, 1, w.type2, 0, iq, a->stride, type, w.pw, pkt,
			 a->pkt_stride, a->trxd_version);
	if (r) return r;
	HeaderParams hp;
	hp.n = m; hp.version = a->trxd_version; hp.type = type; hp.rc = rc; hp.toa = toa; hp.ci = ci; hp.pw = w.pw; hp.energy = energy;
	hp.tsc_out = tsc_out; hp.fn = fn; hp.tn = tn; hp.full_scale = a->rx_full_scale; hp.rssi_offset = a->rssi_offset; hp.pkt = pkt;
	hp.pkt_stride = a->pkt_stride; hp.pkt_len = pkt_len; hp.flags = flags;
	prof_pre(ctx, st);
	header_kernel<<<std::max(1, std::min((m + 32 * kHdrWarps - 1) / (32 * kHdrWarps), ctx->sm_count * 8)), kHdrWarps * 32, 0, st>>>(hp);
	prof_post(ctx, st, "header_kernel");
	return post_launch(ctx, "header_kernel");
}

int trxb200_pull_batch(trxb200_ctx *ctx, const trxb200_pull_args *a)
{
	DevGuard dg(ctx ? ctx->device : -1);
	int r = pull_check(ctx, a);
	if (r || a->n == 0) return r;
	const int cap = ctx->tune.pull_chunk;
	int s_min_, W_;
	pull_window(ctx, a->max_toa_bound, s_min_, W_);
	r = pull_scratch_get(ctx, ctx->stream, ctx->pull, std::min(cap, a->n), pull_soft_stride(a), W_);
	if (r) return r;
	for (long lo = 0; lo < a->n; lo += ctx->pull.cap) {
		const int m = (int)std::min<long>(ctx->pull.cap, a->n - lo);
		r = pull_chunk(ctx, ctx->stream, ctx->pull, a, m, a->iq + (size_t)lo * a->stride * 2, a->type + lo, a->tsc + lo,
			       a->max_toa + lo, a->fn + lo, a->tn + lo, a->rc + lo, a->energy + lo, a->pkt + (size_t)lo * a->pkt_stride,
			       a->pkt_len + lo, a->flags ? a->flags + lo : nullptr, a->amp ? a->amp + (size_t)lo * 2 : nullptr,
			       a->toa ? a->toa + lo : nullptr, a->ci ? a->ci + lo : nullptr, a->tsc_out ? a->tsc_out + lo : nullptr);
		if (r) return r;
	}
	return TRXB200_OK;
}

int trxb200_expected_corr_type_batch(trxb200_ctx *ctx, const trxb200_sched_cfg *cfg, const uint32_t *fn, const uint8_t *tn,
				     const uint16_t *chan, int n, uint8_t *type, uint16_t *max_toa)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx) return TRXB200_EINVAL;
	if (!cfg || n < 0 || cfg->n_chan < 1 || !cfg->chan_type || !cfg->handover || (n > 0 && (!fn || !tn || !type)))
		return fail(ctx, TRXB200_EINVAL, "expected_corr_type: bad argument");
	if (n == 0) return TRXB200_OK;
	if (!ctx->sched_tables) {
		// SDCCH/4 and SDCCH/8 sub-slot per 102-multiframe position (Transceiver.cpp:517-520), as (value, run length)
		static const unsigned char sd4_rl[] = { 3,4, 0,2, 2,4, 3,4, 0,27, 1,4, 0,2, 2,4, 3,4, 0,6, 1,4, 0,27, 1,4, 0,2, 2,4 };
		static const unsigned char sd8_rl[] = { 5,4, 6,4, 7,4, 0,7, 1,4, 2,4, 3,4, 4,4, 5,4, 6,4, 7,4, 0,4,
							1,4, 2,4, 3,4, 0,7, 1,4, 2,4, 3,4, 4,4, 5,4, 6,4, 7,4, 4,4 };
		unsigned char t4[102], t8[102];
		int k = 0;
		for (size_t i = 0; i < sizeof(sd4_rl); i += 2) for (int c = 0; c < sd4_rl[i + 1]; c++) t4[k++] = sd4_rl[i];
		if (k != 102) return fail(ctx, TRXB200_EINVAL, "expected_corr_type: internal table");
		k = 0;
		for (size_t i = 0; i < sizeof(sd8_rl); i += 2) for (int c = 0; c < sd8_rl[i + 1]; c++) t8[k++] = sd8_rl[i];
		if (k != 102) return fail(ctx, TRXB200_EINVAL, "expected_corr_type: internal table");
		CK(cudaMemcpyToSymbol(c_sd4, t4, 102));
		CK(cudaMemcpyToSymbol(c_sd8, t8, 102));
		ctx->sched_tables = true;
	}
	SchedParams p;
	p.n = n; p.n_chan = cfg->n_chan; p.fn = fn; p.tn = tn; p.chan = chan; p.chan_type = cfg->chan_type; p.handover = cfg->handover;
	p.ext_rach = cfg->ext_rach; p.egprs = cfg->egprs; p.max_toa_nb = cfg->max_toa_nb; p.max_toa_ab = cfg->max_toa_ab;
	p.type_out = type; p.max_toa_out = max_toa;
	sched_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(p);
	return post_launch(ctx, "sched_kernel");
}

struct PullStage {
	static constexpr int kSlots = 3;
	int chunk = 0, stride = 0, pkt_stride = 0;
	cudaStream_t streams[kSlots] = {};
	PullScratch w[kSlots];
	int16_t *d_iq[kSlots] = {};
	uint8_t *d_pkt[kSlots] = {};
	uint8_t *d_in[kSlots] = {}, *d_out[kSlots] = {}; // packed small arrays (device)
	uint8_t *h_in[kSlots] = {}, *h_out[kSlots] = {}; // their pinned mirrors
	int16_t *h_iq[kSlots] = {};			   // pinned staging, allocated only for pageable callers
	uint8_t *h_pkt[kSlots] = {};
	int pend_lo[kSlots] = {}, pend_m[kSlots] = {};
};
// packed per-slot arrays of the pull pipeline, m16 = round16(slots in the chunk):
//   in : fn u32 @0 | max_toa u16 @4 m16 | type u8 @6 m16 | tsc u8 @7 m16 | tn u8 @8 m16                  ( 9 B per slot)
//   out: rc i32 @0 | energy f32 @4 m16 | pkt_len u16 @8 m16 | flags u8 @10 m16 | tsc_out u8 @11 m16     (12 B per slot)
//        | amp f32x2 @12 m16 | toa f32 @20 m16 | ci f32 @24 m16   (copied back only when the caller asks for one of them)
constexpr int kPullInBytes = 9, kPullOutBytes = 12, kPullOutBytesAll = 28;

static void pull_stage_free(PullStage *s)
{
	for (int k = 0; k < PullStage::kSlots; k++) {
		pull_scratch_free(s->w[k]);
		cudaFree(s->d_iq[k]); cudaFree(s->d_pkt[k]); cudaFree(s->d_in[k]); cudaFree(s->d_out[k]);
		cudaFreeHost(s->h_in[k]); cudaFreeHost(s->h_out[k]); cudaFreeHost(s->h_iq[k]); cudaFreeHost(s->h_pkt[k]);
		if (s->streams[k]) cudaStreamDestroy(s->streams[k]);
	}
	delete s;
}

int trxb200_pull_host(trxb200_ctx *ctx, const trxb200_pull_args *a)
{
	DevGuard dg(ctx ? ctx->device : -1);
	int r = pull_check(ctx, a);
	if (r || a->n == 0) return r;
	const int chunk = round16(ctx->tune.host_chunk);
	const bool pin_in = host_ptr_is_pinned(a->iq), pin_out = host_ptr_is_pinned(a->pkt);
	PullStage *s = ctx->pull_stage;
	if (s && (s->stride != a->stride || s->pkt_stride != a->pkt_stride || s->chunk != chunk)) {
		pull_stage_free(s);
		ctx->pull_stage = s = nullptr;
	}
	{
		cudaError_t e = cudaSuccess;
		auto A = [&](auto **p_, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p_, bytes); };
		auto H = [&](auto **p_, size_t bytes) { if (e == cudaSuccess && !*p_) e = cudaMallocHost(p_, bytes); };
		if (!s) {
			// built locally and published only when complete (a failed allocation leaves nothing half-built behind)
			s = new PullStage();
			s->chunk = chunk; s->stride = a->stride; s->pkt_stride = a->pkt_stride;
			for (int k = 0; k < PullStage::kSlots && e == cudaSuccess; k++) {
				e = cudaStreamCreateWithFlags(&s->streams[k], cudaStreamNonBlocking);
				A(&s->d_iq[k], (size_t)chunk * a->stride * 4);
				A(&s->d_pkt[k], (size_t)chunk * a->pkt_stride);
				A(&s->d_in[k], (size_t)chunk * kPullInBytes); A(&s->d_out[k], (size_t)chunk * kPullOutBytesAll);
				H(&s->h_in[k], (size_t)chunk * kPullInBytes); H(&s->h_out[k], (size_t)chunk * kPullOutBytesAll);
			}
			if (e != cudaSuccess) {
				pull_stage_free(s);
				return fail(ctx, e == cudaErrorMemoryAllocation ? TRXB200_ENOMEM : TRXB200_ECUDA, "pull_host: staging buffers", e);
			}
			ctx->pull_stage = s;
		}
		for (int k = 0; k < PullStage::kSlots; k++) {
			if (!pin_in) H(&s->h_iq[k], (size_t)chunk * a->stride * 4);
			if (!pin_out) H(&s->h_pkt[k], (size_t)chunk * a->pkt_stride);
		}
		if (e != cudaSuccess) return fail(ctx, TRXB200_ENOMEM, "pull_host: pinned staging", e);
	}
	const bool want_ebp = a->amp || a->toa || a->ci;
	const int out_bytes = want_ebp ? kPullOutBytesAll : kPullOutBytes;
	auto finish = [&](int slot) -> int {
		CK(cudaStreamSynchronize(s->streams[slot]));
		const int m = s->pend_m[slot], lo = s->pend_lo[slot];
		if (!m) return TRXB200_OK;
		const int m16 = round16(m);
		const uint8_t *o = s->h_out[slot];
		std::memcpy(a->rc + lo, o, (size_t)m * 4);
		std::memcpy(a->energy + lo, o + (size_t)4 * m16, (size_t)m * 4);
		std::memcpy(a->pkt_len + lo, o + (size_t)8 * m16, (size_t)m * 2);
		if (a->flags) std::memcpy(a->flags + lo, o + (size_t)10 * m16, m);
		if (a->tsc_out) std::memcpy(a->tsc_out + lo, o + (size_t)11 * m16, m);
		if (a->amp) std::memcpy(a->amp + (size_t)lo * 2, o + (size_t)12 * m16, (size_t)m * 8);
		if (a->toa) std::memcpy(a->toa + lo, o + (size_t)20 * m16, (size_t)m * 4);
		if (a->ci) std::memcpy(a->ci + lo, o + (size_t)24 * m16, (size_t)m * 4);
		if (!pin_out) std::memcpy(a->pkt + (size_t)lo * a->pkt_stride, s->h_pkt[slot], (size_t)m * a->pkt_stride);
		s->pend_m[slot] = 0;
		return TRXB200_OK;
	};
	const int ss = pull_soft_stride(a);
	int s_min_, W_;
	pull_window(ctx, a->max_toa_bound, s_min_, W_);
	int slot = 0;
	for (int lo = 0; lo < a->n; lo += chunk, slot = (slot + 1) % PullStage::kSlots) {
		const int m = std::min(chunk, a->n - lo), m16 = round16(m);
		cudaStream_t st = s->streams[slot];
		r = finish(slot); // the slot's previous chunk (incl. its D2H) has fully landed
		if (r) return r;
		r = pull_scratch_get(ctx, st, s->w[slot], chunk, ss, W_);
		if (r) return r;
		PullScratch &w = s->w[slot];
		uint8_t *hi = s->h_in[slot];
		std::memcpy(hi, a->fn + lo, (size_t)m * 4);
		std::memcpy(hi + (size_t)4 * m16, a->max_toa + lo, (size_t)m * 2);
		std::memcpy(hi + (size_t)6 * m16, a->type + lo, m);
		std::memcpy(hi + (size_t)7 * m16, a->tsc + lo, m);
		std::memcpy(hi + (size_t)8 * m16, a->tn + lo, m);
		const int16_t *src = a->iq + (size_t)lo * a->stride * 2;
		if (!pin_in) {
			std::memcpy(s->h_iq[slot], src, (size_t)m * a->stride * 4);
			src = s->h_iq[slot];
		}
		CK(cudaMemcpyAsync(s->d_iq[slot], src, (size_t)m * a->stride * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(s->d_in[slot], hi, (size_t)m16 * kPullInBytes, cudaMemcpyHostToDevice, st));
		uint8_t *di = s->d_in[slot], *d_o = s->d_out[slot];
		const uint32_t *d_fn = reinterpret_cast<const uint32_t *>(di);
		const uint16_t *d_max_toa = reinterpret_cast<const uint16_t *>(di + (size_t)4 * m16);
		const uint8_t *d_type = di + (size_t)6 * m16, *d_tsc = di + (size_t)7 * m16, *d_tn = di + (size_t)8 * m16;
		int32_t *d_rc = reinterpret_cast<int32_t *>(d_o);
		float *d_energy = reinterpret_cast<float *>(d_o + (size_t)4 * m16);
		uint16_t *d_pkt_len = reinterpret_cast<uint16_t *>(d_o + (size_t)8 * m16);
		uint8_t *d_flags = d_o + (size_t)10 * m16, *d_tsc_out = d_o + (size_t)11 * m16;
		float *d_amp = reinterpret_cast<float *>(d_o + (size_t)12 * m16), *d_toa = reinterpret_cast<float *>(d_o + (size_t)20 * m16);
		float *d_ci = reinterpret_cast<float *>(d_o + (size_t)24 * m16);
		// datagram rows of slots that emit nothing are never written by the kernels: defined as zero for host callers
		CK(cudaMemsetAsync(s->d_pkt[slot], 0, (size_t)m * a->pkt_stride, st));
		r = pull_chunk(ctx, st, w, a, m, s->d_iq[slot], d_type, d_tsc, d_max_toa, d_fn, d_tn, d_rc, d_energy, s->d_pkt[slot], d_pkt_len,
			       d_flags, d_amp, d_toa, d_ci, d_tsc_out);
		if (r) return r;
		CK(cudaMemcpyAsync(pin_out ? a->pkt + (size_t)lo * a->pkt_stride : s->h_pkt[slot], s->d_pkt[slot], (size_t)m * a->pkt_stride,
				   cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(s->h_out[slot], d_o, (size_t)m16 * out_bytes, cudaMemcpyDeviceToHost, st));
		s->pend_lo[slot] = lo; s->pend_m[slot] = m;
	}
	for (int k = 0; k < PullStage::kSlots; k++) {
		r = finish(k);
		if (r) return r;
	}
	return TRXB200_OK;
}

/* ---------------- helpers ---------------- */
int trxb200_energy_detect_batch(trxb200_ctx *ctx, const float *bursts, int stride, int blen, int n, unsigned window,
				float *energy)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !bursts || !energy || n < 0 || blen < 1 || stride < blen) return fail(ctx, TRXB200_EINVAL, "energy_detect: bad argument");
	if (window > (unsigned)blen) window = blen;
	if (window && 4 * (size_t)(window - 1) >= (size_t)stride) return fail(ctx, TRXB200_EINVAL, "energy_detect: window*4 exceeds the row");
	if (n == 0) return TRXB200_OK;
	energy_detect_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(bursts, stride, blen, n, window, energy);
	return post_launch(ctx, "energy_detect_kernel");
}

int trxb200_vector_slicer(trxb200_ctx *ctx, float *dst, const float *src, size_t len)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !dst || !src) return fail(ctx, TRXB200_EINVAL, "vector_slicer: bad argument");
	if (len == 0) return TRXB200_OK;
	const int vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
	vector_slicer_kernel<<<grid_for(ctx, (long)len, 256 * 8, 8), 256, 0, ctx->stream>>>(dst, src, len, vec);
	return post_launch(ctx, "vector_slicer_kernel");
}

int trxb200_delay_vector_batch(trxb200_ctx *ctx, const float *in, int stride, int len, int n, const float *delay,
			       float *out, int out_stride)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !in || !out || !delay || len < 1 || stride < len || out_stride < len || n < 0)
		return fail(ctx, TRXB200_EINVAL, "delay_vector: bad argument");
	if (n == 0) return TRXB200_OK;
	const long tiles = (long)n * ((len + kCvTile - 1) / kCvTile);
	if (tiles < (1L << 30)) {
		delay_vector_blk_kernel<<<grid_for(ctx, tiles * 32, 256, 8), 256, 0, ctx->stream>>>(in, stride, len, n, delay, out, out_stride, -0.0f);
		return post_launch(ctx, "delay_vector_blk_kernel");
	}
	delay_vector_kernel<<<std::min(n, ctx->sm_count * 8), 256, 0, ctx->stream>>>(in, stride, len, n, delay, out, out_stride);
	return post_launch(ctx, "delay_vector_kernel");
}

static int conv_common(trxb200_ctx *ctx, const float *x, int x_len, int x_stride, const float *h, int h_len, float *y,
		       int y_len, int y_stride, int start, int len, int n, int mode, bool always_check)
{
	if (!ctx || !x || !h || !y || n < 0) return fail(ctx, TRXB200_EINVAL, "convolve: bad argument");
	// bounds_check (convolve_base.c:114-131); the SSE entry points skip it in optimised builds, but reading
	// outside the row would fault on a device, so it is always enforced here except for the head-room.
	(void)always_check;
	if (x_len < 1 || h_len < 1 || y_len < 1 || len < 1) return TRXB200_EBOUNDS;
	if (start + len > x_len || len > y_len || x_len < h_len) return TRXB200_EBOUNDS;
	if (h_len > 4096) return fail(ctx, TRXB200_EINVAL, "convolve: h_len > 4096");
	if (n == 0) return TRXB200_OK;
	if ((mode == 0 || mode == 1) && !(h_len % 4) && h_len <= 24 && (long)n * ((len + kCvTile - 1) / kCvTile) < (1L << 30)) {
		// SSE-order cases the reference actually runs (decimator 16, fractional delay 20, pulse shapes 4..24)
		const long tiles = (long)n * ((len + kCvTile - 1) / kCvTile);
		const int grid = grid_for(ctx, tiles * 32, 256, 8);
		const bool ok = mode ? launch_convolve_blk<true>(h_len, grid, ctx->stream, x, x_stride, h, y, y_stride, start, len, n)
				     : launch_convolve_blk<false>(h_len, grid, ctx->stream, x, x_stride, h, y, y_stride, start, len, n);
		if (ok) return post_launch(ctx, "convolve_blk_kernel");
	}
	const size_t smem = (size_t)h_len * 8;
	convolve_kernel<<<grid_for(ctx, (long)n * len, 256, 8), 256, smem, ctx->stream>>>(x, x_stride, h, h_len, y, y_stride, start, len, n, mode);
	return post_launch(ctx, "convolve_kernel");
}

int trxb200_convolve_real_batch(trxb200_ctx *ctx, const float *x, int x_len, int x_stride, const float *h, int h_len,
				float *y, int y_len, int y_stride, int start, int len, int n, int base)
{
	return conv_common(ctx, x, x_len, x_stride, h, h_len, y, y_len, y_stride, start, len, n, base ? 2 : 0, base != 0);
}

int trxb200_convolve_complex_batch(trxb200_ctx *ctx, const float *x, int x_len, int x_stride, const float *h, int h_len,
				   float *y, int y_len, int y_stride, int start, int len, int n, int base)
{
	return conv_common(ctx, x, x_len, x_stride, h, h_len, y, y_len, y_stride, start, len, n, base ? 3 : 1, base != 0);
}

int trxb200_convert_float_short_mode(trxb200_ctx *ctx, int16_t *out, const float *in, float scale, size_t len, int mode)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !out || !in || mode < 0 || mode > 2) return fail(ctx, TRXB200_EINVAL, "convert: bad argument");
	if (len == 0) return TRXB200_OK;
	if (mode == 1 && !(len & 7)) mode = 0;
	const int vec = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in)) & 15) == 0;
	convert_float_short_kernel<<<grid_for(ctx, (long)len, 256 * 8, 8), 256, 0, ctx->stream>>>(out, in, scale, len, vec, mode);
	return post_launch(ctx, "convert_float_short_kernel");
}

int trxb200_convert_float_short(trxb200_ctx *ctx, int16_t *out, const float *in, float scale, size_t len)
{
	return trxb200_convert_float_short_mode(ctx, out, in, scale, len, 0);
}

int trxb200_convert_short_float(trxb200_ctx *ctx, float *out, const int16_t *in, size_t len)
{
	DevGuard dg(ctx ? ctx->device : -1);
	if (!ctx || !out || !in) return fail(ctx, TRXB200_EINVAL, "convert: bad argument");
	if (len == 0) return TRXB200_OK;
	const int vec = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in)) & 15) == 0;
	convert_short_float_kernel<<<grid_for(ctx, (long)len, 256 * 8, 8), 256, 0, ctx->stream>>>(out, in, len, vec);
	return post_launch(ctx, "convert_short_float_kernel");
}

} // extern "C"

#include "capi_stream.cu" // vitac / resampler / filterbank entry points
