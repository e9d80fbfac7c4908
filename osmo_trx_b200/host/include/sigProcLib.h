// sigProcLib.h — the reference's per-burst C++ API (Transceiver52M/sigProcLib.h:57-152) re-created on top of
// the batched C ABI of libtrxb200.so (include/trxb200.h).  Same names, argument meaning, ownership (returned
// vectors are new-allocated, caller deletes) and error conventions, so Transceiver.cpp:107-123,207,392-396,
// 725,768-803 and utils/va-test/burst-gen.cpp compile against it unchanged.  Every call is a batch of one on
// the GPU: this layer is for drop-in correctness; throughput comes from calling the batched ABI directly
// (or the *Batch helpers at the bottom).  There is no CPU implementation behind any of these functions:
// without a usable sm_100 device sigProcLibSetup() returns false and everything else fails.
#pragma once
#include "BitVector.h"
#include "Complex.h"
#include "Vector.h"
#include "signalVector.h"
#include <cstdint>
#include <vector>

#define NORMAL_BURST_NBITS 148
#define EDGE_BURST_NBITS 444
#define EDGE_BURST_NSYMS (EDGE_BURST_NBITS / 3)

// sigProcLib.h:29-45
enum CorrType { OFF, TSC, EXT_RACH, RACH, SCH, EDGE, IDLE };
enum SignalError { SIGERR_NONE, SIGERR_BOUNDS, SIGERR_CLIP, SIGERR_UNSUPPORTED, SIGERR_INTERNAL };
#define BURST_THRESH 4.0 // sigProcLib.h:54

// sigProcLib.h:113-118
struct estim_burst_params {
	complex amp;
	float toa;
	uint8_t tsc;
	float ci;
};

enum class sch_detect_type { SCH_DETECT_FULL, SCH_DETECT_NARROW, SCH_DETECT_BUFFER };

bool sigProcLibSetup();	    // sigProcLib.cpp:2139: builds the tables on the host, uploads them, opens the device
void sigProcLibDestroy(void); // sigProcLib.cpp:137

void vectorSlicer(float *dest, const float *src, size_t len);							     // :546
signalVector *modulateBurst(const BitVector &wBurst, int guardPeriodLength, int sps, bool emptyPulse = false);	     // :970 (sps = 4 only)
signalVector *modulateEdgeBurst(const BitVector &bits, int sps, bool emptyPulse = false);			     // :917 (sps = 4 only)
signalVector *generateEdgeBurst(int tsc);									     // :868
signalVector *generateEmptyBurst(int sps, int tn);								     // :843
signalVector *genRandNormalBurst(int tsc, int sps, int tn);							     // :768
signalVector *genRandAccessBurst(int delay, int sps, int tn);							     // :811
signalVector *generateDummyBurst(int sps, int tn);								     // :856
void scaleVector(signalVector &x, complex scale);								     // :1188
signalVector *delayVector(const signalVector *in, signalVector *out, float delay);				     // :1046
float energyDetect(const signalVector &rxBurst, unsigned windowLength);						     // :1573
int detectAnyBurst(const signalVector &burst, unsigned tsc, float threshold, int sps, CorrType type, unsigned max_toa,
		   struct estim_burst_params *ebp);								     // :1926
int detectSCHBurst(signalVector &rxBurst, float detectThreshold, int sps, sch_detect_type state,
		   struct estim_burst_params *ebp);								     // :1805 (SCH_DETECT_FULL; other states return -1)
SoftVector *demodAnyBurst(const signalVector &burst, CorrType type, int sps, struct estim_burst_params *ebp);	     // :2130

// ---- batched helpers (this repository's extension): N bursts per call, host vectors in and out ----
struct BurstResult {
	int rc;
	estim_burst_params ebp;
	std::vector<float> soft; // 148 (GMSK) or 444 (EDGE) values when rc > 0
};
// detectAnyBurst + demodAnyBurst for bursts[k] (>= 625 samples each) with per-burst type / tsc / max_toa
std::vector<BurstResult> detectDemodBursts(const std::vector<const signalVector *> &bursts, const std::vector<CorrType> &type,
					   const std::vector<unsigned> &tsc, const std::vector<unsigned> &max_toa, float threshold);

// the C ABI context behind this layer (NULL before sigProcLibSetup()); for callers that mix both levels
struct trxb200_ctx;
trxb200_ctx *sigProcLibContext();
