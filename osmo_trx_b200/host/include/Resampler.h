// Resampler.h — Transceiver52M/Resampler.h:31-61 on the GPU library: rational p/q polyphase resampler.
#pragma once
#include <cstddef>
struct trxb200_resampler;

class Resampler {
public:
	Resampler(size_t p, size_t q, size_t filt_len = 16);
	~Resampler();
	bool init(float bw = 1.0f); // builds the partitions (Resampler.cpp:47-96) and uploads them
	// in points at the first new sample; filt_len samples of history must precede it in memory
	// (Resampler.cpp:131-150 reads before `in`); returns out_len or a negative value on error
	int rotate(const float *in, size_t in_len, float *out, size_t out_len);
	size_t len();

private:
	size_t p, q, filt_len;
	trxb200_resampler *h = nullptr;
};
