// ChannelizerBase.h — common part of the analysis / synthesis filterbanks (Transceiver52M/ChannelizerBase.h):
// parameters, the per-channel host buffers callers read / fill, and the GPU object holding taps and history.
#pragma once
#include <cstddef>
#include <vector>
struct trxb200_filterbank;

class ChannelizerBase {
protected:
	ChannelizerBase(size_t m, size_t blockLen, size_t hLen, bool synthesis);
	~ChannelizerBase();
	size_t m, hLen, blockLen;
	bool synthesis;
	trxb200_filterbank *fb = nullptr;
	std::vector<std::vector<float>> chanBuf; // [m][2 * blockLen] host side of outputBuffer() / inputBuffer()
	bool checkLen(size_t innerLen, size_t outerLen);

public:
	bool init(); // ChannelizerBase.cpp:182-207
};
