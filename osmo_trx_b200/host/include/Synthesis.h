// Synthesis.h — Transceiver52M/Synthesis.h:13-32: M-channel polyphase synthesis filterbank.
#pragma once
#include "ChannelizerBase.h"

class Synthesis : public ChannelizerBase {
public:
	Synthesis(size_t m, size_t blockLen, size_t hLen = 16);
	~Synthesis();
	size_t inputLen() const;  // blockLen
	size_t outputLen() const; // blockLen * m
	bool rotate(float *out, size_t oLen);
	float *inputBuffer(size_t chan) const;
	bool resetBuffer(size_t chan);
};
