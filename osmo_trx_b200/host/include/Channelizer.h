// Channelizer.h — Transceiver52M/Channelizer.h:13-31: M-channel polyphase analysis filterbank.
#pragma once
#include "ChannelizerBase.h"

class Channelizer : public ChannelizerBase {
public:
	Channelizer(size_t m, size_t blockLen, size_t hLen = 16);
	~Channelizer();
	size_t inputLen() const;  // blockLen * m complex samples
	size_t outputLen() const; // blockLen
	bool rotate(const float *in, size_t iLen);
	float *outputBuffer(size_t chan) const; // valid until the next rotate()
};
