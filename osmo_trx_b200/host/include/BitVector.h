// BitVector.h — host-side mirror of CommonLibs/BitVector.h:35-156,171-246: one bit per char (only bit 0
// is meaningful, BitVector.cpp:59-68) and SoftVector = Vector<float> with values in [-1, 1], bit = (v > 0).
#pragma once
#include "Vector.h"
#include <cstdint>
#include <cstring>

class BitVector : public Vector<char> {
public:
	BitVector(size_t n = 0) : Vector<char>(n) {}
	// from a string of '0' / '1'
	BitVector(const char *bits) : Vector<char>(std::strlen(bits))
	{
		for (size_t k = 0; k < size(); k++) mStart[k] = (char)(bits[k] == '1');
	}
	BitVector(const Vector<char> &o) : Vector<char>(o) {}
	unsigned bit(size_t k) const { return (unsigned)(*this)[k] & 1u; }
	// MSB-first field extraction / insertion (BitVector.h peekField / fillField)
	uint64_t peekField(size_t pos, unsigned len) const
	{
		uint64_t v = 0;
		for (unsigned k = 0; k < len; k++) v = (v << 1) | bit(pos + k);
		return v;
	}
	void fillField(size_t pos, uint64_t value, unsigned len)
	{
		for (unsigned k = 0; k < len; k++) (*this)[pos + k] = (char)((value >> (len - 1 - k)) & 1);
	}
};

class SoftVector : public Vector<float> {
public:
	SoftVector(size_t n = 0) : Vector<float>(n) {}
	SoftVector(const Vector<float> &o) : Vector<float>(o) {}
	bool bit(size_t k) const { return (*this)[k] > 0.0f; } // BitVector.h:236-241
	BitVector sliced() const
	{
		BitVector b(size());
		for (size_t k = 0; k < size(); k++) b[k] = (char)bit(k);
		return b;
	}
};
