// signalVector.h — host-side mirror of Transceiver52M/signalVector.h:13-53 / signalVector.cpp:3-107:
// a Vector<complex> with head-room in front of the visible samples (convolution history) and the
// real / aligned / symmetry flags the reference's convolve() dispatches on.
#pragma once
#include "Complex.h"
#include "Vector.h"

enum Symmetry { NONE = 0, ABSSYM = 1 };

class signalVector : public Vector<complex> {
public:
	signalVector(size_t n = 0, vector_alloc_func af = nullptr, vector_free_func ff = nullptr) : Vector<complex>(n, af, ff) {}
	// n visible samples behind `start` samples of (zeroed) head-room (signalVector.cpp:9-14)
	signalVector(size_t n, size_t start, vector_alloc_func af = nullptr, vector_free_func ff = nullptr) : Vector<complex>(n + start, af, ff)
	{
		mStart = mData + start;
	}
	// view of caller-owned samples (signalVector.cpp:16-21)
	signalVector(complex *data, size_t start, size_t span, vector_alloc_func = nullptr, vector_free_func = nullptr)
		: Vector<complex>(nullptr, data + start, data + start + span) {}
	signalVector(const signalVector &o) : Vector<complex>(o.size() + o.getStart()), mReal(o.mReal), mAligned(o.mAligned), mSym(o.mSym)
	{
		mStart = mData + o.getStart();
		for (size_t k = 0; k < o.size(); k++) mStart[k] = o.begin()[k];
	}
	// copy with extra head / tail room (signalVector.cpp:31-39)
	signalVector(const signalVector &o, size_t start, size_t tail = 0) : Vector<complex>(start + o.size() + tail), mReal(o.mReal), mAligned(o.mAligned), mSym(o.mSym)
	{
		mStart = mData + start;
		mEnd = mStart + o.size();
		for (size_t k = 0; k < o.size(); k++) mStart[k] = o.begin()[k];
	}
	signalVector &operator=(const signalVector &o)
	{
		if (this == &o) return *this;
		resize(o.size() + o.getStart());
		mStart = mData + o.getStart();
		for (size_t k = 0; k < o.size(); k++) mStart[k] = o.begin()[k];
		mReal = o.mReal; mAligned = o.mAligned; mSym = o.mSym;
		return *this;
	}
	signalVector segment(size_t start, size_t span) { return signalVector(mStart, start, span); }
	size_t getStart() const { return mData ? (size_t)(mStart - mData) : 0; }
	// keep the newest getStart() samples as history in the head-room (signalVector.cpp:65-77)
	size_t updateHistory()
	{
		const size_t h = getStart();
		const size_t n = h < size() ? h : size();
		for (size_t k = 0; k < n; k++) mData[h - n + k] = mEnd[-(long)n + (long)k];
		return n;
	}
	Symmetry getSymmetry() const { return mSym; }
	void setSymmetry(Symmetry s) { mSym = s; }
	bool isReal() const { return mReal; }
	void isReal(bool v) { mReal = v; }
	bool isAligned() const { return mAligned; }
	void setAligned(bool v) { mAligned = v; }

private:
	bool mReal = false, mAligned = false;
	Symmetry mSym = NONE;
};
