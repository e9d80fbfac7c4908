// Vector.h — host-side mirror of CommonLibs/Vector.h:56-318: a fixed-size vector that either owns its
// storage or aliases a segment of another vector.  Subset used by callers of the burst-DSP API.
#pragma once
#include <cassert>
#include <new>
#include <cstddef>
#include <cstring>
#include <ostream>

typedef void *(*vector_alloc_func)(size_t);
typedef void (*vector_free_func)(void *);

template <class T> class Vector {
public:
	Vector(size_t n = 0, vector_alloc_func af = nullptr, vector_free_func ff = nullptr) : mAlloc(af), mFree(ff) { allocate(n); }
	// alias [start, end) of storage owned elsewhere
	Vector(T *data, T *start, T *end) : mData(data), mStart(start), mEnd(end), mOwned(false) {}
	Vector(const Vector &o) : mAlloc(o.mAlloc), mFree(o.mFree) { allocate(o.size()); copy_from(o); }
	Vector(Vector &&o) noexcept { steal(o); }
	~Vector() { release(); }

	Vector &operator=(const Vector &o)
	{
		if (this != &o) { resize(o.size()); copy_from(o); }
		return *this;
	}
	Vector &operator=(Vector &&o) noexcept
	{
		if (this != &o) { release(); steal(o); }
		return *this;
	}

	void resize(size_t n) { release(); allocate(n); }
	void clear() { release(); }
	size_t size() const { return (size_t)(mEnd - mStart); }
	size_t bytes() const { return size() * sizeof(T); }
	T *begin() { return mStart; }
	const T *begin() const { return mStart; }
	T *end() { return mEnd; }
	const T *end() const { return mEnd; }
	T &operator[](size_t k) { assert(mStart + k < mEnd); return mStart[k]; }
	const T &operator[](size_t k) const { assert(mStart + k < mEnd); return mStart[k]; }

	// non-owning views (Vector.h:203-218)
	Vector segment(size_t start, size_t span) { assert(start + span <= size()); return Vector(nullptr, mStart + start, mStart + start + span); }
	const Vector segment(size_t start, size_t span) const { return const_cast<Vector *>(this)->segment(start, span); }
	Vector head(size_t span) { return segment(0, span); }
	Vector tail(size_t start) { return segment(start, size() - start); }

	void fill(const T &v) { for (T *p = mStart; p < mEnd; p++) *p = v; }
	void fill(const T &v, size_t start, size_t span) { for (size_t k = 0; k < span; k++) (*this)[start + k] = v; }
	void copyTo(Vector &dst) const { assert(dst.size() >= size()); for (size_t k = 0; k < size(); k++) dst.mStart[k] = mStart[k]; }
	void copyToSegment(Vector &dst, size_t start, size_t span) const { for (size_t k = 0; k < span; k++) dst[start + k] = mStart[k]; }

protected:
	T *mData = nullptr;  // owned allocation (nullptr for views)
	T *mStart = nullptr; // first visible element
	T *mEnd = nullptr;
	vector_alloc_func mAlloc = nullptr;
	vector_free_func mFree = nullptr;
	bool mOwned = true;

	void allocate(size_t n)
	{
		mOwned = true;
		if (!n) { mData = mStart = mEnd = nullptr; return; }
		mData = mAlloc ? static_cast<T *>(mAlloc(n * sizeof(T))) : new T[n]();
		if (mAlloc) for (size_t k = 0; k < n; k++) new (mData + k) T();
		mStart = mData;
		mEnd = mData + n;
	}
	void release()
	{
		if (mOwned && mData) { if (mFree) mFree(mData); else if (!mAlloc) delete[] mData; }
		mData = mStart = mEnd = nullptr;
	}
	void copy_from(const Vector &o) { for (size_t k = 0; k < o.size(); k++) mStart[k] = o.mStart[k]; }
	void steal(Vector &o)
	{
		mData = o.mData; mStart = o.mStart; mEnd = o.mEnd; mAlloc = o.mAlloc; mFree = o.mFree; mOwned = o.mOwned;
		o.mData = o.mStart = o.mEnd = nullptr;
	}
};

template <class T> std::ostream &operator<<(std::ostream &os, const Vector<T> &v)
{
	for (size_t k = 0; k < v.size(); k++) os << v[k] << " ";
	return os;
}
