// Complex.h — host-side mirror of the boundary type Complex<T> (reference: Transceiver52M/Complex.h:29-141).
// Only what callers of the burst-DSP API touch: interleaved {re, im} storage (so a signalVector can be
// handed to the C ABI as float*), construction, arithmetic, conj/norm2/abs/arg.  Written for this
// repository; the arithmetic of the hot path itself runs on the GPU, not here.
#pragma once
#include <cmath>
#include <ostream>

template <class Real> struct Complex {
	Real r, i;

	Complex() : r(0), i(0) {}
	Complex(Real re) : r(re), i(0) {}
	Complex(Real re, Real im) : r(re), i(im) {}
	template <class Other> Complex(const Complex<Other> &o) : r((Real)o.r), i((Real)o.i) {}

	Real real() const { return r; }
	Real imag() const { return i; }
	void real(Real v) { r = v; }
	void imag(Real v) { i = v; }

	Complex conj() const { return Complex(r, -i); }
	Real norm2() const { return i * i + r * r; }
	Real abs() const { return (Real)std::sqrt((double)norm2()); }
	Real arg() const { return (Real)std::atan2((double)i, (double)r); }
	Complex inv() const { const Real n = norm2(); return Complex(r / n, -i / n); }

	Complex operator+(const Complex &o) const { return Complex(r + o.r, i + o.i); }
	Complex operator-(const Complex &o) const { return Complex(r - o.r, i - o.i); }
	Complex operator-() const { return Complex(-r, -i); }
	Complex operator*(const Complex &o) const { return Complex(r * o.r - i * o.i, r * o.i + i * o.r); }
	Complex operator*(Real s) const { return Complex(r * s, i * s); }
	Complex operator/(const Complex &o) const { return *this * o.inv(); }
	Complex operator/(Real s) const { return Complex(r / s, i / s); }
	Complex &operator+=(const Complex &o) { r += o.r; i += o.i; return *this; }
	Complex &operator-=(const Complex &o) { r -= o.r; i -= o.i; return *this; }
	Complex &operator*=(Real s) { r *= s; i *= s; return *this; }
	bool operator==(const Complex &o) const { return r == o.r && i == o.i; }
	bool operator!=(const Complex &o) const { return !(*this == o); }
	// ordering by power, as the reference's peak searches use it (Complex.h:99-100)
	bool operator<(const Complex &o) const { return norm2() < o.norm2(); }
	bool operator>(const Complex &o) const { return norm2() > o.norm2(); }
};

template <class Real> std::ostream &operator<<(std::ostream &os, const Complex<Real> &z)
{
	return os << z.r << (z.i < 0 ? "-" : "+") << std::fabs(z.i) << "j";
}

typedef Complex<float> complex;
