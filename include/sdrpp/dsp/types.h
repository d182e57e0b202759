// Host-side mirror of the reference's sample types (core/src/dsp/types.h:6-127).
// Layout-compatible PODs: complex_t is interleaved {re, im} fp32 = sdrpp_cf32 of the C ABI.
#pragma once
#include <cmath>

namespace dsp {
    struct complex_t {
        float re;
        float im;

        complex_t operator*(float b) const { return complex_t{ re * b, im * b }; }
        complex_t operator*(double b) const { return complex_t{ (float)(re * b), (float)(im * b) }; }
        complex_t operator/(float b) const { return complex_t{ re / b, im / b }; }
        // same operation order as the reference (types.h:23-25): results of host-side arithmetic agree bit for bit
        complex_t operator*(const complex_t& b) const { return complex_t{ (re * b.re) - (im * b.im), (im * b.re) + (re * b.im) }; }
        complex_t operator+(const complex_t& b) const { return complex_t{ re + b.re, im + b.im }; }
        complex_t operator-(const complex_t& b) const { return complex_t{ re - b.re, im - b.im }; }
        complex_t& operator+=(const complex_t& b) { re += b.re; im += b.im; return *this; }
        complex_t& operator-=(const complex_t& b) { re -= b.re; im -= b.im; return *this; }
        complex_t& operator*=(float b) { re *= b; im *= b; return *this; }
        complex_t conj() const { return complex_t{ re, -im }; }
        float phase() const { return atan2f(im, re); }
        float amplitude() const { return sqrtf(re * re + im * im); }
        float fastAmplitude() const {
            const float a = fabsf(re), b = fabsf(im);
            return a > b ? a + 0.4f * b : b + 0.4f * a;
        }
    };

    struct stereo_t {
        float l;
        float r;
        stereo_t operator*(float b) const { return stereo_t{ l * b, r * b }; }
        stereo_t operator+(const stereo_t& b) const { return stereo_t{ l + b.l, r + b.r }; }
        stereo_t operator-(const stereo_t& b) const { return stereo_t{ l - b.l, r - b.r }; }
        stereo_t& operator+=(const stereo_t& b) { l += b.l; r += b.r; return *this; }
    };
}
