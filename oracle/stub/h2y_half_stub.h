/*
 * Minimal stand-in for OpenEXR's `half` (ImathHalf), enough to COMPILE the reference's dpx.cpp for the oracle
 * (OpenEXR is not in this image).  TEST INFRASTRUCTURE ONLY.  IEEE 754 binary16 <-> binary32 with round-to-nearest-even
 * on the way down, which is what half.h's conversion does for every finite value, infinity and NaN.
 */
#ifndef H2Y_HALF_STUB_H
#define H2Y_HALF_STUB_H
#include <stdint.h>
#include <string.h>

namespace Imath {}
namespace Imf {}

class half {
public:
    half() : bits_(0) {}
    half(float f) { bits_ = from_float(f); }
    operator float() const { return to_float(bits_); }
    half &operator=(float f) { bits_ = from_float(f); return *this; }
    uint16_t bits() const { return bits_; }

private:
    uint16_t bits_;
    static float to_float(uint16_t h)
    {
        const uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u;
        uint32_t u;
        if (e == 0) {
            if (m == 0) u = s;
            else {
                int sh = 0;
                uint32_t mm = m;
                while (!(mm & 1024u)) { mm <<= 1; sh++; }
                u = s | ((uint32_t)(113 - sh) << 23) | ((mm & 1023u) << 13);
            }
        } else if (e == 31) u = s | 0x7f800000u | (m << 13);
        else u = s | ((e + 112u) << 23) | (m << 13);
        float f;
        memcpy(&f, &u, 4);
        return f;
    }
    static uint16_t from_float(float f)
    {
        uint32_t u;
        memcpy(&u, &f, 4);
        const uint32_t s = (u >> 16) & 0x8000u;
        const int32_t e = (int32_t)((u >> 23) & 255u) - 127 + 15;
        uint32_t m = u & 0x7fffffu;
        if (((u >> 23) & 255u) == 255u) return (uint16_t)(s | 0x7c00u | (m ? (0x200u | (m >> 13)) : 0u));
        if (e >= 31) return (uint16_t)(s | 0x7c00u);
        if (e <= 0) {
            if (e < -10) return (uint16_t)s;
            m |= 0x800000u;
            const int sh = 14 - e;
            uint32_t h = m >> sh;
            const uint32_t rem = m & ((1u << sh) - 1u), halfway = 1u << (sh - 1);
            if (rem > halfway || (rem == halfway && (h & 1u))) h++;
            return (uint16_t)(s | h);
        }
        uint32_t h = ((uint32_t)e << 10) | (m >> 13);
        const uint32_t rem = m & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
        return (uint16_t)(s | h);
    }
};
#endif
