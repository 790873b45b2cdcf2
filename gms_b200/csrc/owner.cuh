// owner.cuh — which device owns a closing vertex of the triangle schedule (tc.cu) when part_count > 1.
// Plain arithmetic, compiled for the device by nvcc and for the host by the CPU test (tests/cpp/owner_test.cpp).
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define GMSB_HD
#else
#define GMSB_HD __host__ __device__ __forceinline__
#endif

namespace gmsb {

// owner of position d of a deal among P devices: 0 1 .. P-1 P-1 .. 1 0 0 1 ..
GMSB_HD int snake_owner(int d, int P) {
    const int r = d % (2 * P);
    return r < P ? r : 2 * P - 1 - r;
}

// The same deal with the division by 2P replaced by a multiplication (the schedule passes evaluate it once per edge):
// with D = 2P <= 2^l and s = 31 + l, magic = floor(2^s / D) + 1 gives floor(d / D) = (d * magic) >> s exactly for every
// d < 2^31 (the error term d * (D - 2^s mod D) stays below 2^31 * 2^l = 2^s; d * magic < 2^63).  parts > 64: the plain form.
struct OwnerDeal {
    uint64_t magic = 0;
    int shift = 0;
    int parts = 1;
};
inline OwnerDeal make_owner_deal(int parts) {
    OwnerDeal o;
    o.parts = parts;
    if (parts > 1 && parts <= 64) {
        int l = 0;
        while ((1 << l) < 2 * parts) ++l;
        o.shift = 31 + l;
        o.magic = (1ull << o.shift) / (uint64_t)(2 * parts) + 1ull;
    }
    return o;
}
// Vertices are dealt from the top of rank space (d = n - 1 - v): that is where the hubs are, heaviest (nearly) first.
GMSB_HD int deal_owner(const OwnerDeal &o, uint32_t d) {
    if (o.magic == 0) return snake_owner((int)d, o.parts);
    const uint32_t D = 2u * (uint32_t)o.parts;
    const uint32_t q = (uint32_t)(((uint64_t)d * o.magic) >> o.shift);
    const uint32_t r = d - q * D;
    return (int)(r < (uint32_t)o.parts ? r : D - 1u - r);
}

}  // namespace gmsb
