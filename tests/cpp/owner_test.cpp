// The multiplicative form of the schedule's vertex deal (gms_b200/csrc/owner.cuh) against the plain modulo form:
// exhaustively on small positions, on the whole 31-bit range in large strides and around every multiple of 2P near
// the top of the range, for every part count the fast form serves.
#include <cstdio>
#include <cstdlib>

#include "../../gms_b200/csrc/owner.cuh"

int main() {
    using namespace gmsb;
    long checked = 0;
    for (int P = 1; P <= 70; ++P) {
        const OwnerDeal o = make_owner_deal(P);
        if ((P > 1 && P <= 64) != (o.magic != 0)) { std::printf("fast form not selected as documented for P=%d\n", P); return 1; }
        auto check = [&](uint32_t d) {
            const int want = snake_owner((int)d, P), got = deal_owner(o, d);
            if (want != got) { std::printf("P=%d d=%u: %d != %d\n", P, d, got, want); std::exit(1); }
            ++checked;
        };
        for (uint32_t d = 0; d < 200000; ++d) check(d);
        for (uint64_t d = 0; d < (1ull << 31); d += 9973) check((uint32_t)d);
        const uint32_t top = 0x7fffffffu;
        for (uint32_t k = 0; k < 4096; ++k) check(top - k);
        for (uint32_t base = top - 100000u * (uint32_t)(2 * P); base < top - (uint32_t)(4 * P); base += 7919u * (uint32_t)(2 * P)) {
            const uint32_t m = base / (uint32_t)(2 * P) * (uint32_t)(2 * P);
            for (int k = -2; k <= 2; ++k) check(m + (uint32_t)k);
        }
        // every device gets the same number of positions out of any 2P consecutive ones
        int cnt[70] = {0};
        for (uint32_t d = 1000; d < 1000u + (uint32_t)(2 * P) * 50u; ++d) cnt[deal_owner(o, d)]++;
        for (int i = 0; i < P; ++i) if (cnt[i] != 100) { std::printf("P=%d: device %d got %d of 100\n", P, i, cnt[i]); return 1; }
    }
    std::printf("all checks passed (%ld positions)\n", checked);
    return 0;
}
