// 7-card Hold'em strength (device + host inline): shared by hand_eval.cu and env_kernels.cu.
// Encoding and quirks: see oracle/hand_eval_oracle.c (pinned bit-for-bit to the reference's lib_hand_eval.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prl_he {

constexpr int kPairBase = 576011, kTwoPairBase = 658508, kTripsBase = 661446, kStraightBase = 664384,
              kFlushBase = 664398, kFullHouseBase = 1240409, kQuadsBase = 1240618, kStraightFlushBase = 1240827;

struct CardSet {
    unsigned long long cnt;  // 4 bits per rank
    unsigned suit[4];        // rank mask per suit
    __host__ __device__ __forceinline__ void add(int c) {
        const int r = c >> 2, s = c & 3;
        cnt += 1ull << (4 * r);
        suit[s] |= 1u << r;
    }
};

__host__ __device__ __forceinline__ int straight_top(unsigned mask) {
    // a run of five set bits ending at `top`; the wheel (A,2,3,4,5) has top = 3
    unsigned m = mask & (mask >> 1) & (mask >> 2) & (mask >> 3) & (mask >> 4);  // bit i set: ranks i..i+4 present
    if (m) {
        int top = 0;
        for (int i = 8; i >= 0; --i)
            if (m & (1u << i)) { top = i + 4; break; }
        return top;
    }
    return ((mask & 0x100Fu) == 0x100Fu) ? 3 : -1;
}

__host__ __device__ __forceinline__ int top5_value(unsigned mask) {
    int v = 0, n = 0;
    for (int r = 12; r >= 0 && n < 5; --r)
        if (mask & (1u << r)) { v = v * 13 + r; ++n; }
    return v;
}

__host__ __device__ __forceinline__ int popc13(unsigned m) {
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

// strength of the best 5-card hand in a 7-card set
__host__ __device__ inline int rank_cardset(const CardSet& cs) {
    const unsigned all = cs.suit[0] | cs.suit[1] | cs.suit[2] | cs.suit[3];
    int flush_suit = -1;
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (popc13(cs.suit[s]) >= 5) flush_suit = s;
    if (flush_suit >= 0) {
        const int st = straight_top(cs.suit[flush_suit]);
        if (st >= 0) return kStraightFlushBase + st;
    }
    int quad = -1, trip1 = -1, trip2 = -1, pair1 = -1, pair2 = -1;
    for (int r = 12; r >= 0; --r) {
        const int n = (int)((cs.cnt >> (4 * r)) & 0xF);
        if (n == 4) quad = r;
        else if (n == 3) { if (trip1 < 0) trip1 = r; else if (trip2 < 0) trip2 = r; }
        else if (n == 2) { if (pair1 < 0) pair1 = r; else if (pair2 < 0) pair2 = r; }
    }
    if (quad >= 0) {
        // quirk of the reference binary: the kicker is the card right above the quads in descending order if there is
        // one, otherwise the best card below
        int k = -1;
        for (int r = quad + 1; r <= 12 && k < 0; ++r)
            if (all & (1u << r)) k = r;
        for (int r = quad - 1; r >= 0 && k < 0; --r)
            if (all & (1u << r)) k = r;
        return kQuadsBase + 13 * quad + k;
    }
    if (trip1 >= 0 && (trip2 >= 0 || pair1 >= 0)) return kFullHouseBase + 13 * trip1 + (trip2 > pair1 ? trip2 : pair1);
    if (flush_suit >= 0) return kFlushBase + top5_value(cs.suit[flush_suit]);
    {
        const int st = straight_top(all);
        if (st >= 0) return kStraightBase + st;
    }
    if (trip1 >= 0) {
        const unsigned rest = all & ~(1u << trip1);
        int k0 = -1, k1 = -1;
        for (int r = 12; r >= 0; --r)
            if (rest & (1u << r)) { if (k0 < 0) k0 = r; else if (k1 < 0) k1 = r; }
        return kTripsBase + 169 * trip1 + 13 * k0 + k1;
    }
    if (pair2 >= 0) {
        const unsigned rest = all & ~((1u << pair1) | (1u << pair2));
        int k = -1;
        for (int r = 12; r >= 0 && k < 0; --r)
            if (rest & (1u << r)) k = r;
        return kTwoPairBase + 169 * pair1 + 13 * pair2 + k;
    }
    if (pair1 >= 0) {
        const unsigned rest = all & ~(1u << pair1);
        int k[3] = {0, 0, 0}, n = 0;
        for (int r = 12; r >= 0 && n < 3; --r)
            if (rest & (1u << r)) k[n++] = r;
        return kPairBase + 2197 * pair1 + 169 * k[0] + 13 * k[1] + k[2];
    }
    return top5_value(all);
}

}  // namespace prl_he
