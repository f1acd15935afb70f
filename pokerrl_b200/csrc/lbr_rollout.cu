// Local Best Response roll-outs on the GPU: the check-down equity of the LBR hand against the agent's
// range over EVERY completion of the board - PokerRL/eval/lbr/LocalLBRWorker.py:377-512 (_LBRRolloutManager), the only
// in-tree consumer of the batched hand evaluator (LocalLBRWorker.py:420).  One block per (query, board completion):
// 1326 seven-card evaluations (hand_eval.cuh, bit-identical to lib_hand_eval.so), range-weighted win / tie mass, weighted
// by the probability of the completion under the agent's card-removal distribution, exactly as the reference computes it:
//   card_probs = normalise(1 - P(agent holds c)), 0 for the LBR cards and the dealt cards      (:429-447)
//   completions in ascending card order, probability of each card drawn without replacement ~ card_probs, times n! (:449-466)
//   per completion: range with the new board cards zeroed and renormalised (uniform if nothing is left, PokerRange.py:
//   45-50), equity = mass of hands ranked below the LBR hand + half the mass of ties                (:496-508)
// Reference quirk (reproduced only on request, `first_board_ranks`): _calc_eq never advances its board counter `_i`
// (LocalLBRWorker.py:468-512), so the reference compares ranks on the FIRST completion for every completion while zeroing /
// weighting with the actual one.  The default evaluates every completion on its own ranks; the quirk mode exists so that
// the kernel can be checked bit-for-bit-in-meaning against fixtures produced by running the reference.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hand_eval.cuh"
#include "pokerrl_b200.h"
#include "prl_common.cuh"

namespace {

using namespace prl_he;
constexpr int kR = 1326, kDeckN = 52, kThreadsL = 256;

__device__ __forceinline__ void hole_cards_of_idx(int idx, int& c1, int& c2) {
    int a = (int)((103.0f - sqrtf(103.0f * 103.0f - 8.0f * (float)idx)) * 0.5f);
    while (a * (103 - a) / 2 > idx) --a;
    while ((a + 1) * (102 - a) / 2 <= idx) ++a;
    c1 = a;
    c2 = idx - a * (103 - a) / 2 + a + 1;
}

__device__ __forceinline__ double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return floor(r + 0.5);
}

// cp[q][c]: normalised probability that card c is NOT in the agent's hand, 0 for LBR / dealt cards (LocalLBRWorker.py:429-447)
__global__ void lbr_card_probs_kernel(const int8_t* __restrict__ hands, const int8_t* __restrict__ boards, int n_dealt,
                                      const float* __restrict__ ranges, double* __restrict__ cp) {
    __shared__ double s[kDeckN];
    const int q = blockIdx.x, c = threadIdx.x;
    const float* r = ranges + (size_t)q * kR;
    double v = 0.0;
    if (c < kDeckN) {
        float acp = 0.0f;  // PokerRange.get_card_probs (:28-36): float32 sums over the hands holding c
        for (int x = 0; x < kDeckN; ++x) {
            if (x == c) continue;
            const int c1 = min(c, x), c2 = max(c, x);
            acp += r[c1 * (103 - c1) / 2 + c2 - c1 - 1];
        }
        v = (double)(1.0f - acp);
        if (c == hands[2 * q] || c == hands[2 * q + 1]) v = 0.0;
        for (int k = 0; k < n_dealt; ++k)
            if (c == boards[5 * q + k]) v = 0.0;
        s[c] = v;
    }
    __syncthreads();
    if (c < kDeckN) {
        double tot = 0.0;
        for (int x = 0; x < kDeckN; ++x) tot += s[x];
        cp[(size_t)q * kDeckN + c] = (tot > 0.0) ? v / tot : v;
    }
}

// block (t, q): completion t of query q
__global__ void __launch_bounds__(kThreadsL) lbr_rollout_kernel(const int8_t* __restrict__ hands, const int8_t* __restrict__ boards,
                                                               int n_dealt, const float* __restrict__ ranges,
                                                               const double* __restrict__ cp, int n_comp, int first_board_ranks,
                                                               double* __restrict__ partial) {
    __shared__ double red[3][kThreadsL / 32];
    __shared__ int s_new[5];
    __shared__ int s_first[5];
    __shared__ double s_reach;
    const int t = blockIdx.x, q = blockIdx.y, n_new = 5 - n_dealt;
    const int l1 = hands[2 * q], l2 = hands[2 * q + 1];
    unsigned long long used = (1ull << l1) | (1ull << l2);
    for (int k = 0; k < n_dealt; ++k) used |= 1ull << boards[5 * q + k];
    if (threadIdx.x == 0) {
        // unrank completion t among the ascending n_new-combinations of the possible cards; its probability: cards drawn
        // one after the other without replacement, proportionally to cp (LocalLBRWorker.py:476-494)
        int poss[kDeckN], np = 0;
        for (int c = 0; c < kDeckN; ++c)
            if (!((used >> c) & 1ull)) poss[np++] = c;
        for (int j = 0; j < n_new; ++j) s_first[j] = poss[j];  // completion 0 = the n_new smallest possible cards
        double reach = 1.0, left = 1.0;
        long long rest = t;
        int start = 0;
        const double* cpq = cp + (size_t)q * kDeckN;
        for (int j = 0; j < n_new; ++j) {
            int x = start;
            for (;; ++x) {
                const long long cnt = (long long)binom(np - x - 1, n_new - j - 1);
                if (rest < cnt) break;
                rest -= cnt;
            }
            const int c = poss[x];
            s_new[j] = c;
            const double pc = cpq[c];
            reach *= (left > 0.0) ? pc / left : 0.0;
            left -= pc;
            start = x + 1;
        }
        s_reach = reach;
    }
    __syncthreads();
    CardSet base = {0ull, {0u, 0u, 0u, 0u}};
    unsigned long long bmask = 0;
    for (int k = 0; k < n_dealt; ++k) {
        base.add(boards[5 * q + k]);
        bmask |= 1ull << boards[5 * q + k];
    }
    for (int j = 0; j < n_new; ++j) {
        base.add(s_new[j]);
        bmask |= 1ull << s_new[j];
    }
    // boards the ranks are taken on: the completion itself, or (reference quirk) the first completion
    CardSet rbase = base;
    unsigned long long rmask = bmask;
    if (first_board_ranks && n_new > 0) {
        rbase = CardSet{0ull, {0u, 0u, 0u, 0u}};
        rmask = 0;
        for (int k = 0; k < n_dealt; ++k) {
            rbase.add(boards[5 * q + k]);
            rmask |= 1ull << boards[5 * q + k];
        }
        for (int j = 0; j < n_new; ++j) {
            rbase.add(s_first[j]);
            rmask |= 1ull << s_first[j];
        }
    }
    CardSet mine = rbase;
    mine.add(l1);
    mine.add(l2);
    const int my_rank = rank_cardset(mine);
    const float* r = ranges + (size_t)q * kR;
    double lt = 0.0, eq = 0.0, z = 0.0;
    int n_lt = 0, n_eq = 0;
    for (int idx = threadIdx.x; idx < kR; idx += kThreadsL) {
        int c1, c2;
        hole_cards_of_idx(idx, c1, c2);
        int v = -1;
        const bool live = !(((bmask >> c1) | (bmask >> c2)) & 1ull);
        if (!(((rmask >> c1) | (rmask >> c2)) & 1ull)) {
            CardSet cs = rbase;
            cs.add(c1);
            cs.add(c2);
            v = rank_cardset(cs);
        }
        const double w = live ? (double)r[idx] : 0.0;  // set_cards_to_zero_prob (PokerRange.py:66-82)
        z += w;
        if (v < my_rank) { lt += w; ++n_lt; }
        else if (v == my_rank) { eq += w; ++n_eq; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double cnt_lt = (double)n_lt, cnt_eq = (double)n_eq;
    for (int o = 16; o > 0; o >>= 1) {
        lt += __shfl_xor_sync(0xffffffffu, lt, o);
        eq += __shfl_xor_sync(0xffffffffu, eq, o);
        z += __shfl_xor_sync(0xffffffffu, z, o);
        cnt_lt += __shfl_xor_sync(0xffffffffu, cnt_lt, o);
        cnt_eq += __shfl_xor_sync(0xffffffffu, cnt_eq, o);
    }
    __shared__ double cnt[2][kThreadsL / 32];
    if (lane == 0) {
        red[0][warp] = lt; red[1][warp] = eq; red[2][warp] = z;
        cnt[0][warp] = cnt_lt; cnt[1][warp] = cnt_eq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0, na = 0.0, nb = 0.0;
        for (int w = 0; w < kThreadsL / 32; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; na += cnt[0][w]; nb += cnt[1][w]; }
        // normalise(): a range with no mass left is reset to uniform over all 1326 hands (PokerRange.py:45-50)
        const double equity = (c > 0.0) ? (a + 0.5 * b) / c : (na + 0.5 * nb) / (double)kR;
        partial[(size_t)q * n_comp + t] = equity * s_reach;
    }
}

__global__ void lbr_sum_kernel(const double* __restrict__ partial, int n_comp, double factorial, float* __restrict__ out) {
    __shared__ double red[kThreadsL];
    const int q = blockIdx.x;
    double s = 0.0;
    for (int t = threadIdx.x; t < n_comp; t += kThreadsL) s += partial[(size_t)q * n_comp + t];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kThreadsL / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = (float)(red[0] * factorial);  // one permutation per board was walked (:464-466)
}

}  // namespace

extern "C" long long prl_lbr_workspace_doubles(int n_queries, int n_dealt) {
    if (n_dealt < 0 || n_dealt > 5) return -1;
    const int np = 52 - n_dealt - 2, n_new = 5 - n_dealt;
    double c = 1.0;
    for (int i = 1; i <= n_new; ++i) c = c * (double)(np - n_new + i) / (double)i;
    return (long long)n_queries * (52 + (long long)(c + 0.5));
}

extern "C" int prl_lbr_checkdown_equity(const int8_t* lbr_hands, const int8_t* boards, int n_dealt, const float* ranges, int n_queries,
                                        int first_board_ranks, double* workspace, float* out, prl_stream_t stream) {
    if (n_queries <= 0) return 0;
    if (n_dealt < 0 || n_dealt > 5 || !lbr_hands || !boards || !ranges || !workspace || !out)
        return prl::fail("prl_lbr_checkdown_equity: bad arguments");
    const int np = 52 - n_dealt - 2, n_new = 5 - n_dealt;
    double c = 1.0, fact = 1.0;
    for (int i = 1; i <= n_new; ++i) {
        c = c * (double)(np - n_new + i) / (double)i;
        fact *= (double)i;
    }
    const long long n_comp = (long long)(c + 0.5);
    if (n_comp > 2147483647LL || n_queries > 65535) return prl::fail("prl_lbr_checkdown_equity: too many completions / queries per call");
    cudaStream_t s = (cudaStream_t)stream;
    double* cp = workspace;
    double* partial = workspace + (size_t)n_queries * 52;
    lbr_card_probs_kernel<<<n_queries, 64, 0, s>>>(lbr_hands, boards, n_dealt, ranges, cp);
    lbr_rollout_kernel<<<dim3((unsigned)n_comp, (unsigned)n_queries), kThreadsL, 0, s>>>(lbr_hands, boards, n_dealt, ranges, cp, (int)n_comp, first_board_ranks, partial);
    lbr_sum_kernel<<<n_queries, kThreadsL, 0, s>>>(partial, (int)n_comp, fact, out);
    prl::count_launch();
    prl::count_launch();
    prl::count_launch();
    return prl::check(cudaGetLastError(), "prl_lbr_checkdown_equity");
}
