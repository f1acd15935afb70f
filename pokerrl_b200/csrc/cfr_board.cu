// Board-resident CFR+ sweeps for two-hole-card games with ONE chance layer (Flop5Holdem, PokerRL/game/games.py:222-254) - sm_100a.
//
// The level-synchronous sweeps (cfr_twocard.cu) spill every node vector of every board subtree to HBM: 186 GB per
// iteration against 40 GB of regret / average tables.  Here ONE persistent CTA walks whole (board, seat)
// units: the 15-node post-deal subtree of a board lives in registers / shared memory, HBM sees only
//     opponent regret rows (strategy by regret matching)  ->  reach of the opponent, top-down        (P1)
//     9 terminal rows (5 showdown + 4 fold) evaluated together in shared memory                      (P2)
//     own regret + average rows read, updated, written; the board's root value accumulated into the
//     chance-node sum as 64-bit FIXED POINT (exactly associative: any grouping over CTAs / GPUs gives
//     the same bits)                                                                                 (P3)
// Rows of a board's table are stored in the board's STRENGTH ORDER and hold only the 1081 hands that do not collide
// with the board (stride 1088 floats instead of 1326 natural-order entries): showdown prefix sums run over consecutive
// addresses, blocked hands cost nothing, 18 % fewer bytes.  The per-board index tables (15 KB: packed per-hand record,
// hand ids, card rows) are staged by cp.async.bulk (TMA 1-D bulk copies) completing on mbarriers, the next board's tables
// in flight while the current board computes.
//
// Arithmetic follows the reference statements generalised to two-card hands (float64 restatement
// in oracle/cfr2_oracle.c): reach StrategyFiller.py:118-146, 159-166; fold / showdown values ValueFiller.py:103-158;
// value backup :64-93; regrets _CFRBase.py:146-185 + CFRPlus.py:37-41; regret matching CFRPlus.py:43-63; averaging
// CFRPlus.py:65-87.  The strategy is never stored: it is a pure function of the regret rows (CFRPlus.py:49-58).
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "pokerrl_b200.h"
#include "prl_common.cuh"

namespace {

// ---- geometry of a 52-card deck with a complete 5-card board
constexpr int kDeck = 52;
constexpr int kBoardCards = 5;
constexpr int kLiveCards = kDeck - kBoardCards;            // 47
constexpr int kLive = kLiveCards * (kLiveCards - 1) / 2;   // 1081 hands that hold no board card
constexpr int kLdb = 1088;                                  // row stride (floats) of the strength-ordered tables
constexpr int kRowLen = kLiveCards - 1;                     // 46 live hands hold a given live card
constexpr int kRowPad = 48;                                 // card rows padded to 4 lanes x 12 entries
constexpr int kRowSeg = 12;
constexpr int kErStride = kLiveCards;                       // centred prefix array of a card row: entries 0..46
constexpr int kZeroSlot = kLive + 1;                        // S[v][1082] is always 0 (target of the row padding)
constexpr int kRange = 1326;

// per-board table blob (bytes): packed records, hand ids, card rows
constexpr int kRecBytes = kLdb * 8;                         // uint64 per strength position
constexpr int kShBytes = kLdb * 2;                          // int16 hand id per strength position
constexpr int kRowIdxBytes = kLiveCards * kRowPad * 2;      // int16 strength position per card-row entry
constexpr int kBlobA = kRecBytes + kShBytes;                // 10 880 B, needed by P1 and P3
constexpr int kBlobBytes = kBlobA + kRowIdxBytes;           // 15 392 B
static_assert(kBlobA % 16 == 0 && kRowIdxBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

// kernel variants (A/B switches; the defaults are the measured winners, profiles/r02_q_sweep_variants.md)
#ifndef PRL_BV_RED
#define PRL_BV_RED 1         // chance sums by 64-bit RED instead of load + add + store
#endif
#ifndef PRL_BV_P1PIPE
#define PRL_BV_P1PIPE 1      // P1 rows: 1 = only the first position requested one unit ahead, the rest inside P1; 0 = all three
#endif
#ifndef PRL_BV_FOLDLIN
#define PRL_BV_FOLDLIN 1     // update form: card-row sums of the fold vectors from the showdown vectors' row totals (linearity)
#endif
#ifndef PRL_BV_ERT
#define PRL_BV_ERT 1         // card-row prefix arrays entry-major (Er[v][k][card]) instead of card-major: fewer store conflicts in P2a
#endif
#ifndef PRL_BV_ROWTOTF
#define PRL_BV_ROWTOTF 1     // kLin: a lane's part of a card row's total summed in float (12 terms), only the quad reduction in double
                             // (measured: +1 %, parity unchanged at <= 1.6e-7)
#endif
#ifndef PRL_BV_NEWTON
#define PRL_BV_NEWTON 1      // regret matching: Newton step after MUFU.RCP (<= 1 ulp); 0 = the approximation as is (2^-23 relative)
#endif
constexpr int kVP1Pipe = PRL_BV_P1PIPE;
constexpr bool kVRed = PRL_BV_RED, kVFoldLin = PRL_BV_FOLDLIN, kVErT = PRL_BV_ERT;
// measured and removed (profiles/r02_q_sweep_variants.md): five-warp / one-warp-per-vector scans, the single-warp stage spread
// over nine warps, the third P3 pass spread over all warps, two positions requested a unit ahead

constexpr int kThreads = 384;  // 12 warps; 3 strength positions per thread (3 * 384 = 1152 >= 1081: 94 % of the lanes busy)
constexpr int kPerThread = 3;
constexpr int kWarps = kThreads / 32;

// ---- compiled shape of the post-deal subtree (breadth-first; Flop5Holdem with pot-size raises, stacks that allow the
//      full raise sequence).  The host checks the game's abstract tree against these arrays.
struct ShapeFHP {
    static constexpr int N = 15;
    static constexpr int kind(int i) { constexpr int a[N] = {1, 0, 0, 4, 1, 3, 4, 1, 3, 4, 0, 3, 4, 3, 4}; return a[i]; }
    static constexpr int parent(int i) { constexpr int a[N] = {-1, 0, 0, 1, 1, 2, 2, 2, 4, 4, 4, 7, 7, 10, 10}; return a[i]; }
    static constexpr int first_child(int i) { constexpr int a[N] = {1, 3, 5, -1, 8, -1, -1, 11, -1, -1, 13, -1, -1, -1, -1}; return a[i]; }
    static constexpr int n_children(int i) { constexpr int a[N] = {2, 2, 3, 0, 3, 0, 0, 2, 0, 0, 2, 0, 0, 0, 0}; return a[i]; }
    // index of terminal i among the showdown / fold vectors
    static constexpr int vec_index(int i) {
        int n = 0;
        for (int k = 0; k < i; ++k) n += (kind(k) == kind(i));
        return n;
    }
    static constexpr int count(int k) {
        int n = 0;
        for (int i = 0; i < N; ++i) n += (kind(i) == k);
        return n;
    }
    static constexpr int n_sd = 5, n_fold = 4;
    // table rows of one board: the rows of seat 0's decision nodes first, then seat 1's, children in breadth-first order
    static constexpr int rows_of_seat(int p) {
        int n = 0;
        for (int i = 1; i < N; ++i) n += (kind(parent(i)) == p);
        return n;
    }
    static constexpr int rows = 14;
    static constexpr int row_of(int c) {  // c = child of a decision node
        const int p = kind(parent(c));
        int n = (p == 0) ? 0 : rows_of_seat(0);
        for (int k = 1; k < c; ++k) n += (kind(parent(k)) == p);
        return n;
    }
};
static_assert(ShapeFHP::count(4) == ShapeFHP::n_sd && ShapeFHP::count(3) == ShapeFHP::n_fold, "shape");
// Reach of the OPPONENT of seat P at fold terminal f as a combination of its reach at the showdown terminals (strategies sum
// to one, own nodes copy the reach): x_fold[f] = sum_v coef(P, f, v) * x_sd[v].  Every linear functional of the fold vectors
// (their card-row sums) follows from the showdown vectors' at no cost.  tools/fold_relations.py derives the table.
struct FoldLinFHP {
    static constexpr int coef(int P, int f, int v) {
        constexpr int c[2][4][5] = {{{0, 1, 0, 0, 0}, {1, 0, -1, 0, -1}, {0, 1, 0, -1, 0}, {0, 0, 0, 0, 1}},
                                    {{1, -1, 1, -1, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 1, 0, -1}}};
        return c[P][f][v];
    }
};
static_assert(ShapeFHP::rows_of_seat(0) + ShapeFHP::rows_of_seat(1) == ShapeFHP::rows, "shape");

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for_down(F&& f) {  // N-1 .. I
    if constexpr (I < N) {
        f(std::integral_constant<int, N - 1>{});
        static_for_down<I, N - 1>(f);
    }
}

// ---- shared memory carve-up (bytes)
constexpr int kNVec = ShapeFHP::n_sd + ShapeFHP::n_fold;                   // 9 terminal vectors
constexpr int kSOff = 0;                                                    // float S[9][1088]
constexpr int kErOff = kSOff + kNVec * kLdb * 4;                            // float Er[5][47*47]
constexpr int kErVec = kLiveCards * kErStride;                              // 2209 floats per showdown vector
constexpr int kErBytes = ((ShapeFHP::n_sd * kErVec * 4) + 15) & ~15;
constexpr int kBlobOff = kErOff + kErBytes;                                 // 2 x (rec + hand ids)
constexpr int kRowIdxOff = kBlobOff + 2 * kBlobA;                           // card rows (single buffer)
constexpr int kCsOff = kRowIdxOff + kRowIdxBytes;                           // float cs[4][48] per-card sums of the fold vectors
constexpr int kCsdOff = kCsOff + ShapeFHP::n_fold * kRowPad * 4;            // double csd[4][48]: the same sums before rounding
constexpr int kMiscOff = kCsdOff + ShapeFHP::n_fold * kRowPad * 8;          // double wsum[5][16], wexc[5][16]; float tf[8]
constexpr int kMiscBytes = (5 * 16 + 5 * 16) * 8 + 8 * 4;
constexpr int kRowTotOff = kMiscOff + kMiscBytes;                           // double rowtot[5][48]: card-row totals of the showdown vectors
constexpr int kRowTotBytes = ShapeFHP::n_sd * kRowPad * 8;
constexpr int kBarOff = kRowTotOff + kRowTotBytes;                          // 3 mbarriers
constexpr int kSmemBytes = kBarOff + 32;
static_assert(kBlobOff % 16 == 0 && kRowIdxOff % 16 == 0 && kCsdOff % 8 == 0 && kMiscOff % 8 == 0 && kRowTotOff % 8 == 0 && kBarOff % 8 == 0, "alignment");
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");

struct SweepArgs {
    prl_board_game_t g;
    const float* trunk_reach_opp;  // natural order row of the opponent's reach at the chance node
    int iter, delay;
    float m_old, m_new;            // CFRPlus.py:68-73
    int src_own, src_opp;          // evaluation: 0 = regret matching of `regret`, 1 = `avg` rows as they are (CFR+ average),
                                   // 2 = `avg` rows normalised (reach-weighted sums of Vanilla / Linear CFR, LinearCFR.py:64-71)
    float rw;                      // weight of the instantaneous regret: 1, Linear CFR iter + 1 (LinearCFR.py:27-28)
    float defer_w;                 // DEFER: weight of the opponent's pending average-strategy contribution (0: none)
    double fx_scale;               // 2^frac_bits
    float sc[16];                  // terminal n: K * pot / 2, negated where the seat of this sweep is the folder
};

// ---- PTX helpers: mbarrier + 1-D bulk copy global -> shared (TMA) + bulk L2 prefetch, sm_90+
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// strategy of one decision node from its table rows: regret matching (CFRPlus.py:43-63; the regrets of Vanilla / Linear
// CFR are clipped first, LinearCFR.py:33-51) or the rows as they are (CFR+ average strategy)
template <int A>
__device__ __forceinline__ void node_strategy(const float (&g)[A], int src, float (&s)[A]) {
    if (src == 1) {
#pragma unroll
        for (int a = 0; a < A; ++a) s[a] = g[a];
        return;
    }
    float sum = 0.0f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        s[a] = fmaxf(g[a], 0.0f);
        sum += s[a];
    }
    const bool pos = sum > 0.0f;
    // reciprocal = MUFU.RCP + one Newton step (<= 1 ulp, no range-check branch; sums below 1e-37 cannot occur: regrets are
    // chip amounts times probabilities)
    const float sm = fmaxf(sum, 1e-37f);
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sm));
    if (PRL_BV_NEWTON) inv = fmaf(inv, fmaf(-sm, inv, 1.0f), inv);
    inv = pos ? inv : 0.0f;
    const float uni = pos ? 0.0f : 1.0f / (float)A;  // CFRPlus.py:53-58: uniform where no regret is positive
#pragma unroll
    for (int a = 0; a < A; ++a) s[a] = fmaf(s[a], inv, uni);
}

__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// =====================================================================================================================
// The sweep kernel.  P = seat whose values are computed; EVAL = false: CFR+ update of seat P (regrets, average);
// EVAL = true: values and best-response values of seat P under the strategies selected by src_own / src_opp.
// Table rows of board j: [j][14][1088] floats, rows ShapeFHP::row_of(child); everything a unit touches is contiguous.
// =====================================================================================================================
// DEFER (Vanilla / Linear CFR, update form): regrets are not clipped, and the average is the reach-weighted SUM of
// strategies (VanillaCFR.py:54-60, LinearCFR.py:53-59) with the seat's reach under its NEW strategy - which includes its
// new trunk reach, known only after this seat's trunk update.  The contribution of seat q's update is therefore added
// during the NEXT sweep that walks q's rows anyway: P1 of the other seat's sweep computes exactly q's strategy and reach
// at every node (defer_w = its weight).  P1ONLY: nothing but that (flush before an evaluation of the average strategy).
template <class SH, int P, bool EVAL, bool DEFER = false, bool P1ONLY = false>
__global__ void __launch_bounds__(kThreads, 2) board_sweep_kernel(const SweepArgs a) {
    static_assert(!(EVAL && DEFER) && (!P1ONLY || DEFER), "variants");
    extern __shared__ __align__(128) unsigned char smem[];
    float* S = reinterpret_cast<float*>(smem + kSOff);
    float* Er = reinterpret_cast<float*>(smem + kErOff);
    float* cs = reinterpret_cast<float*>(smem + kCsOff);
    double* csd = reinterpret_cast<double*>(smem + kCsdOff);
    double* wsum = reinterpret_cast<double*>(smem + kMiscOff);  // [5][16] warp totals of the main scans
    double* wexc = wsum + 5 * 16;                                // [5][16] exclusive prefix of the warp totals - total / 2
    float* tf = reinterpret_cast<float*>(wexc + 5 * 16);         // [8]     totals of the fold vectors
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);  // [0], [1]: blob A buffers; [2]: card rows
    const int16_t* rowidx = reinterpret_cast<const int16_t*>(smem + kRowIdxOff);
    double* rowtot = reinterpret_cast<double*>(smem + kRowTotOff);  // [5][48]
    constexpr bool kLin = kVFoldLin && !EVAL;  // evaluation may read average-strategy rows, which sum to one only up to rounding

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const prl_board_game_t& G = a.g;
    const int nb = G.n_boards;
    constexpr int OPP = 1 - P;
    constexpr int NSD = SH::n_sd, NF = SH::n_fold;
    constexpr int ROWS = SH::rows;
    constexpr int OWN0 = (P == 0) ? 0 : SH::rows_of_seat(0);      // first table row of the seat / of the opponent
    constexpr int OPP0 = (P == 0) ? SH::rows_of_seat(0) : 0;
    constexpr int NOWN = SH::rows_of_seat(P), NOPP = SH::rows_of_seat(OPP);
    constexpr size_t kBoardFloats = (size_t)ROWS * kLdb;
    const unsigned char* blob_g = reinterpret_cast<const unsigned char*>(G.tables);
    // tables the two seats' strategies come from (evaluation: regret matching of `regret` or the rows of `avg`)
    const float* tab_opp = (EVAL && a.src_opp >= 1) ? G.avg : G.regret;
    const float* tab_own = (EVAL && a.src_own >= 1) ? G.avg : G.regret;
    const int asis_opp = (EVAL && a.src_opp == 1) ? 1 : 0, asis_own = (EVAL && a.src_own == 1) ? 1 : 0;
    const bool do_avg = !EVAL && !DEFER && a.iter >= a.delay;
    const bool defer_now = DEFER && a.defer_w != 0.0f;
    const bool read_avg = do_avg && a.m_old != 0.0f;

    // private chance-sum accumulators of this CTA (global, L2-resident): [2][kRange] int64
    long long* wp = reinterpret_cast<long long*>(G.w_private) + (size_t)blockIdx.x * 2 * kRange;
    for (int h = tid; h < 2 * kRange; h += kThreads) wp[h] = 0;
    for (int v = 0; v < kNVec; ++v)
        for (int i = kLive + tid; i < kLdb; i += kThreads) S[v * kLdb + i] = 0.0f;  // incl. the always-zero slot
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_async_shared();
    }
    __syncthreads();
    int j = blockIdx.x;
    // rows of board jj into L2 ahead of their use (one bulk prefetch per contiguous piece; nothing waits on them)
    auto prefetch_rows = [&](int jj) {
        bulk_prefetch_l2(tab_opp + (size_t)jj * kBoardFloats + (size_t)OPP0 * kLdb, NOPP * kLdb * 4);
        bulk_prefetch_l2(tab_own + (size_t)jj * kBoardFloats + (size_t)OWN0 * kLdb, NOWN * kLdb * 4);
        if (read_avg) bulk_prefetch_l2(G.avg + (size_t)jj * kBoardFloats + (size_t)OWN0 * kLdb, NOWN * kLdb * 4);
        if (defer_now) bulk_prefetch_l2(G.avg + (size_t)jj * kBoardFloats + (size_t)OPP0 * kLdb, NOPP * kLdb * 4);
    };
    if (tid == 0 && j < nb) {  // first board's tables
        mbar_expect_tx(&bars[0], kBlobA);
        bulk_g2s(smem + kBlobOff, blob_g + (size_t)j * kBlobBytes, kBlobA, &bars[0]);
        if (!P1ONLY) {
            mbar_expect_tx(&bars[2], kRowIdxBytes);
            bulk_g2s(smem + kRowIdxOff, blob_g + (size_t)j * kBlobBytes + kBlobA, kRowIdxBytes, &bars[2]);
        }
        prefetch_rows(j);
    }

    // P1 inputs of the thread's three strength positions (opponent rows + trunk reach).  kVP1Pipe: only the first position
    // is requested one unit ahead (8 registers live across the unit's last barrier - 24 spilled to local memory and made the
    // warp wait for the loads there); the other two are requested inside P1, one position ahead of their use.
    constexpr int kP1Ahead = kVP1Pipe ? kVP1Pipe : kPerThread;
    float p1_g[kP1Ahead][NOPP], p1_x0[kP1Ahead];
    auto p1_load_k = [&](int jj, const int16_t* sh_jj, int k, float (&g)[NOPP], float& x0) {
        // lanes past the last hand read the last hand's values (never used): unconditional loads keep the arrays in registers
        const int i = min(tid + k * kThreads, kLive - 1);
        const float* rows = tab_opp + (size_t)jj * kBoardFloats + (size_t)OPP0 * kLdb + i;
#pragma unroll
        for (int r = 0; r < NOPP; ++r) g[r] = ld_stream(rows + (size_t)r * kLdb);
        x0 = __ldg(a.trunk_reach_opp + sh_jj[i]);
    };
    auto p1_load = [&](int jj, const int16_t* sh_jj) {
#pragma unroll
        for (int k = 0; k < kP1Ahead; ++k) p1_load_k(jj, sh_jj, k, p1_g[k], p1_x0[k]);
    };

    for (int it = 0; j < nb; j += gridDim.x, ++it) {
        const int buf = it & 1;
        const unsigned char* blob = smem + kBlobOff + buf * kBlobA;
        const uint64_t* rec = reinterpret_cast<const uint64_t*>(blob);
        const int16_t* sh = reinterpret_cast<const int16_t*>(blob + kRecBytes);
        const int jn = j + gridDim.x;
        if (tid == 0 && jn < nb) {  // next board: records + hand ids into the other buffer (free since the last barrier), rows into L2
            mbar_expect_tx(&bars[buf ^ 1], kBlobA);
            bulk_g2s(smem + kBlobOff + (buf ^ 1) * kBlobA, blob_g + (size_t)jn * kBlobBytes, kBlobA, &bars[buf ^ 1]);
            prefetch_rows(jn);
        }
        const float prob = __ldg(G.board_prob + j);
        if (it == 0) {
            mbar_wait(&bars[buf], 0);
            p1_load(j, sh);
        }

        // ------------------------------------------------------------------------------------------ P1: reach, top-down
        // x[i] = reach of the OPPONENT at local node i (StrategyFiller.py:118-146); terminal rows go to S in strength order.
        // The rows were requested before the previous unit's last barrier (p1_load): their latency is off this path.
        auto p1_hand = [&](int k, const float (&gk)[NOPP], float x0k) {
            const int i = tid + k * kThreads;
            if (i < kLive) {
                float x[SH::N];
                x[0] = x0k * prob;  // the deal (StrategyFiller.py:137-140); blocked hands are not stored at all
                static_for<0, SH::N>([&](auto I) {
                    constexpr int n = decltype(I)::value;
                    constexpr int A = SH::n_children(n);
                    if constexpr (SH::kind(n) <= 1) {
                        constexpr int fc = SH::first_child(n);
                        if constexpr (SH::kind(n) == OPP) {
                            float gg[A], s[A];
#pragma unroll
                            for (int c = 0; c < A; ++c) gg[c] = gk[SH::row_of(fc + c) - OPP0];
                            node_strategy<A>(gg, asis_opp, s);
#pragma unroll
                            for (int c = 0; c < A; ++c) x[fc + c] = x[n] * s[c];
                            if constexpr (DEFER) {
                                if (defer_now) {  // avg_strat_sum += strategy * reach * weight (VanillaCFR.py:56-59, LinearCFR.py:55-58)
                                    float* arow = G.avg + (size_t)j * kBoardFloats + i;
#pragma unroll
                                    for (int c = 0; c < A; ++c) {
                                        float* ap = arow + (size_t)SH::row_of(fc + c) * kLdb;
                                        st_stream(ap, __fadd_rn(ld_stream(ap), __fmul_rn(x[fc + c], a.defer_w)));
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < A; ++c) x[fc + c] = x[n];
                        }
                    } else if constexpr (SH::kind(n) == 4) {
                        S[SH::vec_index(n) * kLdb + i] = x[n];
                    } else {
                        S[(NSD + SH::vec_index(n)) * kLdb + i] = x[n];
                    }
                });
            }
        };
        if constexpr (kVP1Pipe == 1) {
            float gB[NOPP], gC[NOPP], xB = 0.0f, xC = 0.0f;
            p1_load_k(j, sh, 1, gB, xB);
            p1_hand(0, p1_g[0], p1_x0[0]);
            p1_load_k(j, sh, 2, gC, xC);
            p1_hand(1, gB, xB);
            p1_hand(2, gC, xC);
        } else {
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) p1_hand(k, p1_g[k], p1_x0[k]);
        }
        __syncthreads();  // B1: S complete
        if constexpr (!P1ONLY) {

        // ------------------------------------------------------------------------------------------ P2a: card rows
        // quad (live card lc, lane q): entries [12 q, 12 q + 12) of the card's row in strength order.  Showdown vectors:
        // centred exclusive prefix sums Er[v][lc][k] = (mass of the k weakest hands holding the card) - half the row's mass;
        // fold vectors: the row's mass cs[f][lc].  Two groups of 47 quads share the nine vectors (SD 0-2 + fold 0-1 | SD 3-4 +
        // fold 2-3); the other threads start on the main scans, which only read S as well.
        mbar_wait(&bars[2], it & 1);
        float a0[NSD], a1[NSD], a2[NSD];
        double pre[NSD];
        constexpr int kQuadThreads = kLiveCards * 4;  // 188
        if (warp < (2 * kQuadThreads + 31) / 32) {   // whole warps (the quad shuffles name every lane)
            const int grp = (tid >= kQuadThreads) ? 1 : 0;
            const int t = tid - grp * kQuadThreads;
            const bool row_live = t < kQuadThreads;
            const int lc = row_live ? (t >> 2) : (kLiveCards - 1), q = tid & 3;
            const unsigned qmask = 0xFu << (lane & 28);  // the two groups run different trip counts: shuffles name the quad only
            const uint2* rp = reinterpret_cast<const uint2*>(rowidx + lc * kRowPad + q * kRowSeg);
            const uint2 w0 = rp[0], w1 = rp[1], w2 = rp[2];
            const unsigned pk[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
            int idx[kRowSeg];
#pragma unroll
            for (int e = 0; e < kRowSeg; ++e) idx[e] = (pk[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
            const int v_lo = grp ? 3 : 0, v_hi = grp ? NSD : 3;
#pragma unroll 1
            for (int v = v_lo; v < v_hi; ++v) {
                const float* Sv = S + v * kLdb;
                float inc[kRowSeg];
                float run = 0.0f;
                double drun = 0.0;  // kLin: the row's total in double (the fold vectors' row sums derive from these)
#pragma unroll
                for (int e = 0; e < kRowSeg; ++e) {
                    const float xv = Sv[idx[e]];
                    inc[e] = run;
                    run += xv;
                    if constexpr (kLin && !PRL_BV_ROWTOTF) drun += (double)xv;
                }
                if constexpr (kLin) {
                    if constexpr (PRL_BV_ROWTOTF) drun = (double)run;
                    drun += __shfl_xor_sync(qmask, drun, 1, 4);
                    drun += __shfl_xor_sync(qmask, drun, 2, 4);
                    if (row_live && q == 0) rowtot[v * kRowPad + lc] = drun;
                }
                float sc = run;  // inclusive scan over the quad
                float tt = __shfl_up_sync(qmask, sc, 1, 4);
                if (q >= 1) sc += tt;
                tt = __shfl_up_sync(qmask, sc, 2, 4);
                if (q >= 2) sc += tt;
                const float half = 0.5f * __shfl_sync(qmask, sc, 3, 4);
                const float off = (sc - run) - half;
                float* row = Er + v * kErVec + (kVErT ? q * kRowSeg * kErStride + lc : lc * kErStride + q * kRowSeg);
#pragma unroll
                for (int e = 0; e < kRowSeg; ++e)
                    if (row_live && q * kRowSeg + e < kErStride) row[kVErT ? e * kErStride : e] = off + inc[e];
            }
#pragma unroll 1
            for (int f = 2 * grp; f < (kLin ? 0 : 2 * grp + 2); ++f) {  // kLin: nothing to gather for the fold vectors
                const float* Sv = S + (NSD + f) * kLdb;
                double run = 0.0;  // double: the fold value subtracts these sums from the total (cancellation)
#pragma unroll
                for (int e = 0; e < kRowSeg; ++e) run += (double)Sv[idx[e]];
                run += __shfl_xor_sync(qmask, run, 1, 4);
                run += __shfl_xor_sync(qmask, run, 2, 4);
                if (row_live && q == 0) csd[f * kRowPad + lc] = run;
            }
        }
        __syncwarp();
        // ------------------------------------------------------------------------------------------ P2b: main scans, part 1
        // centred exclusive prefix sums over the strength order: E[k] = (mass of the k weakest hands) - total / 2, k = 0 ..
        // 1081, sums in double (the showdown value is a difference of two prefixes).
        // Thread t owns positions 3t .. 3t+2.
        {
            const int b0 = 3 * tid;
#pragma unroll
            for (int v = 0; v < NSD; ++v) {
                const float* Sv = S + v * kLdb;
                a0[v] = (b0 < kLive) ? Sv[b0] : 0.0f;
                a1[v] = (b0 + 1 < kLive) ? Sv[b0 + 1] : 0.0f;
                a2[v] = (b0 + 2 < kLive) ? Sv[b0 + 2] : 0.0f;
                const double loc = ((double)a0[v] + (double)a1[v]) + (double)a2[v];
                double incw = loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double tt = __shfl_up_sync(0xffffffffu, incw, o);
                    if (lane >= o) incw += tt;
                }
                pre[v] = incw - loc;  // exclusive within the warp
                if (lane == 31) wsum[v * 16 + warp] = incw;
            }
        }
        // P3 inputs of the first strength position: requested here, consumed after the prefix sums are written back
        const float mult = __ldg(G.board_mult + j);
        const double fx = (double)mult * a.fx_scale;
        const float* own_rows = tab_own + (size_t)j * kBoardFloats + (size_t)OWN0 * kLdb;
        float* reg_rows = G.regret + (size_t)j * kBoardFloats + (size_t)OWN0 * kLdb;
        float* avg_rows = G.avg + (size_t)j * kBoardFloats + (size_t)OWN0 * kLdb;
        auto p3_pos = [&](int k) -> int { return tid + k * kThreads; };  // strength position of the thread's k-th hand of P3
        float gA[NOWN], aA[NOWN], gB[NOWN], aB[NOWN];
        auto p3_load = [&](int k, float (&g)[NOWN], float (&av)[NOWN]) {
            const int i = min(p3_pos(k), kLive - 1);  // unconditional (see p1_load_k)
#pragma unroll
            for (int r = 0; r < NOWN; ++r) g[r] = ld_stream(own_rows + (size_t)r * kLdb + i);
#pragma unroll
            for (int r = 0; r < NOWN; ++r) av[r] = 0.0f;
            if (read_avg) {
#pragma unroll
                for (int r = 0; r < NOWN; ++r) av[r] = ld_stream(avg_rows + (size_t)r * kLdb + i);
            }
        };
        p3_load(0, gA, aA);
        __syncthreads();  // B2: card rows done, S may be overwritten, the row table may be replaced
        if (tid == 0 && jn < nb) {
            mbar_expect_tx(&bars[2], kRowIdxBytes);
            bulk_g2s(smem + kRowIdxOff, blob_g + (size_t)jn * kBlobBytes + kBlobA, kRowIdxBytes, &bars[2]);
        }
        // card-row sums cs[f][card] and total tf[f] of fold vector f (total = half the sum of its card rows): from the gathered
        // sums csd, or - kLin - as the combination FoldLinFHP of the showdown vectors' row totals
        auto fold_finish = [&](auto F) {
            constexpr int f = decltype(F)::value;
            double c0 = 0.0, c1 = 0.0;
            if constexpr (kLin) {
                static_assert(std::is_same<SH, ShapeFHP>::value, "FoldLinFHP belongs to ShapeFHP");
                static_for<0, NSD>([&](auto V) {
                    constexpr int v = decltype(V)::value;
                    constexpr int cf = FoldLinFHP::coef(P, f, v);
                    if constexpr (cf != 0) {
                        const double r0 = (lane < kLiveCards) ? rowtot[v * kRowPad + lane] : 0.0;
                        const double r1 = (lane + 32 < kLiveCards) ? rowtot[v * kRowPad + lane + 32] : 0.0;
                        c0 = (cf > 0) ? c0 + r0 : c0 - r0;
                        c1 = (cf > 0) ? c1 + r1 : c1 - r1;
                    }
                });
            } else {
                c0 = (lane < kLiveCards) ? csd[f * kRowPad + lane] : 0.0;
                c1 = (lane + 32 < kLiveCards) ? csd[f * kRowPad + lane + 32] : 0.0;
            }
            double s2 = c0 + c1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (lane == 0) tf[f] = (float)(0.5 * s2);
            if (lane < kLiveCards) cs[f * kRowPad + lane] = (float)c0;  // float copies for the per-hand epilogue
            if (lane + 32 < kLiveCards) cs[f * kRowPad + lane + 32] = (float)c1;
        };
        {
            auto warp_totals_scan = [&](int v) {  // exclusive scan of the 12 warp totals of vector v
                const double w = (lane < kWarps) ? wsum[v * 16 + lane] : 0.0;
                double sc = w;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const double tt = __shfl_up_sync(0xffffffffu, sc, o);
                    if (lane >= o) sc += tt;
                }
                const double total = __shfl_sync(0xffffffffu, sc, kWarps - 1);
                if (lane < kWarps) wexc[v * 16 + lane] = (sc - w) - 0.5 * total;
            };
            if (warp == 0) {
#pragma unroll
                for (int v = 0; v < NSD; ++v) warp_totals_scan(v);
            } else if (warp == 1) {
                static_for<0, NF>([&](auto F) { fold_finish(F); });
            }
            __syncthreads();  // B3
            const int b0 = 3 * tid;
#pragma unroll
            for (int v = 0; v < NSD; ++v) {
                float* Sv = S + v * kLdb;
                double run = pre[v] + wexc[v * 16 + warp];
                if (b0 <= kLive) Sv[b0] = (float)run;
                run += (double)a0[v];
                if (b0 + 1 <= kLive) Sv[b0 + 1] = (float)run;
                run += (double)a1[v];
                if (b0 + 2 <= kLive) Sv[b0 + 2] = (float)run;
            }
        }
        __syncthreads();  // B4: prefix arrays complete

        // ------------------------------------------------------------------------------------------ P3: values, bottom-up
        auto p3_hand = [&](int k, const float (&gown)[NOWN], const float (&av)[NOWN]) {
            const int i = p3_pos(k);
            const uint64_t w = rec[i];
            const int hand = sh[i];
            long long* wacc = wp + hand;
            long long w_ev = 0, w_br = 0;
            if constexpr (!kVRed) {
                w_ev = *wacc;
                if constexpr (EVAL) w_br = wacc[kRange];
            }
            const int gs = (int)(w & 0x7ffu), ge = (int)((w >> 11) & 0x7ffu);
            const int lc1 = (int)((w >> 22) & 0x3fu), lc2 = (int)((w >> 28) & 0x3fu);
            const int k1 = (int)((w >> 34) & 0x3fu), t1 = (int)((w >> 40) & 0x3fu), k2 = (int)((w >> 46) & 0x3fu), t2 = (int)((w >> 52) & 0x3fu);
            const int o1 = kVErT ? k1 * kErStride + lc1 : lc1 * kErStride + k1, o1e = o1 + (kVErT ? t1 * kErStride : t1);
            const int o2 = kVErT ? k2 * kErStride + lc2 : lc2 * kErStride + k2, o2e = o2 + (kVErT ? t2 * kErStride : t2);
            float e[SH::N], br[EVAL ? SH::N : 1];
            // terminal rows (ValueFiller.py:103-158): ev = equity * K * pot / 2, the folder loses
            static_for<0, SH::N>([&](auto I) {
                constexpr int n = decltype(I)::value;
                if constexpr (SH::kind(n) == 4) {
                    constexpr int v = SH::vec_index(n);
                    const float* Ev = S + v * kLdb;
                    const float* Rv = Er + v * kErVec;
                    const float all = Ev[gs] + Ev[ge];
                    const float rows = (Rv[o1] + Rv[o1e]) + (Rv[o2] + Rv[o2e]);
                    e[n] = (all - rows) * a.sc[n];
                    if constexpr (EVAL) br[n] = e[n];
                } else if constexpr (SH::kind(n) == 3) {
                    constexpr int f = SH::vec_index(n);
                    const float mass = ((tf[f] - cs[f * kRowPad + lc1]) - cs[f * kRowPad + lc2]) + S[(NSD + f) * kLdb + i];
                    e[n] = mass * a.sc[n];
                    if constexpr (EVAL) br[n] = e[n];
                }
            });
            // decision nodes, deepest first (ValueFiller.py:64-93); regrets / matching / average at P's nodes
            static_for_down<0, SH::N>([&](auto I) {
                constexpr int n = decltype(I)::value;
                constexpr int A = SH::n_children(n);
                if constexpr (SH::kind(n) <= 1) {
                    constexpr int fc = SH::first_child(n);
                    if constexpr (SH::kind(n) == OPP) {
                        float v = e[fc];
#pragma unroll
                        for (int c = 1; c < A; ++c) v += e[fc + c];
                        e[n] = v;
                        if constexpr (EVAL) {
                            float b = br[fc];
#pragma unroll
                            for (int c = 1; c < A; ++c) b += br[fc + c];
                            br[n] = b;
                        }
                    } else {
                        constexpr int r0 = SH::row_of(fc) - OWN0;  // rows r0 .. r0 + A - 1 of the seat's block
                        float g[A], s[A];
#pragma unroll
                        for (int c = 0; c < A; ++c) g[c] = gown[r0 + c];
                        node_strategy<A>(g, asis_own, s);
                        float v = s[0] * e[fc];
#pragma unroll
                        for (int c = 1; c < A; ++c) v += s[c] * e[fc + c];
                        e[n] = v;
                        if constexpr (EVAL) {
                            float b = br[fc];
#pragma unroll
                            for (int c = 1; c < A; ++c) b = fmaxf(b, br[fc + c]);
                            br[n] = b;
                        } else {
#pragma unroll
                            for (int c = 0; c < A; ++c) {
                                if constexpr (DEFER) g[c] = __fadd_rn(__fmul_rn(a.rw, e[fc + c] - v), g[c]);  // VanillaCFR.py:26-27, LinearCFR.py:27-28
                                else g[c] = fmaxf((e[fc + c] - v) + g[c], 0.0f);                          // CFRPlus.py:37-41
                            }
                            if constexpr (!DEFER) node_strategy<A>(g, 0, s);
#pragma unroll
                            for (int c = 0; c < A; ++c) st_stream(reg_rows + (size_t)(r0 + c) * kLdb + i, g[c]);
                            if (do_avg) {  // CFRPlus.py:65-87 (not reach-weighted)
#pragma unroll
                                for (int c = 0; c < A; ++c)
                                    st_stream(avg_rows + (size_t)(r0 + c) * kLdb + i, a.m_old * av[r0 + c] + a.m_new * s[c]);
                            }
                        }
                    }
                }
            });
            // the board's contribution to its parent's sum (ValueFiller.py:76-78), 64-bit fixed point
            if constexpr (kVRed) {  // fire-and-forget 64-bit RED into the CTA's private vector: nothing to wait for
                atomicAdd(reinterpret_cast<unsigned long long*>(wacc), (unsigned long long)__double2ll_rn((double)e[0] * fx));
                if constexpr (EVAL)
                    atomicAdd(reinterpret_cast<unsigned long long*>(wacc + kRange), (unsigned long long)__double2ll_rn((double)br[0] * fx));
            } else {
                *wacc = w_ev + __double2ll_rn((double)e[0] * fx);
                if constexpr (EVAL) wacc[kRange] = w_br + __double2ll_rn((double)br[0] * fx);
            }
        };
        // software pipeline over the thread's three strength positions: the next position's rows are in flight while the
        // current one is evaluated (position 0 was requested before the prefix sums were written back)
        p3_load(1, gB, aB);
        p3_hand(0, gA, aA);
        p3_load(2, gA, aA);
        if (p3_pos(1) < kLive) p3_hand(1, gB, aB);
        if (p3_pos(2) < kLive) p3_hand(2, gA, aA);
        }  // !P1ONLY
        // next unit's P1 inputs: requested before the barrier below, consumed after it (the other table buffer is long there)
        if (jn < nb) {
            mbar_wait(&bars[buf ^ 1], ((it + 1) >> 1) & 1);
            p1_load(jn, reinterpret_cast<const int16_t*>(smem + kBlobOff + (buf ^ 1) * kBlobA + kRecBytes));
        }
        __syncthreads();  // B5: S / Er / tables of this board are free
    }
    // merge into the device-wide sums (integer adds: exact in any order)
    __syncthreads();
    unsigned long long* wt = reinterpret_cast<unsigned long long*>(G.w_total) + (EVAL ? 2 * P * kRange : 0);  // eval: [seat][ev, ev_br]
    for (int h = tid; h < (EVAL ? 2 : 1) * kRange; h += kThreads) {
        const long long v = __ldcg(wp + h);  // L2: where the REDs landed
        if (v != 0) atomicAdd(wt + h, (unsigned long long)v);
    }
}

// =====================================================================================================================
// Per-board tables from the hand strengths (prl_hand_rank_boards): strength order, packed records, card rows.
// One CTA per board.  blob layout: uint64 rec[1088] | int16 hand[1088] | int16 rowidx[47][48]
//   rec[s] (s = strength position, ties ordered by hand id): bits 0-10 gs (# strictly weaker), 11-21 ge (# weaker or equal),
//          22-27 / 28-33 compact index of the hand's first / second card among the 47 live cards, 34-39 # hands of the first
//          card's row strictly weaker, 40-45 # tied in that row (incl. the hand itself), 46-51 / 52-57 same for the second card
// =====================================================================================================================
__global__ void __launch_bounds__(256) board_tables_kernel(const int32_t* __restrict__ ranks, const uint64_t* __restrict__ board_mask,
                                                           const int8_t* __restrict__ hand_cards, int n_boards,
                                                           unsigned char* __restrict__ out) {
    __shared__ int srk[kRange];
    __shared__ short spos[kRange], sgs[kRange], sge[kRange];
    const int b = blockIdx.x;
    const uint64_t bm = board_mask[b];
    unsigned char* blob = out + (size_t)b * kBlobBytes;
    uint64_t* rec = reinterpret_cast<uint64_t*>(blob);
    int16_t* sh = reinterpret_cast<int16_t*>(blob + kRecBytes);
    int16_t* rowidx = reinterpret_cast<int16_t*>(blob + kBlobA);
    for (int h = threadIdx.x; h < kRange; h += blockDim.x) srk[h] = ranks[(size_t)b * kRange + h];
    for (int i = threadIdx.x; i < kLdb; i += blockDim.x) {
        rec[i] = 0;
        sh[i] = 0;
    }
    for (int i = threadIdx.x; i < kLiveCards * kRowPad; i += blockDim.x) rowidx[i] = (int16_t)kZeroSlot;
    __syncthreads();
    for (int h = threadIdx.x; h < kRange; h += blockDim.x) {
        const int r = srk[h];
        int lt = 0, le = 0, tb = 0;
        if (r >= 0) {
            for (int q = 0; q < kRange; ++q) {
                const int x = srk[q];
                if (x < 0) continue;
                lt += x < r;
                le += x <= r;
                tb += (x == r) && (q < h);
            }
        }
        sgs[h] = (short)(r >= 0 ? lt : -1);
        sge[h] = (short)(r >= 0 ? le : -1);
        spos[h] = (short)(r >= 0 ? lt + tb : -1);
    }
    __syncthreads();
    // card rows: item (card c, other card x)
    for (int it = threadIdx.x; it < kDeck * (kDeck - 1); it += blockDim.x) {
        const int c = it / (kDeck - 1), xr = it % (kDeck - 1);
        const int x = xr + (xr >= c);
        if (((bm >> c) | (bm >> x)) & 1ull) continue;
        const int c1 = min(c, x), c2 = max(c, x);
        const int h = c1 * (2 * kDeck - 1 - c1) / 2 + (c2 - c1 - 1);
        const int g = sgs[h], ps = spos[h];
        int lt = 0, le = 0, kpos = 0;
        for (int y = 0; y < kDeck; ++y) {
            if (y == c || ((bm >> y) & 1ull)) continue;
            const int d1 = min(c, y), d2 = max(c, y);
            const int h2 = d1 * (2 * kDeck - 1 - d1) / 2 + (d2 - d1 - 1);
            const int g2 = sgs[h2];
            lt += g2 < g;
            le += g2 <= g;
            kpos += spos[h2] < ps;
        }
        const int lc = c - __popcll(bm & ((1ull << c) - 1ull));
        rowidx[lc * kRowPad + kpos] = (int16_t)ps;
        // this card's fields of the hand's record (first card = the smaller id)
        const uint64_t fld = ((uint64_t)lc) << (c == c1 ? 22 : 28) | ((uint64_t)lt) << (c == c1 ? 34 : 46) |
                             ((uint64_t)(le - lt)) << (c == c1 ? 40 : 52);
        atomicOr(reinterpret_cast<unsigned long long*>(rec + ps), (unsigned long long)fld);
    }
    __syncthreads();
    for (int h = threadIdx.x; h < kRange; h += blockDim.x) {
        const int ps = spos[h];
        if (ps < 0) continue;
        sh[ps] = (int16_t)h;
        atomicOr(reinterpret_cast<unsigned long long*>(rec + ps),
                 (unsigned long long)((uint64_t)sgs[h] | ((uint64_t)sge[h] << 11)));
    }
}

// chance-node rows from the fixed-point sums: out[a][h] = 2^-frac * sum over the suit permutations s of W[a][s(h)]
// (integer sum: exact), or W[a][h] itself without isomorphism (suit symmetrisation)
__global__ void board_collect_kernel(const long long* __restrict__ w_total, int n_arr, const int16_t* __restrict__ sym_perm,
                                     int n_sym, double inv_scale, float* __restrict__ out, int ld) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= kRange) return;
    for (int a = 0; a < n_arr; ++a) {
        long long s = 0;
        if (n_sym > 1) {
            for (int q = 0; q < n_sym; ++q) s += w_total[(size_t)a * kRange + sym_perm[(size_t)q * kRange + h]];
        } else {
            s = w_total[(size_t)a * kRange + h];
        }
        out[(size_t)a * ld + h] = (float)((double)s * inv_scale);
    }
}

// strength-ordered rows <-> natural-order rows of ONE table (interfaces to the level engine, checkpoints, agents)
__global__ void board_permute_kernel(const unsigned char* __restrict__ tables, int n_boards, int rows_per_board,
                                     const int64_t* __restrict__ row_src, const int64_t* __restrict__ row_dst, float* sorted_tab,
                                     float* natural_tab, int ld, int to_natural) {
    // block = (board j, row r of the board); row_src[r] / row_dst[r]: row index on board 0 and stride per board packed
    const int j = blockIdx.x / rows_per_board, r = blockIdx.x % rows_per_board;
    const int16_t* sh = reinterpret_cast<const int16_t*>(tables + (size_t)j * kBlobBytes + kRecBytes);
    float* srow = sorted_tab + ((size_t)row_src[2 * r] + (size_t)j * row_src[2 * r + 1]) * kLdb;
    float* nrow = natural_tab + ((size_t)row_dst[2 * r] + (size_t)j * row_dst[2 * r + 1]) * ld;
    if (to_natural) {
        for (int h = threadIdx.x; h < ld; h += blockDim.x) nrow[h] = 0.0f;
        __syncthreads();
        for (int i = threadIdx.x; i < kLive; i += blockDim.x) nrow[sh[i]] = srow[i];
    } else {
        for (int i = threadIdx.x; i < kLive; i += blockDim.x) srow[i] = nrow[sh[i]];
    }
}

// =====================================================================================================================
// The pre-deal trunk in ONE launch (one CTA): chance-node rows from the fixed-point sums (integer sum over the suit
// permutations), fold terminals, value backup, and - update form - regrets / matching / average of seat p's trunk nodes and
// its new reach rows; evaluation form: values + best response of both seats and the root exploitability.  Same statements as
// the level kernels (ValueFiller.py:64-125, CFRPlus.py:37-87, StrategyFiller.py:118-146) on <= 8 nodes in natural hand order.
// =====================================================================================================================
constexpr int kTrunkThreads = 1024;

__device__ __forceinline__ float trunk_sigma(const prl_trunk_t& t, int src, int fs, int c, int A, int h) {
    if (src == PRL_STRAT_UNIFORM64) return 1.0f / (float)A;
    if (src == PRL_STRAT_AVG_SUM) {  // reach-weighted sums, normalised on the fly (LinearCFR.py:64-71, VanillaCFR.py:65-72)
        float tot = 0.0f;
        for (int k = 0; k < A; ++k) tot += t.avg[(size_t)(fs + k) * t.ld + h];
        return (tot == 0.0f) ? 1.0f / (float)A : t.avg[(size_t)(fs + c) * t.ld + h] / tot;
    }
    const float* tab = (src == PRL_STRAT_F32) ? t.strat : t.avg;
    return tab[(size_t)(fs + c) * t.ld + h];
}

// peers != NULL: the cross-GPU sum is done HERE - every rank's fixed-point vector sits in symmetric (peer-mapped) memory and
// is read over NVLink with coalesced 8-byte loads, rank 0 .. n_peers-1 in order (integers: any order gives the same bits),
// into w_scratch; the caller has placed a cross-rank barrier between the sweep kernels and this launch.
template <bool EVAL>
__global__ void __launch_bounds__(kTrunkThreads) trunk_kernel(const prl_trunk_t t, const long long* __restrict__ w_total,
                                                              const long long* const* __restrict__ peers, int n_peers,
                                                              long long peer_offset, long long* __restrict__ w_scratch,
                                                              const int16_t* __restrict__ sym_perm, int n_sym, double inv_scale,
                                                              int p_upd, int iter, int delay, float m_old, float m_new, int algo,
                                                              float rw, float* out_expl) {
    __shared__ float ro[kRange + 2];
    __shared__ float cs[64];
    __shared__ float red[32];
    __shared__ double dred[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t N = (size_t)t.n_buf_nodes, ld = (size_t)t.ld;
    const int seat_lo = EVAL ? 0 : p_upd, seat_hi = EVAL ? 1 : p_upd;
    if (peers != nullptr) {
        for (int i = tid; i < (EVAL ? 4 : 1) * kRange; i += kTrunkThreads) {
            long long s = 0;
            for (int r = 0; r < n_peers; ++r) s += peers[r][peer_offset + i];
            w_scratch[i] = s;
        }
        __syncthreads();
        w_total = w_scratch;
    }
    // 1. the chance node's rows
    for (int p = seat_lo; p <= seat_hi; ++p)
        for (int k = 0; k < (EVAL ? 2 : 1); ++k) {
            const long long* W = w_total + (size_t)(EVAL ? (2 * p + k) : 0) * kRange;
            float* dst = (k ? t.ev_br : t.ev) + ((size_t)p * N + t.chance_node) * ld;
            for (int h = tid; h < kRange; h += kTrunkThreads) {
                long long s = 0;
                if (n_sym > 1) {
                    for (int q = 0; q < n_sym; ++q) s += W[sym_perm[(size_t)q * kRange + h]];
                } else {
                    s = W[h];
                }
                dst[h] = (float)((double)s * inv_scale);
            }
        }
    // 2. fold terminals (no board): ValueFiller.py:103-113 for two-card hands
    for (int n = 0; n < t.n_nodes; ++n) {
        if (t.kind[n] != PRL_KIND_FOLD) continue;
        for (int p = seat_lo; p <= seat_hi; ++p) {
            const float* rg = t.reach + ((size_t)(1 - p) * N + n) * ld;
            __syncthreads();
            float part = 0.0f;
            for (int h = tid; h < kRange; h += kTrunkThreads) {
                const float r = rg[h];
                ro[h] = r;
                part += r;
            }
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) red[warp] = part;
            __syncthreads();
            float T = 0.0f;
            for (int w = 0; w < kTrunkThreads / 32; ++w) T += red[w];
            if (tid < kDeck) {  // per-card sums over the lexicographic range layout, fixed order
                float acc = 0.0f;
                const int c = tid;
                for (int r = 0; r < c; ++r) acc += ro[r * (2 * kDeck - 1 - r) / 2 + c - r - 1];
                const int b0 = c * (2 * kDeck - 1 - c) / 2;
                for (int k = 0; k < kDeck - 1 - c; ++k) acc += ro[b0 + k];
                cs[c] = acc;
            }
            __syncthreads();
            const float sc = t.eq_const * t.pot[n] * 0.5f * ((t.acted_last[n] == p) ? -1.0f : 1.0f);
            float* e = t.ev + ((size_t)p * N + n) * ld;
            float* b = t.ev_br + ((size_t)p * N + n) * ld;
            for (int h = tid; h < kRange; h += kTrunkThreads) {
                const float v = (T - cs[t.hand_cards[2 * h]] - cs[t.hand_cards[2 * h + 1]] + ro[h]) * sc;
                e[h] = v;
                if (EVAL) b[h] = v;
            }
        }
    }
    __syncthreads();
    // 3. decision nodes bottom-up, per hand (children have larger ids); 4. regrets / matching / average; 5. reach of seat p
    double ex[2] = {0.0, 0.0};
    for (int h = tid; h < kRange; h += kTrunkThreads) {
        for (int n = t.n_nodes - 1; n >= 0; --n) {
            const int k = t.kind[n];
            if (k > PRL_KIND_P1) continue;
            const int A = t.n_children[n], fc = t.first_child[n], fs = t.first_slot[n];
            for (int p = seat_lo; p <= seat_hi; ++p) {
                float* ev_p = t.ev + (size_t)p * N * ld;
                float* br_p = t.ev_br + (size_t)p * N * ld;
                float v = 0.0f, b = 0.0f;
                if (k != p) {
                    for (int c = 0; c < A; ++c) v += ev_p[(size_t)(fc + c) * ld + h];
                    if (EVAL)
                        for (int c = 0; c < A; ++c) b += br_p[(size_t)(fc + c) * ld + h];
                } else {
                    for (int c = 0; c < A; ++c) v += trunk_sigma(t, t.mode[p], fs, c, A, h) * ev_p[(size_t)(fc + c) * ld + h];
                    if (EVAL) {
                        b = br_p[(size_t)fc * ld + h];
                        for (int c = 1; c < A; ++c) b = fmaxf(b, br_p[(size_t)(fc + c) * ld + h]);
                    } else {  // CFRPlus.py:37-63; VanillaCFR.py:26-52 / LinearCFR.py:27-51: weighted, unclipped, matching on the positive part
                        float ssum = 0.0f;
                        for (int c = 0; c < A; ++c) {
                            float* rg = t.regret + (size_t)(fs + c) * ld + h;
                            const float d = ev_p[(size_t)(fc + c) * ld + h] - v;
                            const float r = (algo == PRL_ALGO_CFR_PLUS) ? fmaxf(d + *rg, 0.0f) : __fadd_rn(__fmul_rn(rw, d), *rg);
                            *rg = r;
                            ssum += fmaxf(r, 0.0f);
                        }
                        const float inv = (ssum > 0.0f) ? 1.0f / ssum : 0.0f;
                        for (int c = 0; c < A; ++c) {
                            const float r = fmaxf(t.regret[(size_t)(fs + c) * ld + h], 0.0f);
                            t.strat[(size_t)(fs + c) * ld + h] = (ssum > 0.0f) ? r * inv : 1.0f / (float)A;
                        }
                    }
                }
                ev_p[(size_t)n * ld + h] = v;
                if (EVAL) br_p[(size_t)n * ld + h] = b;
            }
        }
        if (EVAL) {
            for (int p = 0; p < 2; ++p)  // ValueFiller.py:95-101 at the root
                ex[p] += (double)t.reach[((size_t)p * N) * ld + h] *
                         ((double)t.ev_br[((size_t)p * N) * ld + h] - (double)t.ev[((size_t)p * N) * ld + h]);
        } else {
            const int p = p_upd;
            float* rp = t.reach + (size_t)p * N * ld;
            for (int n = 0; n < t.n_nodes; ++n) {  // StrategyFiller.py:118-146 with the new strategy; average CFRPlus.py:65-87
                const int k = t.kind[n];
                if (k > PRL_KIND_P1) continue;
                const int A = t.n_children[n], fc = t.first_child[n], fs = t.first_slot[n];
                const float r = rp[(size_t)n * ld + h];
                for (int c = 0; c < A; ++c) {
                    float s = 1.0f;
                    if (k == p) {
                        s = t.strat[(size_t)(fs + c) * ld + h];
                        float* a = t.avg + (size_t)(fs + c) * ld + h;
                        if (algo != PRL_ALGO_CFR_PLUS) *a = __fadd_rn(*a, __fmul_rn(__fmul_rn(s, r), rw));  // VanillaCFR.py:56-59, LinearCFR.py:55-58
                        else if (iter >= delay) *a = m_old * (*a) + m_new * s;
                    }
                    rp[(size_t)(fc + c) * ld + h] = s * r;
                }
            }
        }
    }
    if (EVAL) {
        for (int p = 0; p < 2; ++p) {
            double v = ex[p];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (lane == 0) dred[warp] = v;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < kTrunkThreads / 32; ++w) s += dred[w];
                out_expl[p] = (float)s;
            }
        }
    }
}

bool shape_matches(const prl_board_game_t* g) {
    if (g->n_local != ShapeFHP::N) return false;
    for (int i = 0; i < ShapeFHP::N; ++i)
        if (g->kind[i] != ShapeFHP::kind(i) || g->parent[i] != ShapeFHP::parent(i) || g->first_child[i] != ShapeFHP::first_child(i) ||
            g->n_children[i] != ShapeFHP::n_children(i))
            return false;
    return true;
}

// the table layout the kernel compiles in: row(i, j) = j * 14 + row_of(i)
bool layout_matches(const prl_board_game_t* g) {
    for (int i = 1; i < ShapeFHP::N; ++i)
        if (g->row0[i] != ShapeFHP::row_of(i) || g->row_m[i] != ShapeFHP::rows) return false;
    return true;
}

int default_grid() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = 2 * sms;
    }
    return cached[dev];
}

template <int P, bool EVAL, bool DEFER = false, bool P1ONLY = false>
int launch_sweep(const SweepArgs& a, int grid, cudaStream_t s) {
    auto kern = board_sweep_kernel<ShapeFHP, P, EVAL, DEFER, P1ONLY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);  // per device: set every time
    if (e != cudaSuccess) return prl::check(e, "prl_board_sweep: shared memory opt-in");
    kern<<<grid, kThreads, kSmemBytes, s>>>(a);
    prl::count_launch();
    return 0;
}

}  // namespace

extern "C" int prl_board_layout(int32_t* out) {
    out[0] = kLive;
    out[1] = kLdb;
    out[2] = kBlobBytes;
    out[3] = kRecBytes;            // offset of the hand ids inside a board's blob
    out[4] = kBlobA;               // offset of the card rows
    out[5] = kLiveCards;
    out[6] = kRowPad;
    out[7] = ShapeFHP::N;
    return 0;
}

extern "C" int prl_board_grid(void) { return default_grid(); }

extern "C" int prl_board_rows(int32_t* row_of, int32_t* rows_per_board) {
    for (int i = 0; i < 16; ++i) row_of[i] = (i >= 1 && i < ShapeFHP::N) ? ShapeFHP::row_of(i) : -1;
    *rows_per_board = ShapeFHP::rows;
    return 0;
}

extern "C" int prl_board_shape_ok(const prl_board_game_t* g) { return (g && shape_matches(g)) ? 1 : 0; }

extern "C" int prl_board_build_tables(const int32_t* ranks, const uint64_t* board_mask, const int8_t* hand_cards, int n_boards,
                                      void* blob, prl_stream_t stream) {
    if (n_boards <= 0) return 0;
    board_tables_kernel<<<n_boards, 256, 0, (cudaStream_t)stream>>>(ranks, board_mask, hand_cards, n_boards, (unsigned char*)blob);
    prl::count_launch();
    return prl::check(cudaGetLastError(), "prl_board_build_tables");
}

extern "C" int prl_board_sweep(const prl_board_game_t* g, int p, int eval, int src_own, int src_opp, const float* trunk_reach_opp,
                               int iter, int delay, int algo, float defer_w, int p1_only, prl_stream_t stream) {
    if (algo != PRL_ALGO_CFR_PLUS && algo != PRL_ALGO_VANILLA && algo != PRL_ALGO_LINEAR) return prl::fail("prl_board_sweep: bad algo");
    const bool defer = !eval && algo != PRL_ALGO_CFR_PLUS;
    if (p1_only && !defer) return prl::fail("prl_board_sweep: p1_only is the average flush of Vanilla / Linear CFR");
    if (!g || !shape_matches(g)) return prl::fail("prl_board_sweep: the post-deal subtree does not have the compiled shape");
    if (g->n_range != kRange || g->n_deck != kDeck) return prl::fail("prl_board_sweep: 52-card deck / 1326 hands only");
    if (!layout_matches(g)) return prl::fail("prl_board_sweep: row0 / row_m must be the board-major layout of prl_board_rows");
    if (p < 0 || p > 1) return prl::fail("prl_board_sweep: bad seat");
    if (!g->tables || !g->regret || !g->avg || !g->w_private || !g->w_total || !trunk_reach_opp)
        return prl::fail("prl_board_sweep: missing buffers");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = g->grid > 0 ? g->grid : default_grid();
    SweepArgs a;
    a.g = *g;
    a.trunk_reach_opp = trunk_reach_opp;
    a.iter = iter;
    a.delay = delay;
    const double cw = 0.5 * ((double)iter * (iter + 1) - (double)delay * (delay + 1));  // CFRPlus.py:68-73
    const double nw = (double)iter - delay + 1;
    a.m_old = (iter > delay) ? (float)(cw / (cw + nw)) : 0.0f;
    a.m_new = (iter > delay) ? (float)(nw / (cw + nw)) : 1.0f;
    a.src_own = src_own;
    a.src_opp = src_opp;
    a.rw = (algo == PRL_ALGO_LINEAR) ? (float)(iter + 1) : 1.0f;
    a.defer_w = defer ? defer_w : 0.0f;
    a.fx_scale = (double)(1ull << g->frac_bits);
    for (int n = 0; n < 16; ++n) {
        const bool folder = n < g->n_local && g->kind[n] == PRL_KIND_FOLD && g->acted_last[n] == p;
        a.sc[n] = (n < g->n_local) ? g->eq_const * g->pot[n] * 0.5f * (folder ? -1.0f : 1.0f) : 0.0f;
    }
    if (p1_only) {
        const int rc1 = (p == 0) ? launch_sweep<0, false, true, true>(a, grid, s) : launch_sweep<1, false, true, true>(a, grid, s);
        return rc1 ? rc1 : prl::check(cudaGetLastError(), "prl_board_sweep(flush)");
    }
    {   // the sums this launch produces: update -> w_total[0]; evaluation of seat p -> w_total[2p], w_total[2p + 1]
        char* base = reinterpret_cast<char*>(g->w_total) + (eval ? sizeof(long long) * 2 * p * kRange : 0);
        if (int e = prl::check(cudaMemsetAsync(base, 0, sizeof(long long) * (eval ? 2 : 1) * kRange, s), "prl_board_sweep: memset")) return e;
    }
    int rc;
    if (eval) rc = (p == 0) ? launch_sweep<0, true>(a, grid, s) : launch_sweep<1, true>(a, grid, s);
    else if (defer) rc = (p == 0) ? launch_sweep<0, false, true>(a, grid, s) : launch_sweep<1, false, true>(a, grid, s);
    else rc = (p == 0) ? launch_sweep<0, false>(a, grid, s) : launch_sweep<1, false>(a, grid, s);
    if (rc) return rc;
    return prl::check(cudaGetLastError(), "prl_board_sweep");
}

extern "C" int prl_board_collect(const prl_board_game_t* g, int n_arr, const int16_t* sym_perm, int n_sym, float* out, int ld,
                                 prl_stream_t stream) {
    if (!g || n_arr < 1 || n_arr > 4) return prl::fail("prl_board_collect: bad arguments");
    board_collect_kernel<<<(kRange + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(g->w_total), n_arr, sym_perm, n_sym, 1.0 / (double)(1ull << g->frac_bits), out, ld);
    prl::count_launch();
    return prl::check(cudaGetLastError(), "prl_board_collect");
}

extern "C" int prl_board_permute(const prl_board_game_t* g, int rows_per_board, const int64_t* row_src, const int64_t* row_dst,
                                 float* sorted_tab, float* natural_tab, int ld, int to_natural, prl_stream_t stream) {
    if (!g || g->n_boards <= 0 || rows_per_board <= 0) return 0;
    board_permute_kernel<<<g->n_boards * rows_per_board, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)g->tables, g->n_boards, rows_per_board, row_src, row_dst, sorted_tab, natural_tab, ld, to_natural);
    prl::count_launch();
    return prl::check(cudaGetLastError(), "prl_board_permute");
}

// Trunk of seat p's half-iteration (eval == 0) or of an evaluation of both seats (eval != 0) in one launch; see prl_trunk_t.
extern "C" int prl_board_trunk(const prl_board_game_t* g, const prl_trunk_t* t, int eval, int p, int n_sym, const int16_t* sym_perm,
                               int iter, int delay, float* out_expl, const int64_t* const* peers, int n_peers,
                               int64_t peer_offset, int64_t* w_scratch, int algo, prl_stream_t stream) {
    if (algo != PRL_ALGO_CFR_PLUS && algo != PRL_ALGO_VANILLA && algo != PRL_ALGO_LINEAR) return prl::fail("prl_board_trunk: bad algo");
    const float rw = (algo == PRL_ALGO_LINEAR) ? (float)(iter + 1) : 1.0f;
    if (!g || !t || t->n_nodes < 1 || t->n_nodes > 8) return prl::fail("prl_board_trunk: 1..8 trunk nodes");
    if (t->n_range != kRange || g->n_deck != kDeck) return prl::fail("prl_board_trunk: 52-card deck / 1326 hands only");
    if (eval && !out_expl) return prl::fail("prl_board_trunk: out_expl missing");
    for (int n = 0; n < t->n_nodes; ++n)
        if (t->kind[n] == PRL_KIND_SHOWDOWN || t->kind[n] == PRL_KIND_SHOWDOWN_ALLIN)
            return prl::fail("prl_board_trunk: showdowns before the deal are not supported");
    const double cw = 0.5 * ((double)iter * (iter + 1) - (double)delay * (delay + 1));  // CFRPlus.py:68-73
    const double nw = (double)iter - delay + 1;
    const float m_old = (iter > delay) ? (float)(cw / (cw + nw)) : 0.0f, m_new = (iter > delay) ? (float)(nw / (cw + nw)) : 1.0f;
    const double inv_scale = 1.0 / (double)(1ull << g->frac_bits);
    const long long* w = reinterpret_cast<const long long*>(g->w_total);
    if (peers && (n_peers < 1 || !w_scratch)) return prl::fail("prl_board_trunk: peer sum needs n_peers >= 1 and w_scratch");
    const long long* const* pp = reinterpret_cast<const long long* const*>(peers);
    long long* ws = reinterpret_cast<long long*>(w_scratch);
    if (eval)
        trunk_kernel<true><<<1, kTrunkThreads, 0, (cudaStream_t)stream>>>(*t, w, pp, n_peers, (long long)peer_offset, ws, sym_perm, n_sym,
                                                                       inv_scale, -1, iter, delay, m_old, m_new, algo, rw, out_expl);
    else
        trunk_kernel<false><<<1, kTrunkThreads, 0, (cudaStream_t)stream>>>(*t, w, pp, n_peers, (long long)peer_offset, ws, sym_perm, n_sym,
                                                                        inv_scale, p, iter, delay, m_old, m_new, algo, rw, out_expl);
    prl::count_launch();
    return prl::check(cudaGetLastError(), "prl_board_trunk");
}
