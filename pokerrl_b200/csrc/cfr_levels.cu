// Level-synchronous public-tree sweeps for tabular CFR / best response (sm_100a).
//
// One thread per (node, hand).  Nodes of one depth are contiguous, the children of a node are contiguous and the
// children groups of adjacent nodes are adjacent, so a warp reads / writes contiguous runs of rows.  These sweeps are
// HBM/L2-bound vector work (no GEMM shape anywhere): the design rules that matter are coalescing and launch count.
//
// Arithmetic contract: every expression is evaluated in the reference's dtype and operation order
// (see oracle/cfr_numpy.py, which is pinned bit-for-bit against the reference).  This translation unit is compiled
// with -fmad=false so that no multiply-add is contracted; the reference (numpy) never fuses.
//
// Reference statements restated here (paths under PokerRL/):
//   reach pass      game/_/tree/_/StrategyFiller.py:118-146, 148-169
//   value pass      game/_/tree/_/ValueFiller.py:21-101, terminals :103-175
//   regrets         cfr/_CFRBase.py:146-185, cfr/CFRPlus.py:37-41, cfr/LinearCFR.py:27-31, cfr/VanillaCFR.py:26-30
//   regret matching cfr/CFRPlus.py:43-63, cfr/LinearCFR.py:33-51, cfr/VanillaCFR.py:32-52
//   averaging       cfr/CFRPlus.py:65-87, cfr/LinearCFR.py:53-76, cfr/VanillaCFR.py:54-77
#include <cuda_runtime.h>
#include <stdint.h>

#include "pokerrl_b200.h"
#include "prl_common.cuh"

namespace {

struct Ctx {
    prl_tree_t T;
    prl_buffers_t B;
    int lo, hi;       // node range of this level
    int mask;         // players to process
    int mode[2];      // strategy source per seat
    // CFR update parameters
    int algo, upd_p, iter, delay, avg_f64;
};

__device__ __forceinline__ bool mode_is_f32(int m) { return m == PRL_STRAT_F32 || m == PRL_STRAT_AVG_F32; }

// strategy probability of the child in table row `slot` (rows of the decision node start at first_slot, A rows)
__device__ __forceinline__ float strat_f32(const Ctx& c, int m, int slot, int h) {
    const float* tab = (m == PRL_STRAT_F32) ? c.B.strat : (const float*)c.B.avg;
    return tab[(size_t)slot * c.T.ld + h];
}

__device__ __forceinline__ double strat_f64(const Ctx& c, int m, int slot, int first_slot, int A, int h) {
    if (m == PRL_STRAT_UNIFORM64) return 1.0 / (double)A;
    if (m == PRL_STRAT_AVG_F64) return ((const double*)c.B.avg)[(size_t)slot * c.T.ld + h];
    // PRL_STRAT_AVG_SUM: float sums, float division, promoted to double (np.where with a float64 branch)
    const float* tab = (const float*)c.B.avg;
    float s = tab[(size_t)first_slot * c.T.ld + h];
    for (int k = 1; k < A; ++k) s = s + tab[(size_t)(first_slot + k) * c.T.ld + h];
    if (s == 0.0f) return 1.0 / (double)A;
    return (double)(tab[(size_t)slot * c.T.ld + h] / s);
}

// one-card games: chance probability of dealing this node's board given hand h (StrategyFiller.py:159-166)
__device__ __forceinline__ float chance_prob_1card(const Ctx& c, int n, int h) {
    return (h == c.T.board[n]) ? 0.0f : (float)(1.0 / (double)(c.T.n_deck - 2));
}

__device__ __forceinline__ int leduc_rank(const Ctx& c, int h, int b) {
    int r = h / c.T.n_suits;
    return (b / c.T.n_suits == r) ? c.T.pair_bonus + r : r;
}

// ------------------------------------------------------------------------------------------------ reach (top-down)
template <bool UPDATE_AVG>
__global__ void reach_level_kernel(Ctx c) {
    const int R = c.T.n_range, ld = c.T.ld;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int n = c.lo + (int)(idx / R);
    int h = (int)(idx % R);
    if (n >= c.hi) return;
    const size_t N = (size_t)c.T.n_nodes;
    const int par = c.T.parent[n];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (!(c.mask & (1 << q))) continue;
        float* reach_q = c.B.reach + (size_t)q * N * ld;
        float r;
        if (par < 0) {
            r = (float)(1.0 / (double)R);  // PublicTree.py:122-124
        } else {
            const float rp = reach_q[(size_t)par * ld + h];
            const int pk = c.T.kind[par];
            if (pk == PRL_KIND_CHANCE) {
                r = rp * chance_prob_1card(c, n, h);
            } else if (pk == q) {
                const int slot = c.T.slot[n];
                const int m = c.mode[q];
                if (mode_is_f32(m)) {
                    const float s = strat_f32(c, m, slot, h);
                    r = s * rp;
                    if (UPDATE_AVG && q == c.upd_p) {
                        if (c.algo == PRL_ALGO_CFR_PLUS) {
                            if (c.iter >= c.delay) {
                                // current_weight = sum(arange(delay+1, iter+1)); new_weight = iter - delay + 1
                                const long long cw = ((long long)c.iter * (c.iter + 1) - (long long)c.delay * (c.delay + 1)) / 2;
                                const long long nw = (long long)c.iter - c.delay + 1;
                                double m_old = (double)cw / (double)(cw + nw);
                                double m_new = (double)nw / (double)(cw + nw);
                                if (c.iter == c.delay) { m_old = 0.0; m_new = 1.0; }
                                if (c.avg_f64) {
                                    double* a = (double*)c.B.avg + (size_t)slot * ld + h;
                                    *a = m_old * (*a) + m_new * (double)s;
                                } else {
                                    float* a = (float*)c.B.avg + (size_t)slot * ld + h;
                                    *a = (float)m_old * (*a) + (float)m_new * s;
                                }
                            }
                        } else {
                            float contrib = r;  // strategy * reach[p]
                            if (c.algo == PRL_ALGO_LINEAR) contrib = contrib * (float)(c.iter + 1);
                            float* a = (float*)c.B.avg + (size_t)slot * ld + h;
                            *a = *a + contrib;
                        }
                    }
                } else {
                    const int fs = c.T.slot[c.T.first_child[par]];
                    const double s = strat_f64(c, m, slot, fs, c.T.n_children[par], h);
                    r = (float)(s * (double)rp);
                }
            } else {
                r = rp;
            }
        }
        reach_q[(size_t)n * ld + h] = r;
    }
}

// ------------------------------------------------------------------------------------------------ terminals (1 card)
__device__ __forceinline__ float terminal_equity_1card(const Ctx& c, int n, int h, int p, int kind) {
    const int R = c.T.n_range, ld = c.T.ld;
    const float* ro = c.B.reach + ((size_t)(1 - p) * c.T.n_nodes + n) * ld;  // opponent reach row
    const float K = (float)((double)c.T.n_deck / (double)(c.T.n_deck - 1));    // ValueFiller.py:19
    const int b = c.T.board[n];
    float eq;
    if (kind == PRL_KIND_FOLD) {  // ValueFiller.py:103-125
        float s = ro[0];
        for (int j = 1; j < R; ++j) s = s + ro[j];
        eq = s - ro[h];
        if (c.T.acted_last[n] == p) eq = -eq;
        eq = eq * K;
    } else if (kind == PRL_KIND_SHOWDOWN) {  // ValueFiller.py:127-158
        eq = 0.0f;
        if (h != b) {
            const int rh = leduc_rank(c, h, b);
            for (int j = 0; j < R; ++j) {
                if (j == h || j == b) continue;
                const int rj = leduc_rank(c, j, b);
                if (rh > rj) eq = eq + ro[j];
                else if (rh < rj) eq = eq - ro[j];
            }
        }
        eq = eq * K;
    } else {  // all-in before the board card: ValueFiller.py:160-175
        eq = 0.0f;
        for (int bb = 0; bb < c.T.n_deck; ++bb) {
            float e = 0.0f;
            if (h != bb) {
                const int rh = leduc_rank(c, h, bb);
                for (int j = 0; j < R; ++j) {
                    if (j == h || j == bb) continue;
                    const int rj = leduc_rank(c, j, bb);
                    if (rh > rj) e = e + ro[j];
                    else if (rh < rj) e = e - ro[j];
                }
            }
            eq = eq + e * K;
        }
        eq = eq / (float)(c.T.n_deck - 2);
    }
    if (h == b) eq = 0.0f;  // ValueFiller.py:57-59
    return eq;
}

// ------------------------------------------------------------------------------------------------ value (bottom-up)
template <bool WITH_BR, bool UPDATE>
__global__ void value_level_kernel(Ctx c) {
    const int R = c.T.n_range, ld = c.T.ld;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int n = c.lo + (int)(idx / R);
    int h = (int)(idx % R);
    if (n >= c.hi) return;
    const size_t N = (size_t)c.T.n_nodes;
    const int kind = c.T.kind[n];
    const int fc = c.T.first_child[n];
    const int A = c.T.n_children[n];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if (!(c.mask & (1 << p))) continue;
        float* ev_p = c.B.ev + (size_t)p * N * ld;
        float* evbr_p = WITH_BR ? c.B.ev_br + (size_t)p * N * ld : nullptr;
        float v, vbr = 0.0f;
        if (kind >= PRL_KIND_FOLD) {
            const float eq = terminal_equity_1card(c, n, h, p, kind);
            v = eq * c.T.pot[n] / 2.0f;  // ValueFiller.py:61
            vbr = v;
        } else if (kind == PRL_KIND_CHANCE || kind != p) {
            // chance node, or the opponent acts: plain sum over children (ValueFiller.py:76-78, 88-90)
            v = ev_p[(size_t)fc * ld + h];
            for (int k = 1; k < A; ++k) v = v + ev_p[(size_t)(fc + k) * ld + h];
            if (WITH_BR) {
                vbr = evbr_p[(size_t)fc * ld + h];
                for (int k = 1; k < A; ++k) vbr = vbr + evbr_p[(size_t)(fc + k) * ld + h];
            }
        } else {
            // p acts here (ValueFiller.py:87, 91)
            const int fs = c.T.slot[fc];
            const int m = c.mode[p];
            if (mode_is_f32(m)) {
                v = strat_f32(c, m, fs, h) * ev_p[(size_t)fc * ld + h];
                for (int k = 1; k < A; ++k) v = v + strat_f32(c, m, fs + k, h) * ev_p[(size_t)(fc + k) * ld + h];
            } else {
                double acc = strat_f64(c, m, fs, fs, A, h) * (double)ev_p[(size_t)fc * ld + h];
                for (int k = 1; k < A; ++k)
                    acc = acc + strat_f64(c, m, fs + k, fs, A, h) * (double)ev_p[(size_t)(fc + k) * ld + h];
                v = (float)acc;
            }
            if (WITH_BR) {
                vbr = evbr_p[(size_t)fc * ld + h];
                for (int k = 1; k < A; ++k) vbr = fmaxf(vbr, evbr_p[(size_t)(fc + k) * ld + h]);
            }
            if (UPDATE && p == c.upd_p) {
                // regret update (_CFRBase.py:146-185) then regret matching into the strategy table
                float* reg = c.B.regret;
                float* st = c.B.strat;
                const float w = (float)(c.iter + 1);
                float s = 0.0f;
                for (int k = 0; k < A; ++k) {
                    const size_t off = (size_t)(fs + k) * ld + h;
                    const float d = ev_p[(size_t)(fc + k) * ld + h] - v;
                    float r;
                    if (c.algo == PRL_ALGO_CFR_PLUS) r = fmaxf(d + reg[off], 0.0f);
                    else if (c.algo == PRL_ALGO_LINEAR) r = w * d + reg[off];
                    else r = d + reg[off];
                    reg[off] = r;
                    const float rp = fmaxf(r, 0.0f);
                    s = (k == 0) ? rp : s + rp;
                }
                const float uni = (float)(1.0 / (double)A);
                for (int k = 0; k < A; ++k) {
                    const size_t off = (size_t)(fs + k) * ld + h;
                    st[off] = (s > 0.0f) ? fmaxf(reg[off], 0.0f) / s : uni;
                }
            }
        }
        ev_p[(size_t)n * ld + h] = v;
        if (WITH_BR) evbr_p[(size_t)n * ld + h] = vbr;
    }
}

// root exploitability (ValueFiller.py:95-101): sequential float sums like numpy's short contiguous reduction
__global__ void root_exploitability_kernel(prl_tree_t T, prl_buffers_t B, float* out) {
    const int p = threadIdx.x;
    if (p >= 2) return;
    const size_t N = (size_t)T.n_nodes;
    const float* ev = B.ev + (size_t)p * N * T.ld;
    const float* evbr = B.ev_br + (size_t)p * N * T.ld;
    const float* reach = B.reach + (size_t)p * N * T.ld;
    float s = 0.0f;
    for (int h = 0; h < T.n_range; ++h) {
        const float e = evbr[h] * reach[h] - ev[h] * reach[h];
        s = (h == 0) ? e : s + e;
    }
    out[p] = s;
}

constexpr int kThreads = 256;

#define PRL_LAUNCH(kernel, grid, stream, ...) do { kernel<<<(grid), kThreads, 0, (stream)>>>(__VA_ARGS__); prl::count_launch(); } while (0)

inline unsigned grid_for(long long n_threads) { return (unsigned)((n_threads + kThreads - 1) / kThreads); }

int check_tree(const prl_tree_t* t) {
    if (!t || !t->level_start || t->n_hole != 1) return prl::fail("prl: level sweeps support one-hole-card games (n_hole == 1)");
    return 0;
}

}  // namespace

extern "C" int prl_reach_pass(const prl_tree_t* tree, const prl_buffers_t* buf, int player_mask, const int* strat_mode,
                              prl_stream_t stream) {
    if (int e = check_tree(tree)) return e;
    Ctx c{*tree, *buf, 0, 0, player_mask, {strat_mode[0], strat_mode[1]}, 0, -1, 0, 0, 0};
    cudaStream_t s = (cudaStream_t)stream;
    for (int d = 0; d < tree->n_levels; ++d) {
        c.lo = (int)tree->level_start[d];
        c.hi = (int)tree->level_start[d + 1];
        long long nt = (long long)(c.hi - c.lo) * tree->n_range;
        if (nt == 0) continue;
        PRL_LAUNCH(reach_level_kernel<false>, grid_for(nt), s, c);
    }
    return prl::check(cudaGetLastError(), "prl_reach_pass");
}

extern "C" int prl_value_pass(const prl_tree_t* tree, const prl_buffers_t* buf, int player_mask, int with_br,
                              const int* strat_mode, prl_stream_t stream) {
    if (int e = check_tree(tree)) return e;
    if (with_br && !buf->ev_br) return prl::fail("prl_value_pass: with_br needs ev_br");
    Ctx c{*tree, *buf, 0, 0, player_mask, {strat_mode[0], strat_mode[1]}, 0, -1, 0, 0, 0};
    cudaStream_t s = (cudaStream_t)stream;
    for (int d = tree->n_levels - 1; d >= 0; --d) {
        c.lo = (int)tree->level_start[d];
        c.hi = (int)tree->level_start[d + 1];
        long long nt = (long long)(c.hi - c.lo) * tree->n_range;
        if (nt == 0) continue;
        if (with_br) PRL_LAUNCH((value_level_kernel<true, false>), grid_for(nt), s, c);
        else PRL_LAUNCH((value_level_kernel<false, false>), grid_for(nt), s, c);
    }
    return prl::check(cudaGetLastError(), "prl_value_pass");
}

extern "C" int prl_root_exploitability(const prl_tree_t* tree, const prl_buffers_t* buf, float* out_expl,
                                       prl_stream_t stream) {
    if (!buf->ev_br) return prl::fail("prl_root_exploitability needs ev_br");
    root_exploitability_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*tree, *buf, out_expl);
    prl::count_launch();
    return prl::check(cudaGetLastError(), "prl_root_exploitability");
}

extern "C" int prl_cfr_sweep(const prl_tree_t* tree, const prl_buffers_t* buf, int algo, int p, int iter, int delay,
                             int avg_f64, const int* strat_mode, int which, prl_stream_t stream) {
    if (int e = check_tree(tree)) return e;
    if (p < 0 || p > 1 || algo < 0 || algo > 2) return prl::fail("prl_cfr_sweep: bad p / algo");
    if (algo != PRL_ALGO_CFR_PLUS && avg_f64) return prl::fail("avg_f64 only applies to CFR+");
    Ctx c{*tree, *buf, 0, 0, 1 << p, {strat_mode[0], strat_mode[1]}, algo, p, iter, delay, avg_f64};
    cudaStream_t s = (cudaStream_t)stream;
    if (which & 1) {
        for (int d = tree->n_levels - 1; d >= 0; --d) {
            c.lo = (int)tree->level_start[d];
            c.hi = (int)tree->level_start[d + 1];
            long long nt = (long long)(c.hi - c.lo) * tree->n_range;
            if (nt == 0) continue;
            PRL_LAUNCH((value_level_kernel<false, true>), grid_for(nt), s, c);
        }
    }
    if (which & 2) {
        c.mode[p] = PRL_STRAT_F32;  // p's strategy now lives in the float table
        for (int d = 0; d < tree->n_levels; ++d) {
            c.lo = (int)tree->level_start[d];
            c.hi = (int)tree->level_start[d + 1];
            long long nt = (long long)(c.hi - c.lo) * tree->n_range;
            if (nt == 0) continue;
            PRL_LAUNCH(reach_level_kernel<true>, grid_for(nt), s, c);
        }
    }
    return prl::check(cudaGetLastError(), "prl_cfr_sweep");
}

extern "C" int prl_cfr_half_iteration(const prl_tree_t* tree, const prl_buffers_t* buf, int algo, int p, int iter,
                                      int delay, int avg_f64, const int* strat_mode, prl_stream_t stream) {
    return prl_cfr_sweep(tree, buf, algo, p, iter, delay, avg_f64, strat_mode, 3, stream);
}
