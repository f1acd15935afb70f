/*
 * CPU restatement of the reference's 7-card Hold'em hand evaluator and Hold'em LUT generators.
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline) - never linked into the product library.
 *
 * The reference ships these as binaries without source: PokerRL/game/_/cpp_wrappers/lib_hand_eval.so (wrapper
 * CppHandeval.py:19-65; call sites game_rules.py:213-223, 296-306) and lib_luts.so (CppLUT.py:14-94).  The algorithm
 * below is therefore a black-box restatement: the int32 strength encoding was recovered from the immediates in the
 * binary's rank_seven_card_2d_hand (0x8ca0b pair, 0xa0c4c two pair, 0xa17c6 trips, 0xa2340 straight, 0xa234e flush,
 * 0x12ed59 full house, 0x12ee2a quads, 0x12eefb straight flush) and by probing.
 *
 * Parity status: PINNED against outputs of the reference binary itself - oracle/gen_golden_holdem.py compares this
 * file with lib_hand_eval.so on 3 million random 7-card hands + every category-targeted family in the build container
 * (bit-exact) and commits a fixture of boards x 1326 hand ranks under tests/golden/.
 *
 * Encoding (ranks 0..12 = 2..A, higher value = stronger hand):
 *   high card        r1*13^4 + r2*13^3 + r3*13^2 + r4*13 + r5
 *   pair             576011  + 2197*p + 169*k1 + 13*k2 + k3
 *   two pair         658508  + 169*hp + 13*lp + k
 *   three of a kind  661446  + 169*t + 13*k1 + k2
 *   straight         664384  + top                       (wheel: top = 3)
 *   flush            664398  + high-card value of the five best suited cards - (value of 5,3,2,1,0 ... see FLUSH_BASE)
 *   full house       1240409 + 13*t + p
 *   four of a kind   1240618 + 13*q + k    QUIRK: k is the card right above the quads in descending order if any,
 *                                          else the best card below (the binary takes the first sorted 5-card window
 *                                          that contains the quads) - SURVEY.md §8(a) row H
 *   straight flush   1240827 + top
 */
#include <stdint.h>
#include <string.h>

#define PAIR_BASE 576011
#define TWO_PAIR_BASE 658508
#define TRIPS_BASE 661446
#define STRAIGHT_BASE 664384
#define FLUSH_BASE 664398
#define FULL_HOUSE_BASE 1240409
#define QUADS_BASE 1240618
#define STRAIGHT_FLUSH_BASE 1240827

/* highest straight top rank in a 13-bit rank mask, or -1 (wheel A-5 has top 3) */
static int straight_top(unsigned mask) {
    for (int top = 12; top >= 4; --top)
        if (((mask >> (top - 4)) & 0x1F) == 0x1F) return top;
    if ((mask & 0x100F) == 0x100F) return 3;
    return -1;
}

/* value of the five highest ranks of a mask, base 13 */
static int top5_value(unsigned mask) {
    int v = 0, n = 0;
    for (int r = 12; r >= 0 && n < 5; --r)
        if (mask & (1u << r)) { v = v * 13 + r; ++n; }
    return v;
}

/* cards: 7 card ids 0..51, c = rank*4 + suit */
int32_t orc_rank7(const int8_t* cards) {
    int cnt[13] = {0};
    unsigned suit_mask[4] = {0, 0, 0, 0}, all = 0;
    int suit_cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < 7; ++i) {
        const int r = cards[i] >> 2, s = cards[i] & 3;
        ++cnt[r];
        ++suit_cnt[s];
        suit_mask[s] |= 1u << r;
        all |= 1u << r;
    }
    /* flush / straight flush */
    for (int s = 0; s < 4; ++s) {
        if (suit_cnt[s] >= 5) {
            const int st = straight_top(suit_mask[s]);
            if (st >= 0) return STRAIGHT_FLUSH_BASE + st;
        }
    }
    int quad = -1, trip1 = -1, trip2 = -1, pair1 = -1, pair2 = -1, pair3 = -1;
    for (int r = 12; r >= 0; --r) {
        if (cnt[r] == 4) quad = r;
        else if (cnt[r] == 3) { if (trip1 < 0) trip1 = r; else if (trip2 < 0) trip2 = r; }
        else if (cnt[r] == 2) { if (pair1 < 0) pair1 = r; else if (pair2 < 0) pair2 = r; else if (pair3 < 0) pair3 = r; }
    }
    if (quad >= 0) {
        int k = -1;
        for (int r = quad + 1; r <= 12; ++r)
            if (cnt[r]) { k = r; break; } /* lowest rank above the quads */
        if (k < 0)
            for (int r = quad - 1; r >= 0; --r)
                if (cnt[r]) { k = r; break; } /* else the best card below */
        return QUADS_BASE + 13 * quad + k;
    }
    if (trip1 >= 0 && (trip2 >= 0 || pair1 >= 0)) {
        const int p = (trip2 > pair1) ? trip2 : pair1;
        return FULL_HOUSE_BASE + 13 * trip1 + p;
    }
    for (int s = 0; s < 4; ++s)
        if (suit_cnt[s] >= 5) return FLUSH_BASE + top5_value(suit_mask[s]);
    {
        const int st = straight_top(all);
        if (st >= 0) return STRAIGHT_BASE + st;
    }
    if (trip1 >= 0) {
        int k[2], n = 0;
        for (int r = 12; r >= 0 && n < 2; --r)
            if (cnt[r] && r != trip1) k[n++] = r;
        return TRIPS_BASE + 169 * trip1 + 13 * k[0] + k[1];
    }
    if (pair2 >= 0) {
        int k = -1;
        for (int r = 12; r >= 0; --r)
            if (cnt[r] && r != pair1 && r != pair2) { k = r; break; }
        return TWO_PAIR_BASE + 169 * pair1 + 13 * pair2 + k;
    }
    if (pair1 >= 0) {
        int k[3], n = 0;
        for (int r = 12; r >= 0 && n < 3; --r)
            if (cnt[r] && r != pair1) k[n++] = r;
        return PAIR_BASE + 2197 * pair1 + 169 * k[0] + 13 * k[1] + k[2];
    }
    return top5_value(all);
}

/* CppHandeval.get_hand_rank_all_hands_on_given_boards_52_holdem (CppHandeval.py:45-65): out[n][1326], -1 where the
 * hand shares a card with the board; hands in LUT order (c1 < c2 lexicographic). */
void orc_rank_boards(int32_t* out, const int8_t* boards, int32_t n_boards) {
    for (int b = 0; b < n_boards; ++b) {
        const int8_t* bd = boards + 5 * b;
        uint64_t bm = 0;
        for (int i = 0; i < 5; ++i) bm |= 1ull << bd[i];
        int idx = 0;
        for (int c1 = 0; c1 < 52; ++c1) {
            for (int c2 = c1 + 1; c2 < 52; ++c2, ++idx) {
                if ((bm >> c1) & 1 || (bm >> c2) & 1) { out[(size_t)b * 1326 + idx] = -1; continue; }
                int8_t cards[7] = {(int8_t)c1, (int8_t)c2, bd[0], bd[1], bd[2], bd[3], bd[4]};
                out[(size_t)b * 1326 + idx] = orc_rank7(cards);
            }
        }
    }
}

/* batch of n 7-card hands, cards[n][7] -> out[n] */
void orc_rank7_batch(int32_t* out, const int8_t* cards, int32_t n) {
    for (int i = 0; i < n; ++i) out[i] = orc_rank7(cards + 7 * i);
}
