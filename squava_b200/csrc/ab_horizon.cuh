// ab_horizon.cuh -- the value of a HORIZON-1 node of the alpha-beta dialects in one bit-parallel shot.
//
// A node at ply maxDepth-1 has only depth-stop children: each child c is worth either a terminal score (the
// mover completes a 4-line / a 3-line through c) or the evaluator's heuristic
//     30 * (three-of-four quads through c) + scores[c]  (+ what the dialect carries down),
// alphabeta.go:121-163,165-223 / opening3.go:136-225.  The sequential search walks the lines through every
// child in turn; here the counts for ALL empty cells come out of shifts of the two masks at once:
//   * a WINDOW is four cells in a row; it is "good" for the mover when it holds exactly two of the mover's
//     stones and two empty cells -- then each of the two empty cells, once played, leaves a three-of-four quad
//     whose hole is the other one;
//   * the number of good windows through a cell is summed over the 4 directions x 4 offsets with bit-sliced
//     adders (one plane per binary digit, all 25 cells per instruction);
//   * the best child is found by walking the planes from the top digit down.
// The node's value is the true maximum over its children -- a legal return for any window of a fail-soft
// search -- so the root values do not change (every root value is the plain minimax value).
//
// Everything here is __host__ __device__: tests/test_ab_horizon_native.py compiles it for the host and compares
// it with a per-child table walk on random boards.
#pragma once
#include <cstdint>
#include "sq_device.cuh"

namespace sq {

#if defined(__CUDA_ARCH__)
#define SQ_POPC(x) __popc(x)
#define SQ_UNROLL _Pragma("unroll")
#else
#define SQ_POPC(x) __builtin_popcount(x)
#define SQ_UNROLL
#endif

constexpr int kHorizonMaxPlanes = 5;       // the bias table may span at most 29 (5 bits) for the digit walk to be exact

struct HorizonTables {
    uint32_t score_plane[kHorizonMaxPlanes];   // padded layout: bit p of plane k = digit k of (scores[cell] - score_min)
    int score_min;
    int n_planes;
    int ok;                                    // max - min <= 29: a larger three-of-four count always outranks the bias
};

inline void build_horizon_tables(const int8_t scores[25], HorizonTables& t) {
    int lo = 127, hi = -128;
    for (int c = 0; c < 25; c++) { lo = scores[c] < lo ? scores[c] : lo; hi = scores[c] > hi ? scores[c] : hi; }
    t.score_min = lo;
    t.ok = hi - lo <= 29;
    t.n_planes = 0;
    while ((1 << t.n_planes) <= hi - lo && t.n_planes < kHorizonMaxPlanes) t.n_planes++;
    for (int k = 0; k < kHorizonMaxPlanes; k++) {
        uint32_t m = 0;
        for (int c = 0; c < 25; c++) if (((scores[c] - lo) >> k) & 1) m |= 1u << c;
        t.score_plane[k] = pad(m);
    }
}

// Starts (padded layout, direction step d) of the windows that hold exactly two stones of m and no cell outside u.
__host__ __device__ __forceinline__ uint32_t good_windows(uint32_t m, uint32_t u, int d) {
    const uint32_t a1 = m >> d, a2 = m >> (2 * d), a3 = m >> (3 * d);
    const uint32_t inside = u & (u >> d) & (u >> (2 * d)) & (u >> (3 * d));
    const uint32_t even = ~(m ^ a1 ^ a2 ^ a3);
    const uint32_t any = m | a1 | a2 | a3, all = m & a1 & a2 & a3;
    return inside & even & any & ~all;
}

// Per-cell number of good windows through the cell, as four binary-digit planes (0..8).
struct CountPlanes { uint32_t b0, b1, b2, b3; };

__host__ __device__ __forceinline__ CountPlanes window_counts(uint32_t m, uint32_t u) {
    uint32_t lo[4], hi[4];
SQ_UNROLL
    for (int k = 0; k < 4; k++) {
        const int d = k == 0 ? 1 : k == 1 ? 6 : k == 2 ? 7 : 5;
        const uint32_t g = good_windows(m, u, d);
        // a cell lies in at most two windows of one direction: the four offsets add up to a 2-bit number
        const uint32_t s0 = g, s1 = g << d, s2 = g << (2 * d), s3 = g << (3 * d);
        const uint32_t x = s0 ^ s1 ^ s2, c = (s0 & s1) | (s2 & (s0 | s1));
        lo[k] = x ^ s3;
        hi[k] = c | (x & s3);
    }
    // (lo0,hi0) + (lo1,hi1) and (lo2,hi2) + (lo3,hi3): 3-bit sums, each <= 4
    const uint32_t p0 = lo[0] ^ lo[1], pc = lo[0] & lo[1];
    const uint32_t p1 = hi[0] ^ hi[1] ^ pc, p2 = (hi[0] & hi[1]) | (pc & (hi[0] | hi[1]));
    const uint32_t q0 = lo[2] ^ lo[3], qc = lo[2] & lo[3];
    const uint32_t q1 = hi[2] ^ hi[3] ^ qc, q2 = (hi[2] & hi[3]) | (qc & (hi[2] | hi[3]));
    CountPlanes r;
    r.b0 = p0 ^ q0;
    const uint32_t c0 = p0 & q0;
    r.b1 = p1 ^ q1 ^ c0;
    const uint32_t c1 = (p1 & q1) | (c0 & (p1 | q1));
    r.b2 = p2 ^ q2 ^ c1;
    r.b3 = (p2 & q2) | (c1 & (p2 | q2));
    return r;
}

// The same for src/alphabeta's DELAYED evaluation: a child marks its own first empty cell j0 before it evaluates, so a
// window whose other empty cell is j0 does not count.  For every child but the one ON the first empty cell f0, j0 = f0:
// windows through f0 are dropped.  For the child on f0, j0 is the second empty cell f1: its count is the number of good
// windows through f0 that do not hold f1 (first_cnt).
__host__ __device__ __forceinline__ CountPlanes window_counts_delayed(uint32_t m, uint32_t u, uint32_t f0, uint32_t f1, int& first_cnt) {
    uint32_t lo[4], hi[4];
    int fc = 0;
SQ_UNROLL
    for (int k = 0; k < 4; k++) {
        const int d = k == 0 ? 1 : k == 1 ? 6 : k == 2 ? 7 : 5;
        uint32_t g = good_windows(m, u, d);
        const uint32_t w0 = f0 | (f0 >> d) | (f0 >> (2 * d)) | (f0 >> (3 * d));     // starts of the windows that hold f0
        const uint32_t w1 = f1 | (f1 >> d) | (f1 >> (2 * d)) | (f1 >> (3 * d));
        fc += SQ_POPC(g & w0 & ~w1);
        g &= ~w0;
        const uint32_t s0 = g, s1 = g << d, s2 = g << (2 * d), s3 = g << (3 * d);
        const uint32_t x = s0 ^ s1 ^ s2, c = (s0 & s1) | (s2 & (s0 | s1));
        lo[k] = x ^ s3;
        hi[k] = c | (x & s3);
    }
    first_cnt = fc;
    const uint32_t p0 = lo[0] ^ lo[1], pc = lo[0] & lo[1];
    const uint32_t p1 = hi[0] ^ hi[1] ^ pc, p2 = (hi[0] & hi[1]) | (pc & (hi[0] | hi[1]));
    const uint32_t q0 = lo[2] ^ lo[3], qc = lo[2] & lo[3];
    const uint32_t q1 = hi[2] ^ hi[3] ^ qc, q2 = (hi[2] & hi[3]) | (qc & (hi[2] | hi[3]));
    CountPlanes r;
    r.b0 = p0 ^ q0;
    const uint32_t c0 = p0 & q0;
    r.b1 = p1 ^ q1 ^ c0;
    const uint32_t c1 = (p1 & q1) | (c0 & (p1 | q1));
    r.b2 = p2 ^ q2 ^ c1;
    r.b3 = (p2 & q2) | (c1 & (p2 | q2));
    return r;
}

// planes += f1 + f2 (two more one-bit inputs per cell; the sum stays below 16)
__host__ __device__ __forceinline__ void add_two_bits(CountPlanes& p, uint32_t f1, uint32_t f2) {
    const uint32_t s0 = p.b0 ^ f1 ^ f2, c0 = (p.b0 & f1) | (f2 & (p.b0 | f1));
    const uint32_t s1 = p.b1 ^ c0, c1 = p.b1 & c0;
    const uint32_t s2 = p.b2 ^ c1, c2 = p.b2 & c1;
    p.b0 = s0; p.b1 = s1; p.b2 = s2; p.b3 ^= c2;
}

// max over the cells of cand (non-empty) of 30 * count(cell) + scores[cell]; exact because the bias spans < 30
__host__ __device__ __forceinline__ int best_of(const CountPlanes& p, uint32_t cand, const HorizonTables& H) {
    int t = 0;
    uint32_t s;
    s = cand & p.b3; if (s) { cand = s; t += 8; }
    s = cand & p.b2; if (s) { cand = s; t += 4; }
    s = cand & p.b1; if (s) { cand = s; t += 2; }
    s = cand & p.b0; if (s) { cand = s; t += 1; }
    int v = 30 * t + H.score_min;
    for (int k = H.n_planes - 1; k >= 0; k--) {
        s = cand & H.score_plane[k];
        if (s) { cand = s; v += 1 << k; }
    }
    return v;
}

// Exact move classes of the mover `m` (padded) over the empty cells `e` (padded): win = cells that complete a
// 4-line through themselves, sui = cells that complete a 3-line but no 4-line (same as ab_move_classes).
__host__ __device__ __forceinline__ void move_classes_padded(uint32_t m, uint32_t e, uint32_t& win, uint32_t& sui) {
    uint32_t w = 0, l = 0;
SQ_UNROLL
    for (int k = 0; k < 4; k++) {
        const int d = k == 0 ? 1 : k == 1 ? 6 : k == 2 ? 7 : 5;
        const uint32_t a1 = m << d, a2 = m << (2 * d), a3 = m << (3 * d);
        const uint32_t b1 = m >> d, b2 = m >> (2 * d), b3 = m >> (3 * d);
        l |= (a1 & (a2 | b1)) | (b1 & b2);
        w |= (a1 & a2 & (a3 | b1)) | (b1 & b2 & (a1 | b3));
    }
    win = w & e;
    sui = l & e & ~win;
}

// Value of a horizon-1 node RELATIVE TO ITS MOVER (positive is good for the side to move there).
//   own, opp, emp   API-layout masks of the mover, the other side and the empty cells; popc(emp) >= 2
//   term            10000 - (ply of the children): what completing a 4-line is worth there
//   konst           what every heuristic child's value carries besides 30*count + scores[cell]
//   DELAYED (src/alphabeta, src/abbook): the child marks its own first empty cell j0 before it evaluates the move
//   that led to it (alphabeta.go:173-177), so a quad whose hole is j0 does not count, and the pending move of THIS
//   node -- walk masks pe1 / pe2, API layout -- loses one (two) counted quads when the child stands on a pe1
//   (pe2) cell; both are folded into the per-cell count.  first_score: scores[] of the FIRST empty cell (the one child
//   whose j0 is the second empty cell; its count comes out of window_counts_delayed).
template <bool DELAYED>
__host__ __device__ __forceinline__ int horizon_node_value(uint32_t own, uint32_t opp, uint32_t emp, int term, int konst,
                                                           uint32_t pe1, uint32_t pe2, int first_score, const HorizonTables& H) {
    (void)opp;
    const uint32_t m = pad(own), e = pad(emp);
    uint32_t win, sui;
    move_classes_padded(m, e, win, sui);
    if (win) return term;
    int best = sui ? -term : -0x40000000;
    uint32_t cand = e & ~sui;
    if (DELAYED) {
        const uint32_t f0 = e & (0u - e), rest = e ^ f0, f1 = rest & (0u - rest);
        const uint32_t f1p = pad(pe1), f2p = pad(pe2);
        int first_cnt;
        CountPlanes p = window_counts_delayed(m, m | e, f0, f1, first_cnt);
        if (cand & f0) {
            const int t0 = first_cnt + ((f1p & f0) ? 1 : 0) + ((f2p & f0) ? 1 : 0);
            const int v0 = 30 * t0 + first_score + konst;
            best = v0 > best ? v0 : best;
        }
        cand &= ~f0;
        if (cand) {
            add_two_bits(p, f1p, f2p);
            const int v = best_of(p, cand, H) + konst;
            best = v > best ? v : best;
        }
    } else if (cand) {
        const CountPlanes p = window_counts(m, m | e);
        const int v = best_of(p, cand, H) + konst;
        best = v > best ? v : best;
    }
    return best;
}

}  // namespace sq
