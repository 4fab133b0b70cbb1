// negascout.cuh -- src/negascout's ChooseMove (negascout.go:81-122) as a GPU batch, one WARP
// per position.
//
// Unlike the alpha-beta dialects this search is not a function of the position alone: null
// windows are accepted two plies above the horizon without a re-search (negascout.go:248) and
// the root hands its RUNNING alpha to movekeeper (:104-109), so every value depends on the exact
// order and windows of the reference's recursion.  The GPU therefore replays that recursion
// step for step; what it parallelises is
//   * the whole-board evaluator staticValue (:168-230): the 28 quads, 48 triplets and 4
//     "deadly" diagonals are spread over the lanes of the warp (one popc pair per lane, one
//     redux.add and a few ballots per node instead of ~330 table loads);
//   * the move generator: lane i tests the i-th cell of the position's ordered move list, one
//     ballot finds the next legal move;
//   * positions: warps pull positions from a queue, so a batch fills the machine.
// The search state (window, score, iterator) is warp-uniform and lives in registers; parents'
// frames sit in shared memory, one copy per warp (every access is a broadcast).
//
// Scan order matters in one corner: staticValue returns at the FIRST complete line of its scan
// (checkable cells in order, each cell's lines in table order).  Lanes hold the lines in that
// first-occurrence order, so the lowest set bit of a ballot IS the reference's first hit.
#pragma once
#include <cstdint>
#include <cstring>
#include "../../include/squava_b200.h"
#include "sq_device.cuh"
#include "sq_tables.h"

namespace sq {

constexpr int kNsWarps = 4;            // warps (positions in flight) per block
constexpr int kNsMaxDepth = 24;
constexpr uint32_t kNsDummy = 0xC0000000u;   // bits no board has: never complete, never three-of-four

struct NsTables {
    uint32_t quad[32];        // first-occurrence scan order (negascout.go:172-186); 28 used
    int32_t quad_weight[32];  // 10 * (number of checkable cells on the quad): the scan meets it that often
    uint32_t trip[64];        // first-occurrence scan order (:188-199); 48 used
    uint32_t deadly_inner[4], deadly_outer[4];   // :147-152
    uint8_t initial_order[32];                    // :392-399
    uint8_t n_cell_quads[25], n_cell_trips[25];
    uint32_t cell_quads[25][8];                   // indexedWinningQuads[cell], table order (:38-51)
    uint32_t cell_trips[25][12];
};
__constant__ NsTables c_ns;

inline void build_ns_tables(NsTables& t) {
    memset(&t, 0, sizeof t);
    ScanTables s;
    build_scan_tables(s);
    const int checkable[9] = {2, 7, 10, 11, 12, 13, 14, 17, 22};            // negascout.go:159-163
    for (int c = 0; c < 25; c++) {
        for (int i = 0; i < 28; i++) if (s.ab_quads[i] >> c & 1u) t.cell_quads[c][t.n_cell_quads[c]++] = s.ab_quads[i];
        for (int i = 0; i < 48; i++) if (s.ab_triplets[i] >> c & 1u) t.cell_trips[c][t.n_cell_trips[c]++] = s.ab_triplets[i];
    }
    int nq = 0, nt = 0;
    for (int i = 0; i < 32; i++) t.quad[i] = kNsDummy;
    for (int i = 0; i < 64; i++) t.trip[i] = kNsDummy;
    for (int c : checkable) {
        for (int k = 0; k < t.n_cell_quads[c]; k++) {
            const uint32_t m = t.cell_quads[c][k];
            int at = -1;
            for (int j = 0; j < nq; j++) if (t.quad[j] == m) at = j;
            if (at < 0) { at = nq++; t.quad[at] = m; }
            t.quad_weight[at] += 10;
        }
        for (int k = 0; k < t.n_cell_trips[c]; k++) {
            const uint32_t m = t.cell_trips[c][k];
            bool seen = false;
            for (int j = 0; j < nt; j++) seen |= t.trip[j] == m;
            if (!seen) t.trip[nt++] = m;
        }
    }
    const int deadly[4][4] = {{5, 11, 17, 23}, {21, 17, 13, 9}, {1, 7, 13, 19}, {15, 11, 7, 3}};
    for (int d = 0; d < 4; d++) {
        t.deadly_outer[d] = (1u << deadly[d][0]) | (1u << deadly[d][3]);
        t.deadly_inner[d] = (1u << deadly[d][1]) | (1u << deadly[d][2]);
    }
    const uint8_t init[25] = {6, 8, 18, 16, 1, 3, 9, 19, 23, 21, 15, 5, 0, 4, 24, 20, 12, 7, 13, 17, 11, 2, 10, 14, 22};
    memcpy(t.initial_order, init, 25);
}

struct NsScores { int8_t s[25]; };

struct NsFrame { int alpha, beta, score, n, k, cell_phase; };

// outputs per position: values[25], visit_order[25], best_value, best_mask, first_best, nodes
struct NsOut {
    int32_t (*values)[25];
    int8_t (*visit_order)[25];
    int32_t* best_value;
    uint32_t* best_mask;
    int8_t* first_best;
    unsigned long long* nodes;
};

// reorderMoves (negascout.go:411-499): lane i classifies the i-th cell of the initial order as
// good / bad / dull; the two ordered lists are stable partitions of that order.
__device__ __forceinline__ void ns_reorder(uint32_t x, uint32_t o, int lane, uint8_t (*order)[32]) {
    const uint32_t full = 0xffffffffu;
    const int c = lane < 25 ? c_ns.initial_order[lane] : 0;
    const bool empty = lane < 25 && !(((x | o) >> c) & 1u);
    int kind = 0;                                             // +1 good, -1 bad, 0 dull
    if (empty) {
        for (int k = 0; k < c_ns.n_cell_quads[c] && kind == 0; k++) {
            const uint32_t m = c_ns.cell_quads[c][k];
            const int sum = __popc(m & x) - __popc(m & o);
            if (sum == 3 || sum == 2) kind = 1; else if (sum == -3 || sum == -2) kind = -1;
        }
        for (int k = 0; k < c_ns.n_cell_trips[c] && kind == 0; k++) {
            const uint32_t m = c_ns.cell_trips[c][k];
            const int sum = __popc(m & x) - __popc(m & o);
            if (sum == -2) kind = 1; else if (sum == 2) kind = -1;
        }
    }
    const uint32_t good = __ballot_sync(full, empty && kind > 0), bad = __ballot_sync(full, empty && kind < 0);
    const uint32_t dull = __ballot_sync(full, empty && kind == 0);
    const uint32_t below = (1u << lane) - 1u;
    const int ng = __popc(good), nb = __popc(bad), nd = __popc(dull);
    if (empty) {
        const int within = __popc((kind > 0 ? good : kind < 0 ? bad : dull) & below);
        // MAXIMIZER: good, dull, bad.  MINIMIZER: bad, dull, good.
        order[1][kind > 0 ? within : kind == 0 ? ng + within : ng + nd + within] = (uint8_t)c;
        order[0][kind < 0 ? within : kind == 0 ? nb + within : nb + nd + within] = (uint8_t)c;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kNsWarps * 32)
negascout_kernel(const sq_board* __restrict__ pos, const int* __restrict__ idx, int n_pos, int max_depth, NsScores sc,
                 NsOut out, unsigned int* __restrict__ queue) {
    __shared__ NsFrame s_frames[kNsWarps][kNsMaxDepth + 2];
    __shared__ uint8_t s_order[kNsWarps][2][32];       // [0] MINIMIZER's list, [1] MAXIMIZER's
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    NsFrame* const frames = s_frames[warp];
    const uint32_t full = 0xffffffffu;

    // this lane's share of the evaluator
    const uint32_t my_quad = c_ns.quad[lane];
    const int my_weight = c_ns.quad_weight[lane] * 3;                 // sum * 10 with sum = +-3
    const uint32_t my_t0 = c_ns.trip[lane], my_t1 = c_ns.trip[lane + 32];
    const uint32_t my_in = lane >= 28 ? c_ns.deadly_inner[lane - 28] : 0u;
    const uint32_t my_out = lane >= 28 ? c_ns.deadly_outer[lane - 28] : 0u;
    int my_score = 0;                                                  // lane c keeps scores[c]
#pragma unroll
    for (int c = 0; c < 25; c++) if (lane == c) my_score = sc.s[c];

    for (;;) {
        unsigned int p = 0;
        if (lane == 0) p = atomicAdd(queue, 1u);
        p = __shfl_sync(full, p, 0);
        if (p >= (unsigned)n_pos) break;
        p = idx[p];
        uint32_t x = pos[p].x & kFull25, o = pos[p].o & kFull25;
        unsigned long long nodes = 0;

        ns_reorder(x, o, lane, s_order[warp]);
        const int n_moves = __popc(kFull25 & ~(x | o));
        int bias = __reduce_add_sync(full, lane < 25 ? (int)((x >> lane) & 1u) * my_score - (int)((o >> lane) & 1u) * my_score : 0);

        // ---- the root loop of ChooseMove (:88-116) and negaScout (:232-266), one state machine
        int ply = 0;                                   // ply of the node whose loop is running
        int alpha = -2 * 10000, beta = 2 * 10000, score = -3 * 10000, n = beta, k = 0, cell = 0, phase = 0;
        int keeper_max = -3 * 10000, first_best = -1, n_visited = 0;
        uint32_t keeper_mask = 0;
        if (lane < 25) { out.values[p][lane] = SQ_NO_MOVE; out.visit_order[p][lane] = -1; }
        int ret = 0, rs_cur = 0;
        bool returned = false;                         // `ret` is a child's value for the running loop
        bool enter = false;                            // the child on `cell` is to be evaluated (phase says which window)
        for (;;) {
            const int player = (ply & 1) ? -1 : 1;
            if (!returned && !enter) {
                // next legal move of the running loop, in this player's order
                const int c = lane < n_moves ? s_order[warp][player > 0 ? 1 : 0][lane] : 0;
                const uint32_t cand = __ballot_sync(full, lane < n_moves && lane >= k && !(((x | o) >> c) & 1u));
                if (cand == 0) {                       // loop over: hand the score up
                    if (ply == 0) break;
                    ret = score;
                    ply--;
                    const NsFrame f = frames[ply];
                    alpha = f.alpha; beta = f.beta; score = f.score; n = f.n; k = f.k;
                    cell = f.cell_phase & 31; phase = f.cell_phase >> 5;
                    returned = true;
                    continue;
                }
                const int at = __ffs(cand) - 1;
                k = at + 1;
                cell = __shfl_sync(full, c, at);
                phase = 1;
                enter = true;
                if (player > 0) x |= 1u << cell; else o |= 1u << cell;
                bias += player * __shfl_sync(full, my_score, cell);
            }
            if (enter) {
                // the child at ply+1, window (-n, -alpha) [phase 1] or (-beta, -cur) [phase 2]
                const int ca = phase == 1 ? -n : -beta, cb = phase == 1 ? -alpha : -rs_cur;
                const int cply = ply + 1;
                enter = false;
                nodes++;
                // staticValue(cply), negascout.go:168-230
                const uint32_t fx = __ballot_sync(full, (my_quad & ~x) == 0u), fo = __ballot_sync(full, (my_quad & ~o) == 0u);
                bool stop = false; int bv = 0;
                if (fx | fo) {
                    const int first = __ffs(fx | fo) - 1;
                    bv = ((fx >> first) & 1u) ? (10000 - cply) : -(10000 - cply);
                    stop = true;
                } else {
                    const uint32_t tx0 = __ballot_sync(full, (my_t0 & ~x) == 0u), to0 = __ballot_sync(full, (my_t0 & ~o) == 0u);
                    const uint32_t tx1 = __ballot_sync(full, (my_t1 & ~x) == 0u), to1 = __ballot_sync(full, (my_t1 & ~o) == 0u);
                    if (tx0 | to0 | tx1 | to1) {
                        bool owner_x;
                        if (tx0 | to0) owner_x = (tx0 >> (__ffs(tx0 | to0) - 1)) & 1u;
                        else owner_x = (tx1 >> (__ffs(tx1 | to1) - 1)) & 1u;
                        bv = owner_x ? -(10000 - cply) : (10000 - cply);      // -sum/3 * (WIN - ply)
                        stop = true;
                    } else {
                        const int cx = __popc(my_quad & x), co = __popc(my_quad & o);
                        int contrib = (cx == 3 && co == 0) ? my_weight : (co == 3 && cx == 0) ? -my_weight : 0;
                        if (lane >= 28) {
                            const int inner = __popc(my_in & x) - __popc(my_in & o);
                            const int outer = __popc(my_out & x) - __popc(my_out & o);
                            if (outer == 0 && (inner == 2 || inner == -2)) contrib -= inner / 2 * 5;
                        }
                        bv = __reduce_add_sync(full, contrib);
                        if (bv == 0) bv = bias;
                        stop = cply > max_depth;
                    }
                }
                if (stop) {
                    ret = -player * bv;                // the child's player is -player
                    returned = true;
                } else {
                    // the child has a loop of its own: push the running frame
                    NsFrame f;
                    f.alpha = alpha; f.beta = beta; f.score = score; f.n = n; f.k = k; f.cell_phase = cell | (phase << 5);
                    if (lane == 0) frames[ply] = f;
                    __syncwarp();
                    ply = cply;
                    alpha = ca; beta = cb; score = -3 * 10000; n = cb; k = 0;
                    continue;
                }
            }
            // `ret` came back to the running loop (:243-262)
            returned = false;
            if (phase == 1) {
                const int cur = -ret;
                if (cur > score) {
                    if (n == beta || (ply > 0 && ply == max_depth - 2)) score = cur;
                    else { phase = 2; rs_cur = cur; enter = true; continue; }     // re-search with (-beta, -cur)
                }
            } else {
                score = -ret;
            }
            if (player > 0) x &= ~(1u << cell); else o &= ~(1u << cell);
            bias -= player * __shfl_sync(full, my_score, cell);
            if (score > alpha) alpha = score;
            if (ply == 0) {                            // moves.SetMove(i, j, alpha), movekeeper.go:26-36
                if (lane == 0) { out.values[p][cell] = alpha; out.visit_order[p][n_visited] = (int8_t)cell; }
                n_visited++;
                if (alpha >= keeper_max) {
                    if (alpha > keeper_max) { keeper_max = alpha; keeper_mask = 0; first_best = cell; }
                    keeper_mask |= 1u << cell;
                }
            }
            if (alpha >= beta) {                       // cut: this loop is over
                if (ply == 0) break;
                ret = score;
                ply--;
                const NsFrame f = frames[ply];
                alpha = f.alpha; beta = f.beta; score = f.score; n = f.n; k = f.k;
                cell = f.cell_phase & 31; phase = f.cell_phase >> 5;
                returned = true;
                continue;
            }
            n = alpha + 1;
        }
        if (lane == 0) {
            out.best_value[p] = keeper_mask ? keeper_max : 0;        // movekeeper.go:40-44
            out.best_mask[p] = keeper_mask;
            out.first_best[p] = (int8_t)first_best;
            out.nodes[p] = nodes;
        }
        __syncwarp();
    }
}

// ---- the fast path ------------------------------------------------------------------------
// Same recursion, same windows, same counts -- but no node is ever evaluated from scratch.  When a
// node becomes the running loop, ONE lane-parallel pass evaluates all of its children at once:
// lane i takes the i-th cell of the mover's ordered list and derives the child's staticValue from
// the parent's (a position whose loop runs has no complete line, so only lines THROUGH the new
// mark can complete or change their three-of-four state; the "deadly" diagonals through the cell
// are recomputed).  Visiting a child is then a shared-memory read.
// At the horizon (children of a node at ply == maxDepth all stop) even the loop disappears: the
// running maximum, the cut-off point and the number of re-searches of negascout.go:240-262 are
// prefix-maximum scans over the lanes.
// Needs: root without a complete line, maxDepth >= 1 (otherwise negascout_kernel runs).
struct NsCellTables {          // [k][cell], padded: conflict-free when lanes hold different cells
    uint32_t quads[8][32];
    int32_t quad_w[8][32];     // 30 * times the scan meets the quad
    uint32_t trips[12][32];
    uint32_t deadly_in[2][32], deadly_out[2][32];
};
__constant__ NsCellTables c_ns_cell;

inline void build_ns_cell_tables(const NsTables& t, NsCellTables& c) {
    memset(&c, 0, sizeof c);
    for (int cell = 0; cell < 32; cell++) {
        for (int k = 0; k < 8; k++) {
            c.quads[k][cell] = (cell < 25 && k < t.n_cell_quads[cell]) ? t.cell_quads[cell][k] : kNsDummy;
            for (int j = 0; j < 28; j++) if (t.quad[j] == c.quads[k][cell]) c.quad_w[k][cell] = 3 * t.quad_weight[j];
        }
        for (int k = 0; k < 12; k++) c.trips[k][cell] = (cell < 25 && k < t.n_cell_trips[cell]) ? t.cell_trips[cell][k] : kNsDummy;
        int nd = 0;
        for (int d = 0; d < 4 && cell < 25; d++)
            if ((t.deadly_inner[d] | t.deadly_outer[d]) >> cell & 1u) {
                c.deadly_in[nd][cell] = t.deadly_inner[d]; c.deadly_out[nd][cell] = t.deadly_outer[d]; nd++;
            }
    }
}

// No popcounts below: POPC issues at a quarter of the ALU rate on its own pipe (8 cycles per warp instruction), and one pass
// over a node's children used to need ~60 of them.  Every count that matters here is "all of these two / three cells", which
// is a mask compare.
__device__ __forceinline__ int ns_deadly_term(uint32_t in, uint32_t out, uint32_t x, uint32_t o) {
    // inner cells both one player's, outer cells balanced (both empty or one each): -5 for X's pair, +5 for O's
    const bool ix = in != 0u && (in & x) == in, io = in != 0u && (in & o) == in;
    const bool balanced = ((out & x) == 0u) == ((out & o) == 0u);
    return balanced ? (ix ? -5 : io ? 5 : 0) : 0;
}

struct NsChild { bool term; int cur; int sum; };   // cur: the child's value as the parent's loop sees it (-ret)

// staticValue of (x, o) + mover `pl` on `cell`, from the parent's raw sum `vsum` and bias; cply = child's ply
__device__ __forceinline__ NsChild ns_child(const NsCellTables& C, uint32_t x, uint32_t o, int pl, int cell, int vsum,
                                            int bias_child, int cply) {
    const uint32_t own = pl > 0 ? x : o, opp = pl > 0 ? o : x;
    const uint32_t bit = 1u << cell;
    bool four = false, three = false;
    int gain = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t q = C.quads[k][cell];
        const uint32_t m = q & ~own;               // cells of the quad the mover does not hold: the new one among them
        four |= m == bit;                          // ... and nothing else: four in a row
        const uint32_t r = m & ~bit;               // the others
        // a new own three: exactly one other cell missing and it is empty; or an opposing three blocked: the other three are all opposing
        const bool own_three = r != 0u && (r & ((r - 1u) | opp)) == 0u;
        const bool blocked = (q & ~opp) == bit;
        gain += (own_three || blocked) ? C.quad_w[k][cell] : 0;
    }
#pragma unroll
    for (int k = 0; k < 12; k++) three |= (C.trips[k][cell] & ~own) == bit;
    const uint32_t x2 = x | (pl > 0 ? bit : 0u), o2 = o | (pl > 0 ? 0u : bit);
    int dd = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t in = C.deadly_in[j][cell], out = C.deadly_out[j][cell];
        dd += ns_deadly_term(in, out, x2, o2) - ns_deadly_term(in, out, x, o);
    }
    NsChild r;
    r.sum = vsum + pl * gain + dd;
    r.term = four || three;
    const int bv = four ? pl * (10000 - cply) : three ? -pl * (10000 - cply) : (r.sum == 0 ? bias_child : r.sum);
    r.cur = pl * bv;
    return r;
}

struct NsFastFrame { int alpha, beta, score, n, k, at_phase; uint32_t active, term; };

// CALL ITEMS.  Every call of negaScout (negascout.go:232-266) is a function of the node and its window (alpha, beta) alone.
// One item = one such call on the node reached from a position by `path` (the root's mover first): the kernel searches
// it exactly as the whole-position kernel would (same order, same windows, same node count) and hands back the value
// and the number of staticValue calls below it (its own included).  The host replays the loops ABOVE the items
// (abi.cu: NsSplit): the first child of a loop alone, then ALL remaining children at once under the alpha they would get
// if nothing improves; a result is used only if the loop really makes that call, so values, visiting order,
// movekeeper set and leaf counter stay bit-identical while the searches of one position spread over many warps.
struct NsItem {
    int32_t pos;        // position index
    int32_t a, b;       // the window the node is called with
    uint8_t n_path;     // >= 1: the node's ply
    uint8_t path[7];    // cells played from the position, the position's mover (X) first; no node before the last is terminal
    uint32_t budget;    // give up after this many staticValue calls (0: never): the host then runs the loop itself
    uint8_t deepest;    // ... at a node of at most this ply
    uint8_t pad[3];
};
// An item that gives up says how far it got: below its node it followed `n_chain` FIRST children (no loop on that way had
// finished a child yet) to a loop that had finished `i_done` children -- that loop's alpha, score and the staticValue calls
// under its finished children come back, so the host resumes exactly there and only the child in progress is searched again.
struct NsItemOut {
    int32_t ret, score;
    unsigned long long nodes;       // kNsAborted: gave up
    int32_t alpha;
    uint32_t nodes_done;
    uint8_t n_chain, i_done, cells[6];
};
constexpr unsigned long long kNsAborted = ~0ull;      // NsOut.nodes of a position that ran out of its node budget

// reorderMoves on the host: the driver needs the children of a node in the mover's order
inline int ns_reorder_host(const NsTables& t, uint32_t x, uint32_t o, uint8_t order_max[25], uint8_t order_min[25]) {
    uint8_t good[25], dull[25], bad[25];
    int ng = 0, nd = 0, nb = 0;
    for (int i = 0; i < 25; i++) {
        const int c = t.initial_order[i];
        if (((x | o) >> c) & 1u) continue;
        int kind = 0;
        for (int k = 0; k < t.n_cell_quads[c] && kind == 0; k++) {
            const uint32_t m = t.cell_quads[c][k];
            const int sum = __builtin_popcount(m & x) - __builtin_popcount(m & o);
            if (sum == 3 || sum == 2) kind = 1; else if (sum == -3 || sum == -2) kind = -1;
        }
        for (int k = 0; k < t.n_cell_trips[c] && kind == 0; k++) {
            const uint32_t m = t.cell_trips[c][k];
            const int sum = __builtin_popcount(m & x) - __builtin_popcount(m & o);
            if (sum == -2) kind = 1; else if (sum == 2) kind = -1;
        }
        if (kind > 0) good[ng++] = (uint8_t)c; else if (kind < 0) bad[nb++] = (uint8_t)c; else dull[nd++] = (uint8_t)c;
    }
    int n = 0, m = 0;
    for (int i = 0; i < ng; i++) order_max[n++] = good[i];
    for (int i = 0; i < nd; i++) order_max[n++] = dull[i];
    for (int i = 0; i < nb; i++) order_max[n++] = bad[i];
    for (int i = 0; i < nb; i++) order_min[m++] = bad[i];
    for (int i = 0; i < nd; i++) order_min[m++] = dull[i];
    for (int i = 0; i < ng; i++) order_min[m++] = good[i];
    return n;
}

// does the mover's mark on `cell` end the game (a four or a three of the mover's through the cell)?  own: the mover's marks before
inline bool ns_move_ends_game(const NsTables& t, uint32_t own, int cell) {
    for (int k = 0; k < t.n_cell_quads[cell]; k++) if (__builtin_popcount(t.cell_quads[cell][k] & own) == 3) return true;
    for (int k = 0; k < t.n_cell_trips[cell]; k++) if (__builtin_popcount(t.cell_trips[cell][k] & own) == 2) return true;
    return false;
}

// An item's result is in host memory (mapped); now its index goes into the flight's landing list, which the host reads
// WHILE the kernel runs: results are consumed one by one, the slow items of a flight hold up nobody else.  The fence orders
// the result before the index for an observer on the host; the list slot comes from a counter in device memory.
__device__ __forceinline__ void ns_land(int* landed, unsigned int* n_landed, unsigned item_index) {
    __threadfence_system();
    const unsigned slot = atomicAdd(n_landed, 1u);
    *((volatile int*)landed + slot) = (int)item_index;
    __threadfence_system();
}

template <bool ITEMS>
__global__ void __launch_bounds__(kNsWarps * 32)
negascout_fast_kernel(const sq_board* __restrict__ pos, const int* __restrict__ idx, int n_pos, int max_depth, NsScores sc,
                      NsOut out, unsigned int* __restrict__ queue, const NsItem* __restrict__ items, NsItemOut* __restrict__ item_out,
                      unsigned long long budget, int* landed, unsigned int* n_landed) {
    __shared__ NsCellTables C;
    __shared__ NsFastFrame s_frames[kNsWarps][kNsMaxDepth + 2];
    __shared__ int s_cur[kNsWarps][kNsMaxDepth + 2][32], s_sum[kNsWarps][kNsMaxDepth + 2][32];
    __shared__ uint8_t s_order[kNsWarps][2][32];
    __shared__ uint32_t s_start[ITEMS ? kNsWarps : 1][kNsMaxDepth + 2], s_mark[ITEMS ? kNsWarps : 1][kNsMaxDepth + 2];   // see the give-up block
    for (int i = threadIdx.x; i < (int)(sizeof(NsCellTables) / 4); i += blockDim.x) ((uint32_t*)&C)[i] = ((const uint32_t*)&c_ns_cell)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    NsFastFrame* const frames = s_frames[warp];
    const uint32_t full = 0xffffffffu;
    const int kNegInf = -(1 << 30);

    const uint32_t my_quad = c_ns.quad[lane];
    const int my_weight = c_ns.quad_weight[lane] * 3;
    const uint32_t my_in = lane >= 28 ? c_ns.deadly_inner[lane - 28] : 0u;
    const uint32_t my_out = lane >= 28 ? c_ns.deadly_outer[lane - 28] : 0u;
    int my_score = 0;
#pragma unroll
    for (int c = 0; c < 25; c++) if (lane == c) my_score = sc.s[c];

    for (unsigned turn = 0;; turn++) {
        unsigned int p = 0;
        if (ITEMS) p = (turn * gridDim.x + blockIdx.x) * kNsWarps + warp;       // items: no queue (nothing to reset between flights)
        else {
            if (lane == 0) p = atomicAdd(queue, 1u);
            p = __shfl_sync(full, p, 0);
        }
        if (p >= (unsigned)n_pos) break;
        const unsigned item_index = p;
        NsItem item;
        int base = 0;                                  // ply of the loop this warp starts in
        if (ITEMS) {
            item = items[p]; p = (unsigned)item.pos; base = item.n_path - 1;
            // (a >= b happens: the reference re-searches (-beta, -cur) even when cur >= beta, negascout.go:251)
            SQ_CHECK(item.n_path >= 1 && item.n_path <= 7 && item.n_path <= max_depth, 20, item.n_path, item_index);
        } else p = idx[p];
        uint32_t x = pos[p].x & kFull25, o = pos[p].o & kFull25;
        unsigned long long nodes = 0;
        ns_reorder(x, o, lane, s_order[warp]);
        const int n_moves = __popc(kFull25 & ~(x | o));
        if (ITEMS)
            for (int j = 0; j < base; j++) { if (j & 1) o |= 1u << item.path[j]; else x |= 1u << item.path[j]; }
        int bias = __reduce_add_sync(full, lane < 25 ? (int)((x >> lane) & 1u) * my_score - (int)((o >> lane) & 1u) * my_score : 0);
        // the starting node's raw sum (no complete line on it: the host checked)
        int vsum;
        {
            const int cx = __popc(my_quad & x), co = __popc(my_quad & o);
            int contrib = (cx == 3 && co == 0) ? my_weight : (co == 3 && cx == 0) ? -my_weight : 0;
            if (lane >= 28) contrib += ns_deadly_term(my_in, my_out, x, o);
            vsum = __reduce_add_sync(full, contrib);
        }

        int ply = base;
        int alpha = -2 * 10000, beta = 2 * 10000, score = -3 * 10000, n = beta, k = 0, at = 0, phase = 0;
        uint32_t active = 0, term = 0;
        int keeper_max = -3 * 10000, first_best = -1, n_visited = 0;
        uint32_t keeper_mask = 0;
        bool aborted = false;
        if (!ITEMS && lane < 25) { out.values[p][lane] = SQ_NO_MOVE; out.visit_order[p][lane] = -1; }

        // one pass over the children of the node that starts its loop at `ply` (mover pl)
        auto children = [&](int at_ply, int raw_sum) {
            const int pl = (at_ply & 1) ? -1 : 1;
            const int c = lane < n_moves ? s_order[warp][pl > 0 ? 1 : 0][lane] : 0;
            const bool act = lane < n_moves && !(((x | o) >> c) & 1u);
            const NsChild ch = ns_child(C, x, o, pl, c, raw_sum, bias + pl * __shfl_sync(full, my_score, c), at_ply + 1);
            active = __ballot_sync(full, act);
            term = __ballot_sync(full, act && ch.term);
            return ch;
        };
        {
            const NsChild ch = children(base, vsum);
            s_cur[warp][base][lane] = ch.cur; s_sum[warp][base][lane] = ch.sum;
            __syncwarp();
        }
        int ret = 0, rs_cur = 0;
        bool returned = false, enter = false;
        if (ITEMS) {
            // the loop at `base` is about to call negaScout on its child path[base] with the window (a, b)
            const int cell = item.path[base], pl = (base & 1) ? -1 : 1;
            at = __ffs(__ballot_sync(full, lane < n_moves && s_order[warp][pl > 0 ? 1 : 0][lane] == cell)) - 1;
            SQ_CHECK(at >= 0 && !(((x | o) >> cell) & 1u), 21, cell, x | o);        // the path's last cell is one of the node's moves
            alpha = -item.b; n = -item.a; phase = 1; k = at + 1; enter = true;
            if (pl > 0) x |= 1u << cell; else o |= 1u << cell;
            bias += pl * __shfl_sync(full, my_score, cell);
        }
        for (;;) {
            const int player = (ply & 1) ? -1 : 1;
            if (!returned && !enter) {
                const uint32_t cand = active & (full << k);
                if (cand == 0) {                       // loop over: hand the score up
                    if (ply == 0) break;
                    ret = score;
                    ply--;
                    const NsFastFrame f = frames[ply];
                    alpha = f.alpha; beta = f.beta; score = f.score; n = f.n; k = f.k;
                    at = f.at_phase & 31; phase = f.at_phase >> 5; active = f.active; term = f.term;
                    returned = true;
                    continue;
                }
                at = __ffs(cand) - 1;
                k = at + 1;
                phase = 1;
                enter = true;
                const int cell = s_order[warp][player > 0 ? 1 : 0][at];
                if (player > 0) x |= 1u << cell; else o |= 1u << cell;
                bias += player * __shfl_sync(full, my_score, cell);
            }
            if (enter) {
                const int ca = phase == 1 ? -n : -beta, cb = phase == 1 ? -alpha : -rs_cur;
                const int cply = ply + 1;
                enter = false;
                if (ITEMS && phase == 1 && lane == 0) s_mark[warp][ply] = (uint32_t)nodes;     // calls so far when this child's first search starts
                nodes++;
                if (ITEMS ? (item.budget != 0 && nodes > item.budget) : nodes > budget) {     // the host takes over (NsSplit)
                    aborted = true;
                    if (ITEMS) {
                        // the loop to resume at: the shallowest one that has finished a child (all above it are at their first)
                        NsItemOut r;
                        r.ret = 0; r.score = -3 * 10000; r.nodes = kNsAborted; r.alpha = item.a; r.nodes_done = 0; r.n_chain = 0; r.i_done = 0;
                        __syncwarp();
                        for (int L = base + 1; L <= ply; L++) {
                            const uint32_t act = L == ply ? active : frames[L].active;
                            const int l_at = L == ply ? at : (frames[L].at_phase & 31);
                            const int fin = __popc(act & ((1u << l_at) - 1u));
                            if (fin > 0 || L == ply || L >= (int)item.deepest) {
                                r.i_done = (uint8_t)fin;
                                r.alpha = L == ply ? alpha : frames[L].alpha;
                                r.score = L == ply ? score : frames[L].score;
                                r.nodes_done = s_mark[warp][L] - s_start[warp][L];
                                SQ_CHECK(s_mark[warp][L] >= s_start[warp][L] && r.n_chain <= 5 && L <= kNsMaxDepth, 22, s_mark[warp][L], s_start[warp][L]);
                                break;
                            }
                            if (r.n_chain < 6) r.cells[r.n_chain] = s_order[warp][(L & 1) ? 0 : 1][l_at];
                            r.n_chain++;
                        }
                        if (lane == 0) { item_out[item_index] = r; ns_land(landed, n_landed, item_index); }
                    }
                    break;
                }
                if ((term >> at) & 1u) {               // the child is a finished game
                    ret = -s_cur[warp][ply][at];
                    returned = true;
                } else if (cply == max_depth) {
                    // the child's own children all stop: its whole loop (:240-262) as scans over the lanes
                    const uint32_t act_save = active, term_save = term;
                    const NsChild ch = children(cply, s_sum[warp][ply][at]);
                    const uint32_t act = active;
                    active = act_save; term = term_save;
                    const bool mine = (act >> lane) & 1u;
                    int incl = mine ? ch.cur : kNegInf;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int up = __shfl_up_sync(full, incl, d);
                        if (lane >= d) incl = max(incl, up);
                    }
                    int excl = __shfl_up_sync(full, incl, 1);
                    if (lane == 0) excl = kNegInf;
                    const int first = __ffs(act) - 1;
                    const int before = max(ca, excl);                        // alpha when this child is reached
                    const bool visited = mine && (lane == first || before < cb);
                    const bool again = visited && lane != first && ch.cur > excl && before + 1 != cb;   // re-searched (:251)
                    const uint32_t vmask = __ballot_sync(full, visited), amask = __ballot_sync(full, again);
                    nodes += __popc(vmask) + __popc(amask);
                    ret = vmask ? __shfl_sync(full, incl, 31 - __clz(vmask)) : -3 * 10000;
                    returned = true;
                } else {
                    NsFastFrame f;
                    f.alpha = alpha; f.beta = beta; f.score = score; f.n = n; f.k = k; f.at_phase = at | (phase << 5);
                    f.active = active; f.term = term;
                    if (lane == 0) frames[ply] = f;
                    SQ_CHECK(cply <= kNsMaxDepth + 1, 23, cply, max_depth);                         // frames and child arrays hold kNsMaxDepth + 2 plies
                    const NsChild ch = children(cply, s_sum[warp][ply][at]);
                    s_cur[warp][cply][lane] = ch.cur; s_sum[warp][cply][lane] = ch.sum;
                    __syncwarp();
                    ply = cply;
                    if (ITEMS && lane == 0) s_start[warp][cply] = (uint32_t)nodes;
                    alpha = ca; beta = cb; score = -3 * 10000; n = cb; k = 0;
                    continue;
                }
            }
            // `ret` came back to the running loop (:243-262)
            if (ITEMS && ply == base) {                  // ... which, for an item, is the host's
                if (lane == 0) { item_out[item_index].ret = ret; item_out[item_index].nodes = nodes; ns_land(landed, n_landed, item_index); }
                break;
            }
            returned = false;
            if (phase == 1) {
                const int cur = -ret;
                if (cur > score) {
                    if (n == beta || (ply > 0 && ply == max_depth - 2)) score = cur;
                    else { phase = 2; rs_cur = cur; enter = true; continue; }
                }
            } else {
                score = -ret;
            }
            const int cell = s_order[warp][player > 0 ? 1 : 0][at];
            if (player > 0) x &= ~(1u << cell); else o &= ~(1u << cell);
            bias -= player * __shfl_sync(full, my_score, cell);
            if (score > alpha) alpha = score;
            if (ply == 0) {                            // moves.SetMove(i, j, alpha), movekeeper.go:26-36
                if (lane == 0) { out.values[p][cell] = alpha; out.visit_order[p][n_visited] = (int8_t)cell; }
                n_visited++;
                if (alpha >= keeper_max) {
                    if (alpha > keeper_max) { keeper_max = alpha; keeper_mask = 0; first_best = cell; }
                    keeper_mask |= 1u << cell;
                }
            }
            if (alpha >= beta) {
                if (ply == 0) break;
                ret = score;
                ply--;
                const NsFastFrame f = frames[ply];
                alpha = f.alpha; beta = f.beta; score = f.score; n = f.n; k = f.k;
                at = f.at_phase & 31; phase = f.at_phase >> 5; active = f.active; term = f.term;
                returned = true;
                continue;
            }
            n = alpha + 1;
        }
        if (!ITEMS && lane == 0) {
            out.best_value[p] = keeper_mask ? keeper_max : 0;
            out.best_mask[p] = keeper_mask;
            out.first_best[p] = (int8_t)first_best;
            out.nodes[p] = aborted ? kNsAborted : nodes;
        }
        __syncwarp();
    }
}

}  // namespace sq
