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

__global__ void __launch_bounds__(kNsWarps * 32)
negascout_kernel(const sq_board* __restrict__ pos, int n_pos, int max_depth, NsScores sc, NsOut out,
                 unsigned int* __restrict__ queue) {
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
        uint32_t x = pos[p].x & kFull25, o = pos[p].o & kFull25;
        unsigned long long nodes = 0;

        // ---- reorderMoves (negascout.go:411-499): lane i classifies the i-th cell of the initial order
        {
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
                s_order[warp][1][kind > 0 ? within : kind == 0 ? ng + within : ng + nd + within] = (uint8_t)c;
                s_order[warp][0][kind < 0 ? within : kind == 0 ? nb + within : nb + nd + within] = (uint8_t)c;
            }
            __syncwarp();
        }
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

}  // namespace sq
