// alphabeta.cuh -- fixed-depth alpha-beta as a GPU batch of independent subtrees.
//
// What is computed: for every root move (or explicit subtree) the value the
// reference's sequential search returns.  Every reference entry point searches
// with a window that contains all reachable values, so
// that value is the plain minimax value of the dialect's tree and does not
// depend on visiting order or on how bounds are shared.  That freedom is what
// the GPU uses:
//   1. the top 1-3 plies under each root move are EXPANDED breadth-first into
//      nodes (kernels ab_roots / ab_expand); every node has a value slot and a
//      count of unfinished children;
//   2. the deepest nodes are UNITS: one thread runs a sequential, explicit-stack
//      alpha-beta over each (kernel ab_dfs, lanes pull units from a queue);
//   3. a finished node folds its value into its parent's slot (atomicMin/Max)
//      and the last child to finish carries the parent's value further up;
//   4. a unit's window is tightened by the current slot values of its ancestors
//      (alpha from MAX ancestors, beta from MIN ancestors), re-read whenever the
//      unit's own root frame gets a child back.
//
// Dialects (SQ_AB_*): 1 src/alphabeta (delayed evaluation, carried value is the
// previous delta only), 2 squavathr (evaluate at entry, cumulative, no2 /
// noMiddle2 penalties), 3/4 opening3 / opening2 (evaluate at entry, cumulative,
// orderedCells child order).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/squava_b200.h"
#include "sq_device.cuh"
#include "sq_tables.h"
#include "ab_horizon.cuh"

namespace sq {

constexpr int kAbMaxDepth = 24;
constexpr int kAbThreads = 128;
constexpr int kWin = 10000;

struct AbTables {
    uint32_t quads[25][8];     // quads through each cell (API layout), 0-terminated
    uint32_t triplets[25][12]; // triplets through each cell
    uint8_t n_quads[25], n_triplets[25];
    uint32_t no2[25][2];        // squavathr.go:305-311 triplets containing the cell (0 = none)
    uint32_t nomid_quad[25][2];   // squavathr.go:297-302: quads in which the cell is a middle cell
    uint32_t nomid_other[25][2];  // ... and the bit of the other middle cell of that quad
    uint8_t order[25];          // opening3.go:328-361: k-th child cell
    uint8_t rank[25];           // inverse
};
__constant__ AbTables c_ab;  // copied to shared memory by each kernel: lanes index it by different cells

inline void build_ab_tables(AbTables& t) {
    memset(&t, 0, sizeof t);
    ScanTables s;
    build_scan_tables(s);
    for (int c = 0; c < 25; c++) {
        for (int i = 0; i < 28; i++)
            if (s.ab_quads[i] >> c & 1u) t.quads[c][t.n_quads[c]++] = s.ab_quads[i];
        for (int i = 0; i < 48; i++)
            if (s.ab_triplets[i] >> c & 1u) t.triplets[c][t.n_triplets[c]++] = s.ab_triplets[i];
    }
    // squavathr.go:305-311 -- the four corner-cutting diagonals of length 3
    const int no2[4][3] = {{10, 6, 2}, {2, 8, 14}, {22, 18, 14}, {22, 16, 10}};
    for (auto& tr : no2) {
        uint32_t m = (1u << tr[0]) | (1u << tr[1]) | (1u << tr[2]);
        for (int c : tr) { if (!t.no2[c][0]) t.no2[c][0] = m; else t.no2[c][1] = m; }
    }
    // squavathr.go:297-302 -- the four diagonals of length 4; [1] and [2] are the middles
    const int nm[4][4] = {{15, 11, 7, 3}, {5, 11, 17, 23}, {1, 7, 13, 19}, {9, 13, 17, 21}};
    for (auto& q : nm) {
        uint32_t m = (1u << q[0]) | (1u << q[1]) | (1u << q[2]) | (1u << q[3]);
        for (int a = 1; a <= 2; a++) {     // each inner-ring corner is a middle of two of them
            int c = q[a], other = q[3 - a], k = t.nomid_quad[c][0] ? 1 : 0;
            t.nomid_quad[c][k] = m; t.nomid_other[c][k] = 1u << other;
        }
    }
    // opening3.go:328-361: inner-ring corners, centre, inner-ring edges, then the outer
    // ring's edge triples (top, left, right, bottom), corners last
    const uint8_t ord[25] = {6, 8, 16, 18, 12, 7, 11, 13, 17, 1, 2, 3, 5, 10, 15, 9, 14, 19, 21, 22, 23, 0, 4, 20, 24};
    for (int k = 0; k < 25; k++) { t.order[k] = ord[k]; t.rank[ord[k]] = (uint8_t)k; }
}

// dialect traits: 1, 5 and 6 share the delayed-evaluation search of src/alphabeta; 6 (src/abbook)
// sums the evaluations along the path; 5 adds the
// deltaValue2 ("avoid") terms, which are squavathr's no2 / noMiddle2 penalties
template <int D> struct AbTraits {
    static constexpr bool delayed = D == 1 || D == 5 || D == 6;
    static constexpr bool delayed_sum = D == 6;        // src/abbook hands boardValue+delta down (abbook.go:228,255)
    static constexpr bool penalties = D == 2 || D == 5;
    static constexpr bool ordered = D == 3;           // dialect 4 runs as 3 with another sentinel
};

struct AbNode {
    uint32_t x, o;       // API-layout masks, the pending move already on the board
    int32_t carried;     // boardValue argument
    int8_t cell;         // pending (last) move
    int8_t ply;
    int8_t side;         // side to move at this node
    int8_t split;        // further plies to expand breadth-first; 0 = DFS unit
    int32_t parent;      // node index, or -1 - output index
    int32_t group;       // position / subtree index the evaluator count is charged to
};

struct AbRun {
    AbNode* nodes;
    int* value;                      // slot values
    int* pending;                    // unfinished children
    int* units;                      // indices of DFS units
    int* aborted;                    // units that ran out of budget in this pass: split one ply deeper
    int* kids;                       // first child of a split node (its children are contiguous, eldest first), -1: none
    int* held;                       // younger children of a split node still waiting for the eldest to finish
    int* released;                   // younger children set free in this pass: units of the next one
    unsigned int* counters;          // [0] n_nodes [1] n_units = queue tail [2] queue head [3] n_aborted [4] node pool overflowed
                                     // [5] n_released [6] smallest ply among the next pass's units
                                     // queue-driven search: [7] units alive (queued or being searched) [8] watchdog fired [9] heartbeat
    int* out;                        // final values
    unsigned long long* node_count;  // evaluator calls per group
    int cstride;                     // counter k lives at counters[k * cstride]: 1 (pass-driven, the host reads them as one block) or
                                     // 32 (queue-driven: every counter on its own 128-byte line, so that the lanes polling one
                                     // do not hold up the atomics on another)
    int node_capacity;
    int queue_capacity;              // slots of units[] (queue-driven: nodes + room for the rare slots handed back empty)
    int max_depth;
    int sentinel;        // |value| of a node with no legal move: 20000, or 10000 (opening2)
    int win_lo, win_hi;  // the dialect's root window
    int shortcuts;       // exact shortcuts allowed (see ab_move_classes); off when |heuristic sums| could reach a win score
    int queued;          // queue-driven search: ONE persistent kernel, units flow through a device-side queue (see ab_dfs_kernel)
    int gen_scaling;           // queue-driven: every split above a unit doubles what it must cost before it is split again
    unsigned int min_budget;   // queue-driven: evaluator calls a unit spends before it may be split for hungry lanes
    long long stall_cycles;    // queue-driven watchdog: idle lanes give up when the heartbeat stands still this long
    int ybw;             // young brothers wait: a split node's younger children start only when the eldest has finished
    int defer_min;       // > 0: horizon evaluations are deferred until this many lanes of the warp need one (SQ_AB_DEFER)
    int ybw_below_ply;   // brothers wait only under split nodes above this ply (SQ_AB_YBW = number of plies below the level-0 nodes)
    int two_killers;     // a second killer per depth (the previous one)
    int killers;         // killer-move ordering inside the per-lane search (SQ_AB_NO_KILLERS switches it off)
    int fast_horizon;    // horizon-1 nodes are valued in one bit-parallel shot (ab_horizon.cuh): needs the shortcuts and a bias table spanning < 30
    int8_t scores[25];
    HorizonTables hz;
};

__host__ __device__ __forceinline__ unsigned int* ab_cnt(const AbRun& R, int k) { return R.counters + k * R.cstride; }

// ---- evaluators -----------------------------------------------------------------
struct Eval { bool terminal; int value; };

// The lines-through-(cell) part common to every dialect (alphabeta.go:123-147,
// squavathr.go:317-341, opening3.go:138-164): own = mask of the colour standing
// on `cell`, opp = the other colour, sign = +1 if own is X.  A sum of +-3 over a
// quad through `cell` can only be three of `own` and one empty cell.
__device__ __forceinline__ Eval eval_lines(const AbTables& T, uint32_t own, uint32_t opp, int cell,
                                           int sign, int ply) {
    Eval e; e.terminal = false; e.value = 0;
    int threes = 0; bool four = false, three = false;
    const int nq = T.n_quads[cell];
    for (int k = 0; k < nq; k++) {
        uint32_t q = T.quads[cell][k];
        uint32_t mine = own & q;
        four |= (mine == q);
        threes += (__popc(mine) == 3 && (opp & q) == 0) ? 1 : 0;
    }
    const int nt = T.n_triplets[cell];
    for (int k = 0; k < nt; k++) {
        uint32_t t = T.triplets[cell][k];
        three |= ((own & t) == t);
    }
    if (four) { e.terminal = true; e.value = sign * (kWin - ply); }        // four beats three
    else if (three) { e.terminal = true; e.value = -sign * (kWin - ply); }
    else e.value = sign * 30 * threes;
    return e;
}

// squavathr.go:343-371: extra terms for the mark on `cell`
__device__ __forceinline__ int eval_penalties(const AbTables& T, uint32_t own, uint32_t opp, int cell, int sign) {
    int v = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        uint32_t t = T.no2[cell][k];
        if (t && __popc(own & t) == 2 && (opp & t) == 0) v -= sign * 100;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        uint32_t q = T.nomid_quad[cell][k];
        if (q && (own & T.nomid_other[cell][k]) && __popc(own & q) - __popc(opp & q) == 2) v -= sign * 100;
    }
    return v;
}

// squavathr.go:244: deltaValue on the board BEFORE the root mark is placed -- only
// the three-of-four sums of the other cells can contribute.
__device__ __forceinline__ int eval_unmarked(const AbTables& T, uint32_t x, uint32_t o, int cell) {
    int v = 0;
    const int nq = T.n_quads[cell];
    for (int k = 0; k < nq; k++) {
        uint32_t q = T.quads[cell][k];
        int sum = __popc(x & q) - __popc(o & q);
        if (sum == 3 || sum == -3) v += sum * 10;
    }
    return v;
}

__device__ __forceinline__ void load_tables(AbTables& dst) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&c_ab);
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    for (int i = threadIdx.x; i < (int)(sizeof(AbTables) / 4); i += blockDim.x) d[i] = src[i];
    __syncthreads();
}

// ---- node bookkeeping ---------------------------------------------------------------
__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// Node data is written by one lane and read by lanes on OTHER multiprocessors while the same kernel runs (queue-driven
// search), and a node shares its 32-byte sector with its neighbours: a plain load could be served from a line this
// multiprocessor's L1 fetched before the node was written.  Everything that is written during the search is therefore
// read from L2 (ld.global.cg).
__device__ __forceinline__ AbNode ab_ld_node(const AbNode* p) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
    const uint2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2);
    AbNode n;
    n.x = a.x; n.o = a.y; n.carried = (int32_t)b.x;
    n.cell = (int8_t)(b.y & 0xffu); n.ply = (int8_t)((b.y >> 8) & 0xffu); n.side = (int8_t)((b.y >> 16) & 0xffu); n.split = (int8_t)(b.y >> 24);
    n.parent = (int32_t)c.x; n.group = (int32_t)c.y;
    return n;
}
static_assert(sizeof(AbNode) == 24, "AbNode is read as three 8-byte words");
__device__ __forceinline__ int ab_node_parent(const AbRun& R, int i) { return __ldcg(reinterpret_cast<const int*>(&R.nodes[i]) + 4); }
__device__ __forceinline__ int ab_node_side(const AbRun& R, int i) { return (int)(int8_t)((__ldcg(reinterpret_cast<const unsigned*>(&R.nodes[i]) + 3) >> 16) & 0xffu); }
__device__ __forceinline__ int ab_node_ply(const AbRun& R, int i) { return (int)(int8_t)((__ldcg(reinterpret_cast<const unsigned*>(&R.nodes[i]) + 3) >> 8) & 0xffu); }

// A node becomes a unit: of the next pass (pass-driven search) or of the running kernel (queue-driven: the slot
// write is what a lane holding that ticket is polling for, so everything about the node is fenced before it).
__device__ __forceinline__ void ab_enqueue(const AbRun& R, int idx, int ply) {
    if (R.queued) atomicAdd(ab_cnt(R, 7), 1u);
    const unsigned slot = atomicAdd(ab_cnt(R, 1), 1u);
    __threadfence();
    SQ_CHECK(slot < (unsigned)R.queue_capacity && idx >= 0 && idx < R.node_capacity, 13, slot, idx);
    if (slot < (unsigned)R.queue_capacity) *reinterpret_cast<volatile int*>(&R.units[slot]) = idx;
    else atomicExch(ab_cnt(R, 4), 1u);             // cannot happen by sizing; never write past the queue
    if (!R.queued) atomicMin(ab_cnt(R, 6), (unsigned)ply);
}

// A finished node hands its value to its parent; the last child to finish carries on upward.
__device__ void ab_complete(const AbRun& R, int node, int value) {
    for (;;) {
        int parent = ab_node_parent(R, node);
        SQ_CHECK(node >= 0 && node < R.node_capacity && parent < R.node_capacity, 10, node, parent);
        if (parent < 0) { R.out[-1 - parent] = value; return; }
        if (ab_node_side(R, parent) > 0) atomicMax(&R.value[parent], value);
        else atomicMin(&R.value[parent], value);
        __threadfence();
        // Young brothers wait: the eldest child of a split node has finished, so its value now bounds the parent's
        // slot -- the younger children become units of the next pass and start with that bound (only the thread
        // that completes the eldest gets here: no race on held[]).
        if (__ldcg(&R.kids[parent]) == node) {
            const int nh = __ldcg(&R.held[parent]);
            if (nh > 0) {
                R.held[parent] = 0;
                if (R.queued) {
                    for (int k = 0; k < nh; k++) ab_enqueue(R, node + 1 + k, 0);
                } else {
                    const unsigned at = atomicAdd(ab_cnt(R, 5), (unsigned)nh);
                    for (int k = 0; k < nh; k++) R.released[at + k] = node + 1 + k;
                    atomicMin(ab_cnt(R, 6), (unsigned)ab_node_ply(R, node));
                }
            }
        }
        int old = atomicSub(&R.pending[parent], 1);
        if (old != 1) return;
        __threadfence();
        value = ld_volatile(&R.value[parent]);
        node = parent;
    }
}

// alpha from the MAX ancestors' slots, beta from the MIN ancestors' slots
__device__ __forceinline__ void ab_bounds(const AbRun& R, int node, int& alpha, int& beta) {
    int p = ab_node_parent(R, node);
    while (p >= 0) {
        SQ_CHECK(p < R.node_capacity, 11, node, p);
        int v = ld_volatile(&R.value[p]);
        if (ab_node_side(R, p) > 0) alpha = max(alpha, v); else beta = min(beta, v);
        p = ab_node_parent(R, p);
    }
}

__device__ __forceinline__ int ab_new_node(const AbRun& R, const AbNode& n) {
    unsigned int idx = atomicAdd(ab_cnt(R, 0), 1u);
    if (idx >= (unsigned)R.node_capacity) {           // the host checks the room before every expansion; if its sizing is
        atomicExch(ab_cnt(R, 4), 1u);               // ever wrong the run is failed (SQ_ERR_NOMEM), never silently short
        return -1;
    }
    R.nodes[idx] = n;
    R.value[idx] = n.side > 0 ? -R.sentinel : R.sentinel;
    R.pending[idx] = 0;
    R.kids[idx] = -1;
    R.held[idx] = 0;
    if (n.split == 0) ab_enqueue(R, (int)idx, n.ply);
    return (int)idx;
}

// Resolve a node on entry, or report that it has children.  Returns true if `value`
// is final.  DIALECT 1 (alphabeta.go:165-223): the pending move is evaluated with the
// FIRST child's mark already on the board and a stop returns from there; DIALECTS 2-4
// evaluate it once at entry (squavathr.go:393-398, opening3.go:138-175).
template <int DIALECT>
__device__ __forceinline__ bool ab_enter(const AbTables& T, const AbRun& R, const AbNode& n, int& value,
                                         int& carried_down) {
    const uint32_t empties = kFull25 & ~(n.x | n.o);
    const bool pend_is_x = (n.x >> n.cell) & 1u;
    const int sign = pend_is_x ? 1 : -1;
    if (AbTraits<DIALECT>::delayed) {
        carried_down = n.carried;
        if (!empties) { value = n.side > 0 ? -R.sentinel : R.sentinel; return true; }
        uint32_t first = empties & (0u - empties);
        uint32_t x = n.x | (n.side > 0 ? first : 0u), o = n.o | (n.side > 0 ? 0u : first);
        Eval e = eval_lines(T, pend_is_x ? x : o, pend_is_x ? o : x, n.cell, sign, n.ply);
        if (e.terminal) { value = e.value; return true; }
        if (n.ply >= R.max_depth) {
            value = e.value + sign * R.scores[n.cell] + n.carried;
            if (AbTraits<DIALECT>::penalties) value += eval_penalties(T, pend_is_x ? x : o, pend_is_x ? o : x, n.cell, sign);
            return true;
        }
        return false;
    } else {
        Eval e = eval_lines(T, pend_is_x ? n.x : n.o, pend_is_x ? n.o : n.x, n.cell, sign, n.ply);
        if (e.terminal) { value = e.value; return true; }
        int d = e.value + sign * R.scores[n.cell];
        if (AbTraits<DIALECT>::penalties) d += eval_penalties(T, pend_is_x ? n.x : n.o, pend_is_x ? n.o : n.x, n.cell, sign);
        if (n.ply == R.max_depth) { value = d + n.carried; return true; }
        carried_down = n.carried + d;
        if (!empties) { value = n.side > 0 ? -R.sentinel : R.sentinel; return true; }
        return false;
    }
}

// Exact shortcuts at a node the mover `own` is about to move from (values of the dialects'
// trees, not heuristics):
//   win  = empty cells that complete a 4-line for the mover: that child is terminal with
//          +-(10000 - child ply), the best value any descendant can have as long as no node
//          below can run out of cells (a node without a legal move returns -+20000) -- so the
//          node's value is known without searching anything;
//   sui  = empty cells that complete a 3-line but no 4-line: that child is terminal with the
//          worst value the mover can get at this ply -- it is folded into the node's value and
//          the child is never visited.
// The caller applies them only when the subtree cannot reach a full board.  Lines are found
// with shifts in the padded layout; only lines THROUGH the candidate cell count, exactly like
// the reference's evaluators, so lines already on the board elsewhere are ignored.
__device__ __forceinline__ void ab_move_classes(uint32_t own, uint32_t empties, uint32_t& win, uint32_t& sui) {
    const uint32_t m = pad(own), e = pad(empties);
    uint32_t w = 0, l = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int d = k == 0 ? 1 : k == 1 ? 6 : k == 2 ? 7 : 5;
        const uint32_t a1 = m << d, a2 = m << (2 * d), a3 = m << (3 * d);   // own stone 1/2/3 steps before
        const uint32_t b1 = m >> d, b2 = m >> (2 * d), b3 = m >> (3 * d);   // ... after
        l |= (a1 & (a2 | b1)) | (b1 & b2);
        w |= (a1 & a2 & (a3 | b1)) | (b1 & b2 & (a1 | b3));
    }
    w &= e;
    l &= e & ~w;
    win = unpad(w);
    sui = unpad(l);
}

// ---- kernels ------------------------------------------------------------------------
// Root moves: ChooseMove's loop body, alphabeta.go:98-110 / squavathr.go:241-253.
// One thread per (position, cell).  split[i] = breadth-first plies below the root move.
template <int DIALECT>
__global__ void __launch_bounds__(kAbThreads)
ab_roots_kernel(AbRun R, const sq_board* __restrict__ pos, const int8_t* __restrict__ split, int n) {
    __shared__ AbTables T;
    load_tables(T);
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 25) return;
    int i = t / 25, c = t % 25;
    uint32_t x = pos[i].x & kFull25, o = pos[i].o & kFull25, bit = 1u << c;
    if ((x | o) & bit) { R.out[t] = INT32_MIN; return; }
    atomicAdd(&R.node_count[i], 1ull);
    AbNode nd;
    nd.x = x | bit; nd.o = o; nd.cell = (int8_t)c; nd.ply = 1; nd.side = -1;
    nd.split = split[i]; nd.parent = -1 - t; nd.group = i;
    if (AbTraits<DIALECT>::delayed) {
        Eval e = eval_lines(T, x | bit, o, c, 1, 0);                 // alphabeta.go:101-102
        if (e.terminal) { R.out[t] = e.value; return; }
        int v = e.value + R.scores[c];
        if (AbTraits<DIALECT>::penalties) v += eval_penalties(T, x | bit, o, c, 1);
        if (0 >= R.max_depth) { R.out[t] = v; return; }
        nd.carried = v;
    } else {
        int v = eval_unmarked(T, x, o, c);                             // squavathr.go:244
        if (0 == R.max_depth) { R.out[t] = v; return; }             // :245-247
        nd.carried = v;
    }
    ab_new_node(R, nd);
}

// Explicit subtrees: the caller's node is the level-0 node.
__global__ void __launch_bounds__(kAbThreads)
ab_subtree_nodes_kernel(AbRun R, const sq_subtree* __restrict__ st, const int8_t* __restrict__ split, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AbNode nd;
    nd.x = st[i].b.x & kFull25; nd.o = st[i].b.o & kFull25;
    nd.carried = st[i].carried; nd.cell = st[i].last_cell; nd.ply = st[i].ply;
    nd.side = st[i].side_to_move; nd.split = split[i]; nd.parent = -1 - i; nd.group = i;
    ab_new_node(R, nd);
}

// One ply of breadth-first expansion.  list == nullptr: the nodes [lo, lo+count) that still
// have split > 0 (static pre-split); otherwise the nodes list[0..count) -- units that ran out
// of budget in the last pass -- whose children become the next pass's units.
template <int DIALECT>
__global__ void __launch_bounds__(kAbThreads)
ab_expand_kernel(AbRun R, const int* __restrict__ list, int lo, int count) {
    __shared__ AbTables T;
    load_tables(T);
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int idx = list ? list[t] : lo + t;
    AbNode n = ab_ld_node(&R.nodes[idx]);
    if (list) n.split = 1;
    else if (n.split <= 0) return;
    int value = 0, carried_down = 0;
    unsigned long long evals = 1;
    if (ab_enter<DIALECT>(T, R, n, value, carried_down)) {
        atomicAdd(&R.node_count[n.group], evals);
        ab_complete(R, idx, value);
        return;
    }
    uint32_t empties = kFull25 & ~(n.x | n.o);
    // the same exact shortcuts as in the per-lane search (ab_move_classes)
    if (R.shortcuts && __popc(empties) > R.max_depth - n.ply) {
        uint32_t win, sui;
        ab_move_classes(n.side > 0 ? n.x : n.o, empties, win, sui);
        const int best = n.side * (kWin - (n.ply + 1));
        if (win) {
            atomicAdd(&R.node_count[n.group], evals);
            ab_complete(R, idx, best);
            return;
        }
        if (sui) {
            empties &= ~sui;
            if (!empties) {                       // every move completes a 3-line
                atomicAdd(&R.node_count[n.group], evals);
                ab_complete(R, idx, -best);
                return;
            }
            R.value[idx] = -best;                 // children fold into this with atomicMin/Max
        }
    }
    // the children take one contiguous run of slots, eldest (first in the dialect's visiting order) first
    const int nch = __popc(empties);
    const unsigned base = atomicAdd(ab_cnt(R, 0), (unsigned)nch);
    if (base + (unsigned)nch > (unsigned)R.node_capacity) {   // the host checks the room before every expansion; if its
        atomicExch(ab_cnt(R, 4), 1u);                       // sizing is ever wrong the run is failed, never silently short
        return;
    }
    // adaptive split of a unit that proved too big: only the eldest child starts now, the others when it is done
    const bool hold = list != nullptr && R.ybw && nch >= 2 && n.ply < R.ybw_below_ply;
    R.kids[idx] = (int)base;
    R.held[idx] = hold ? nch - 1 : 0;
    R.pending[idx] = nch;
    __threadfence();
    const bool pend_is_x = (n.x >> n.cell) & 1u;
    const int sign = pend_is_x ? 1 : -1;
    int j = 0;
    for (int k = 0; k < 25; k++) {               // children in the dialect's visiting order
        const int cell = AbTraits<DIALECT>::ordered ? T.order[k] : k;
        const uint32_t bit = 1u << cell;
        if (!(empties & bit)) continue;
        AbNode c;
        c.x = n.x | (n.side > 0 ? bit : 0u);
        c.o = n.o | (n.side > 0 ? 0u : bit);
        c.cell = (int8_t)cell;
        c.ply = n.ply + 1;
        c.side = -n.side;
        c.split = n.split - 1;
        c.parent = idx;
        c.group = n.group;
        if (AbTraits<DIALECT>::delayed) {   // the parent's move re-evaluated with this child's mark on
            Eval e = eval_lines(T, pend_is_x ? c.x : c.o, pend_is_x ? c.o : c.x, n.cell, sign, n.ply);
            c.carried = e.value + sign * R.scores[n.cell] + (AbTraits<DIALECT>::delayed_sum ? n.carried : 0);
            if (AbTraits<DIALECT>::penalties) c.carried += eval_penalties(T, pend_is_x ? c.x : c.o, pend_is_x ? c.o : c.x, n.cell, sign);
            evals++;
        } else {
            c.carried = carried_down;
        }
        const int ci = (int)base + j;
        R.nodes[ci] = c;
        R.value[ci] = c.side > 0 ? -R.sentinel : R.sentinel;
        R.pending[ci] = 0;
        R.kids[ci] = -1;
        R.held[ci] = 0;
        if (c.split == 0 && (!hold || j == 0)) ab_enqueue(R, ci, c.ply);
        j++;
    }
    atomicAdd(&R.node_count[n.group], evals);
}

// Queue-driven search: a lane hands its unit over to hungry lanes.  The unit's node gets children for the root
// moves the lane has NOT finished (`cells`, API layout: the untried ones and the one it was inside), its value slot
// is seeded with what the finished ones gave (a bound like any other child's), and the children go to the queue --
// all of them if there is a bound already, else the eldest only (the brothers wait for it, see ab_complete).
// Returns false when the node pool has no room: the lane then simply carries on searching.
template <int DIALECT>
__device__ __noinline__ bool ab_split_unit(const AbTables& T, const AbRun& R, int unit, uint32_t cells, int seed_value,
                                           bool have_seed, int carried_down, unsigned long long* evals) {
    const AbNode n = ab_ld_node(&R.nodes[unit]);
    const int nch = __popc(cells);
    const bool hold = R.ybw && !have_seed && nch >= 2;
    // The queue slots of the children that start now are reserved FIRST: the tail jumps at once, so the other lanes
    // stop seeing hungry tickets while this one is still building the nodes (no stampede of splits for one idle lane).
    const int npush = hold ? 1 : nch;
    unsigned base = *reinterpret_cast<volatile unsigned*>(ab_cnt(R, 0));
    if (base + (unsigned)nch > (unsigned)R.node_capacity) return false;       // clearly no room: nothing reserved, nothing to undo
    if (*reinterpret_cast<volatile unsigned*>(ab_cnt(R, 1)) + 64u > (unsigned)R.queue_capacity) return false;
    atomicAdd(ab_cnt(R, 7), (unsigned)npush);
    const unsigned slot0 = atomicAdd(ab_cnt(R, 1), (unsigned)npush);
    for (;;) {                                                   // reserve nch contiguous node slots, or none
        if (base + (unsigned)nch > (unsigned)R.node_capacity) {
            // no room after all (lost a race for the last slots): the reserved queue slots are handed to their ticket
            // holders as "nothing here, draw again".  Rare: the check above keeps a full pool from burning queue slots.
            for (int q = 0; q < npush; q++) *reinterpret_cast<volatile int*>(&R.units[slot0 + q]) = -2;
            atomicSub(ab_cnt(R, 7), (unsigned)npush);
            return false;
        }
        const unsigned seen = atomicCAS(ab_cnt(R, 0), base, base + (unsigned)nch);
        if (seen == base) break;
        base = seen;
    }
    R.kids[unit] = (int)base;
    R.held[unit] = hold ? nch - 1 : 0;
    if (have_seed) { if (n.side > 0) atomicMax(&R.value[unit], seed_value); else atomicMin(&R.value[unit], seed_value); }
    R.pending[unit] = nch;
    __threadfence();
    const bool pend_is_x = (n.x >> n.cell) & 1u;
    const int sign = pend_is_x ? 1 : -1;
    int j = 0;
    for (int k = 0; k < 25; k++) {               // children in the dialect's visiting order
        const int cell = AbTraits<DIALECT>::ordered ? T.order[k] : k;
        const uint32_t bit = 1u << cell;
        if (!(cells & bit)) continue;
        AbNode c;
        c.x = n.x | (n.side > 0 ? bit : 0u);
        c.o = n.o | (n.side > 0 ? 0u : bit);
        c.cell = (int8_t)cell;
        c.ply = n.ply + 1;
        c.side = -n.side;
        c.split = (int8_t)(n.split < 100 ? n.split + 1 : n.split);   // queue-driven: how many splits above this node
        c.parent = unit;
        c.group = n.group;
        if (AbTraits<DIALECT>::delayed) {   // the unit's pending move re-evaluated with this child's mark on
            Eval e = eval_lines(T, pend_is_x ? c.x : c.o, pend_is_x ? c.o : c.x, n.cell, sign, n.ply);
            c.carried = e.value + sign * R.scores[n.cell] + (AbTraits<DIALECT>::delayed_sum ? n.carried : 0);
            if (AbTraits<DIALECT>::penalties) c.carried += eval_penalties(T, pend_is_x ? c.x : c.o, pend_is_x ? c.o : c.x, n.cell, sign);
            (*evals)++;
        } else {
            c.carried = carried_down;
        }
        const int ci = (int)base + j;
        R.nodes[ci] = c;
        R.value[ci] = c.side > 0 ? -R.sentinel : R.sentinel;
        R.pending[ci] = 0;
        R.kids[ci] = -1;
        R.held[ci] = 0;
        j++;
    }
    // EVERY child is written before ANY is published: the eldest may finish at once and set its brothers free, and
    // whoever draws them must find their nodes complete
    __threadfence();
    for (int q = 0; q < npush; q++) *reinterpret_cast<volatile int*>(&R.units[slot0 + (unsigned)q]) = (int)base + q;
    return true;
}

// Sequential alpha-beta per lane over the queued units.
//
// The frame being searched lives in registers; the frames below it in shared memory
// ([frame][word][thread]: a warp's accesses are conflict-free).  A child that cannot have
// children of its own (depth stop, terminal, full board) is resolved in place and never
// pushed, so most evaluator calls touch no stack memory at all.
//
// One LINE WALK per child: the lines through the child's cell are walked once, with a fixed
// trip count over transposed, padded tables (conflict-free, no divergence), giving
//   four / three   the mover completed a 4-line / 3-line through the cell
//   e1, e2         the empty cells that complete one / two three-of-four quads through it
// DIALECT 1 evaluates a move once per CHILD of the node it leads to, the boards differing by
// the child's mark only: a mark on an e1 (e2) cell removes one (two) of the counted quads, so
// the per-child evaluation of the pending move is two bit tests on the walk done when the
// node was entered.  The same identity handles the mark a leaf places before it evaluates.
struct AbWalkTables {
    uint32_t quads[8][32];      // [k][cell]; unused slots hold kDummyLine
    uint32_t triplets[12][32];
};
constexpr uint32_t kDummyLine = 0xC0000000u;   // two bits no board has: never owned, never three-of-four

__device__ __forceinline__ void ab_load_walk_tables(AbWalkTables& W) {
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) {
        int k = i >> 5, c = i & 31;
        W.quads[k][c] = (c < 25 && k < c_ab.n_quads[c]) ? c_ab.quads[c][k] : kDummyLine;
    }
    for (int i = threadIdx.x; i < 12 * 32; i += blockDim.x) {
        int k = i >> 5, c = i & 31;
        W.triplets[k][c] = (c < 25 && k < c_ab.n_triplets[c]) ? c_ab.triplets[c][k] : kDummyLine;
    }
    __syncthreads();
}

struct AbWalk { bool four, three; uint32_t e1, e2; };

__device__ __forceinline__ AbWalk ab_walk(const AbWalkTables& W, uint32_t own, uint32_t opp, int cell) {
    uint32_t e1 = 0, e2 = 0, min_q = 0xffffffffu, min_t = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t hole = W.quads[k][cell] & ~own;          // cells of the quad the mover lacks
        min_q = min(min_q, hole);                                // 0 <=> the quad is complete
        const uint32_t bad = hole & ((hole - 1u) | opp);         // more than one hole, or the hole is taken
        const uint32_t h = bad ? 0u : hole;
        e2 |= e1 & h;
        e1 |= h;
    }
#pragma unroll
    for (int k = 0; k < 12; k++) min_t = min(min_t, W.triplets[k][cell] & ~own);
    AbWalk r;
    r.four = min_q == 0; r.three = min_t == 0; r.e1 = e1; r.e2 = e2;
    return r;
}

// words of a frame in shared memory: the saved register frame (5 or 3) + the KILLER of that depth: the child that
// last cut a node off there is tried first at the next node of the same depth.  Any visiting order gives the same
// root values; the reference's own order (row-major in src/alphabeta) prunes poorly.
template <int DIALECT> struct AbFrameWords { static constexpr int saved = AbTraits<DIALECT>::delayed ? 5 : 3; static constexpr int n = saved + 2; };

//
// QUEUED = false: pass-driven.  The host launches one pass per generation of units; a unit that exceeds `budget`
// is handed back (aborted[]) and split by ab_expand_kernel before the next pass.
// QUEUED = true: queue-driven.  ONE persistent launch per search.  units[] is a ticket queue in global memory:
// counters[1] = tail (bumped by whoever makes a unit: the root kernels, a lane that splits its unit, ab_complete
// setting waiting brothers free), counters[2] = head (a lane without work draws a ticket and polls ITS slot until
// a unit lands there), counters[7] = units alive; the search is over when that count reaches zero.  A lane splits
// its unit only when it has spent min_budget evaluator calls on it AND some lane is hungry (head > tail): small
// subtrees and busy machines keep sequential pruning, big subtrees on an idle machine are spread one generation
// at a time, eldest child first (young brothers wait), with no host round trip and no barrier in between.
template <int DIALECT, bool QUEUED>
__global__ void __launch_bounds__(kAbThreads)
ab_dfs_kernel(AbRun R, int n_units, unsigned long long budget, int n_frames) {
    __shared__ AbTables T;
    __shared__ AbWalkTables W;
    extern __shared__ uint32_t s_stack[];      // [n_frames][words][kAbThreads]
    load_tables(T);
    ab_load_walk_tables(W);
    constexpr bool kOrdered = AbTraits<DIALECT>::ordered;    // children in orderedCells order
    constexpr bool kDelayed = AbTraits<DIALECT>::delayed;
    constexpr bool kPenalties = AbTraits<DIALECT>::penalties;
    constexpr int kWords = AbFrameWords<DIALECT>::n;
    uint32_t* const my = s_stack + threadIdx.x;
#define SQ_FR(f, w) my[((f) * kWords + (w)) * kAbThreads]

    int sp = -1, unit = -1, base_ply = 0, base_side = 1, group = 0;
    uint32_t x = 0, o = 0, emp = 0;            // emp: empty cells in child-order space
    // the frame in registers
    uint32_t moves = 0;                        // children still to try (child-order space)
    int value = 0, alpha = 0, beta = 0, carried = 0;
    uint32_t fe1 = 0, fe2 = 0;                 // DIALECT 1: walk of the pending move ...
    int fpc = 0;                               // ... which stands on this cell
    unsigned long long evals = 0;
    bool fresh = false;                        // the register frame was just entered
    (void)n_frames;
    // deferred horizon evaluations (pass-driven search, R.defer_min > 0): a lane that needs one waits until enough lanes
    // of its warp need one too, so that the most expensive phase of the loop runs with a fuller warp
    bool pend_h = false, done = false;
    uint32_t d_bit = 0, d_kbit = 0, d_e1 = 0, d_e2 = 0;
    int d_konst = 0, stall = 0;
    const bool kDefer = !QUEUED && R.defer_min > 0;
    // queue-driven search
    unsigned ticket = 0xffffffffu;             // the queue slot this lane is waiting on
    bool no_split = false;                     // the node pool had no room: finish the unit alone
    int gen = 0;                               // splits above the unit: each one doubles what it must cost before it is split again
    unsigned iter = 0;
    unsigned beat_seen = 0; long long beat_at = 0;
    unsigned idle_polls = 0, looks = 0;
    unsigned long long busy_turns = 0, idle_turns = 0;   // statistics (SQ_AB_TRACE): loop turns with / without a unit

    auto gen_shift = [&](int g) -> int { return R.gen_scaling ? g : 0; };
    auto order_mask = [&](uint32_t cells) -> uint32_t {   // API layout -> child-order space
        if (!kOrdered) return cells;
        uint32_t m = 0;
        while (cells) { int c = __ffs(cells) - 1; cells &= cells - 1; m |= 1u << T.rank[c]; }
        return m;
    };

    for (;;) {
        int ret = 0;
        uint32_t cur_kbit = 0;    // the child (child-order space) whose value is being handed to frame sp
        bool returning = false;   // `ret` is to be handed to frame sp ...
        bool restore = false;     // ... which is in shared memory (a pushed child came back)
        bool served = false;
        if (kDefer) {
            const unsigned hm = __ballot_sync(0xffffffffu, pend_h), dm = __ballot_sync(0xffffffffu, done);
            if (dm == 0xffffffffu) break;
            if (hm) {
                // serve when enough lanes wait, when nobody else can make progress, or when someone has waited two turns
                const bool go = __popc(hm) >= R.defer_min || (hm | dm) == 0xffffffffu || __any_sync(0xffffffffu, pend_h && stall >= 2);
                if (pend_h) {
                    if (!go) { stall++; continue; }
                    const int s_side = (sp & 1) ? -base_side : base_side;
                    const int s_cply = base_ply + sp + 1;
                    const uint32_t xs = x | (s_side > 0 ? d_bit : 0u), os = o | (s_side > 0 ? 0u : d_bit);
                    const uint32_t own2 = s_side > 0 ? os : xs, opp2 = s_side > 0 ? xs : os;
                    int rel;
                    if (kDelayed) {
                        const uint32_t emp_after = emp & ~d_kbit;                 // child-order space == cell space here
                        const uint32_t f0 = emp_after & (0u - emp_after);
                        rel = horizon_node_value<true>(own2, opp2, emp_after, kWin - (s_cply + 1), d_konst, d_e1, d_e2, R.scores[__ffs(f0) - 1], R.hz);
                    } else {
                        rel = horizon_node_value<false>(own2, opp2, kFull25 & ~(xs | os), kWin - (s_cply + 1), d_konst, 0u, 0u, 0, R.hz);
                    }
                    ret = -s_side * rel;
                    cur_kbit = d_kbit;
                    returning = true; served = true; pend_h = false; stall = 0;
                }
            }
            if (done) continue;
        }
        if (!served) {
        iter++;
        if (QUEUED) { if (sp < 0) idle_turns++; else busy_turns++; }
        if (sp < 0) {
            if (QUEUED) {
                // a waiting lane shares its warp with working ones: it looks at its slot on every 32nd turn only
                if (ticket != 0xffffffffu && (iter & 31u) != 0) continue;
                if (ticket == 0xffffffffu) ticket = atomicAdd(ab_cnt(R, 2), 1u);
                const int got = ticket < (unsigned)R.queue_capacity ? *reinterpret_cast<const volatile int*>(&R.units[ticket]) : -1;
                if (got == -2) { ticket = 0xffffffffu; continue; }       // a split that found no room gave its slots back
                if (got < 0) {
                    // nothing in my slot yet.  Every eighth look (the slot itself is the lane's own line; these are shared
                    // ones): done if no unit is alive anywhere, and the watchdog -- working lanes advance the heartbeat; if
                    // it stands still for stall_cycles, give up
                    if ((looks++ & 7u) == 0) {
                        if (ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 7))) <= 0 || ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 8))) != 0) break;
                        const unsigned beat = (unsigned)ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 9)));
                        const long long now = clock64();
                        if (beat != beat_seen || beat_at == 0) { beat_seen = beat; beat_at = now; }
                        else if (now - beat_at > R.stall_cycles) { atomicExch(ab_cnt(R, 8), 1u); break; }
                    }
                    // A warp in which EVERY lane is waiting sleeps between looks, longer each time (up to 16 us): a hundred
                    // thousand lanes polling flat out would take the L2 away from the lanes that work.  A lane that waits
                    // beside working lanes must not sleep -- the warp moves at the pace of its slowest branch.
                    if (__activemask() == 0xffffffffu) { __nanosleep(250u << (idle_polls < 6 ? idle_polls : 6)); idle_polls++; }
                    continue;
                }
                __threadfence();
                ticket = 0xffffffffu; no_split = false; beat_at = 0; idle_polls = 0; looks = 0;
                unit = got;
            } else {
                unsigned int u = atomicAdd(ab_cnt(R, 2), 1u);
                if (u >= (unsigned)n_units) { if (kDefer) { done = true; continue; } break; }
                unit = R.units[u];
            }
            SQ_CHECK(unit >= 0 && unit < R.node_capacity, 12, unit, R.node_capacity);
            const AbNode n = ab_ld_node(&R.nodes[unit]);
            SQ_CHECK((n.side == 1 || n.side == -1) && n.cell >= 0 && n.cell < 25 && !(n.x & n.o) && (((n.x | n.o) >> n.cell) & 1u), 14, unit, n.x | n.o);
            int a0 = R.win_lo, b0 = R.win_hi;
            ab_bounds(R, unit, a0, b0);
            if (a0 >= b0) {
                // Already cut off by its ancestors' slots (or by a root-window edge: a MIN ancestor at -20000).
                // It hands up a value its parent's atomicMin/Max ignores -- the parent is the other kind of
                // node -- so the unit contributes nothing, as if the sequential search had never reached it.
                ab_complete(R, unit, n.side > 0 ? 30000 : -30000);
                if (QUEUED) atomicSub(ab_cnt(R, 7), 1u);
                continue;
            }
            int v0 = 0, carried_down = 0;
            evals = 1; group = n.group;
            if (ab_enter<DIALECT>(T, R, n, v0, carried_down)) {
                atomicAdd(&R.node_count[group], evals);
                ab_complete(R, unit, v0);
                if (QUEUED) atomicSub(ab_cnt(R, 7), 1u);
                continue;
            }
            x = n.x; o = n.o; base_ply = n.ply; base_side = n.side;
            if (QUEUED) gen = n.split > 8 ? 8 : n.split;
            emp = order_mask(kFull25 & ~(x | o));
            sp = 0;
            moves = emp;
            value = n.side > 0 ? -R.sentinel : R.sentinel;
            alpha = a0; beta = b0; carried = carried_down;
            if (kDelayed) {       // the pending move is not terminal (ab_enter said so): keep its walk
                const bool px = (x >> n.cell) & 1u;
                AbWalk w = ab_walk(W, px ? x : o, px ? o : x, n.cell);
                fe1 = w.e1; fe2 = w.e2; fpc = n.cell;
            }
            fresh = true;
            continue;
        }
        if (!QUEUED && evals > budget) {   // too big for one lane: hand the unit back to be split one ply deeper
            atomicAdd(&R.node_count[group], evals);
            R.aborted[atomicAdd(ab_cnt(R, 3), 1u)] = unit;
            sp = -1;
            continue;
        }
        if (QUEUED && (iter & 31u) == 0) {
            if ((iter & 255u) == 0) *reinterpret_cast<volatile unsigned*>(ab_cnt(R, 9)) = (unsigned)clock64();   // heartbeat of the working lanes
            if (ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 8))) != 0) break;            // the watchdog fired somewhere
            // hungry lanes (tickets drawn beyond the tail) and this unit has had its minimum: split it
            // The node pool cannot grow while the kernel runs and nodes are not recycled: the fuller it gets, the more a
            // unit must have cost before it may be split (x4 above one half, x32 above three quarters, x256 above 7/8).
            bool split_now = false;
            if (!no_split && !fresh && evals >= ((unsigned long long)R.min_budget << gen_shift(gen)) &&
                (unsigned)ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 2))) > (unsigned)ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 1)))) {
                const unsigned used = (unsigned)ld_volatile(reinterpret_cast<const int*>(ab_cnt(R, 0))), cap = (unsigned)R.node_capacity;
                const int shift = gen_shift(gen) + (used > cap / 2 ? (used > cap / 4 * 3 ? (used > cap / 8 * 7 ? 8 : 5) : 2) : 0);
                split_now = evals >= ((unsigned long long)R.min_budget << shift);
            }
            if (split_now) {
                uint32_t rmoves; int rvalue, rcarried;
                if (sp == 0) { rmoves = moves; rvalue = value; rcarried = carried; }
                else {
                    const uint32_t w0 = SQ_FR(0, 0);
                    rmoves = (w0 & kFull25) | (1u << (w0 >> 25));                 // the child the lane is inside goes back too
                    rvalue = (int)(short)(SQ_FR(0, 1) & 0xffffu);
                    rcarried = (int)(short)(SQ_FR(0, 2) >> 16);
                }
                uint32_t cells = rmoves;
                if (kOrdered) { cells = 0; for (uint32_t m = rmoves; m; m &= m - 1) cells |= 1u << T.order[__ffs(m) - 1]; }
                const bool have_seed = rvalue != (base_side > 0 ? -R.sentinel : R.sentinel);
                atomicAdd(ab_cnt(R, 12), 1u);
                if (cells && ab_split_unit<DIALECT>(T, R, unit, cells, rvalue, have_seed, rcarried, &evals)) {
                    atomicAdd(&R.node_count[group], evals);
                    atomicSub(ab_cnt(R, 7), 1u);          // after its children were counted in
                    sp = -1;
                    continue;
                }
                no_split = true;
            }
        }

        const int ply = base_ply + sp;
        const int side = (sp & 1) ? -base_side : base_side;
        if (fresh) {
            fresh = false;
            // exact shortcuts; need: no node below can be left without a legal move
            if (R.shortcuts && __popc(emp) > R.max_depth - ply) {
                uint32_t win, sui;
                ab_move_classes(side > 0 ? x : o, kFull25 & ~(x | o), win, sui);
                const int best = side * (kWin - (ply + 1));        // the mover completes a 4-line now
                if (win) {
                    ret = best; returning = true; sp--; restore = true;
                } else {
                    if (side > 0) beta = min(beta, best); else alpha = max(alpha, best);   // mate distance
                    if (sui) {
                        value = -best;                            // what a 3-line move is worth
                        if (side > 0) alpha = max(alpha, value); else beta = min(beta, value);
                        moves &= ~order_mask(sui);
                    }
                    if (beta <= alpha) {
                        if (side > 0) ret = value >= beta ? value : beta;
                        else ret = value <= alpha ? value : alpha;
                        returning = true; sp--; restore = true;
                    }
                }
            }
        }
        if (returning) {
            // resolved by the shortcuts above
        } else if (moves == 0) {
            ret = value; returning = true;            // node done (also: a node with no legal move)
            sp--; restore = true;
        } else {
            // the killer of this depth first (stale or never-set words are harmless: only legal moves pass the mask)
            uint32_t pref = R.killers ? (moves & SQ_FR(sp, AbFrameWords<DIALECT>::saved)) : 0u;
            if (R.killers && R.two_killers && !pref) pref = moves & SQ_FR(sp, AbFrameWords<DIALECT>::saved + 1);   // the one before
            // then (delayed dialects, where the walk of the pending move is at hand) the holes of the opponent's
            // three-of-four quads: a move there blocks a four
            if (kDelayed && R.killers >= 2) { const uint32_t blk = moves & fe1; if (R.killers == 2) { if (!pref) pref = blk; } else if (blk) pref = blk; }
            const uint32_t from = pref ? pref : moves;
            const uint32_t kbit = from & (0u - from);
            cur_kbit = kbit;
            moves ^= kbit;
            const int kidx = __ffs(kbit) - 1;
            const int cell = kOrdered ? T.order[kidx] : kidx;
            const uint32_t bit = 1u << cell;
            const uint32_t emp_after = emp & ~kbit;
            const uint32_t xs = x | (side > 0 ? bit : 0u), os = o | (side > 0 ? 0u : bit);
            const int cply = ply + 1;
            evals++;
            int d = 0;
            if (kDelayed) {
                // the pending move, evaluated with this child's mark on the board
                const int cnt = __popc(fe1 & ~bit) + __popc(fe2 & ~bit);
                d = -side * (30 * cnt + R.scores[fpc]);
                if (kPenalties) d += eval_penalties(T, side > 0 ? os : xs, side > 0 ? xs : os, fpc, -side);
                if (AbTraits<DIALECT>::delayed_sum) d += carried;
            }
            if (kDelayed && emp_after == 0) {
                ret = side > 0 ? R.sentinel : -R.sentinel;            // the child (other side) has no move
                returning = true;
            } else {
                const AbWalk w = ab_walk(W, side > 0 ? xs : os, side > 0 ? os : xs, cell);
                if (w.four) { ret = side * (kWin - cply); returning = true; }
                else if (w.three) { ret = -side * (kWin - cply); returning = true; }
                else if (kDelayed) {
                    if (cply >= R.max_depth) {
                        // the child stops at its first empty cell, which it marks before evaluating
                        const uint32_t j0 = emp_after & (0u - emp_after);          // order space == cell space
                        const int cnt = __popc(w.e1 & ~j0) + __popc(w.e2 & ~j0);
                        ret = side * (30 * cnt + R.scores[cell]) + d;
                        if (kPenalties) ret += eval_penalties(T, side > 0 ? xs : os, (side > 0 ? os : xs) | j0, cell, side);
                        evals++;
                        returning = true;
                    } else if (!kPenalties && R.fast_horizon && cply == R.max_depth - 1 && (emp_after & (emp_after - 1u))) {
                        // the child is a horizon-1 node that cannot run out of cells: its value in one shot
                        const uint32_t own2 = side > 0 ? os : xs, opp2 = side > 0 ? xs : os;      // the child's mover is the other side
                        const uint32_t f0 = emp_after & (0u - emp_after);
                        int konst = -30 * (__popc(w.e1) + __popc(w.e2)) - R.scores[cell];
                        if (AbTraits<DIALECT>::delayed_sum) konst -= side * d;
                        evals += 3;          // about three table walks of instructions
                        if (kDefer) { d_bit = bit; d_kbit = kbit; d_e1 = w.e1; d_e2 = w.e2; d_konst = konst; pend_h = true; }
                        else {
                            const int rel = horizon_node_value<true>(own2, opp2, emp_after, kWin - (cply + 1), konst, w.e1, w.e2, R.scores[__ffs(f0) - 1], R.hz);
                            ret = -side * rel;
                            returning = true;
                        }
                    }
                } else {
                    int dv = side * (30 * (__popc(w.e1) + __popc(w.e2)) + R.scores[cell]);
                    if (kPenalties) dv += eval_penalties(T, side > 0 ? xs : os, side > 0 ? os : xs, cell, side);
                    if (cply == R.max_depth) { ret = dv + carried; returning = true; }
                    else if (emp_after == 0) { ret = side > 0 ? R.sentinel : -R.sentinel; returning = true; }
                    else if (!kPenalties && R.fast_horizon && cply == R.max_depth - 1 && (emp_after & (emp_after - 1u))) {
                        const uint32_t own2 = side > 0 ? os : xs, opp2 = side > 0 ? xs : os;
                        evals += 2;
                        if (kDefer) { d_bit = bit; d_kbit = kbit; d_konst = -side * (dv + carried); pend_h = true; }
                        else {
                            const int rel = horizon_node_value<false>(own2, opp2, kFull25 & ~(xs | os), kWin - (cply + 1), -side * (dv + carried), 0u, 0u, 0, R.hz);
                            ret = -side * rel;
                            returning = true;
                        }
                    }
                    else d = dv + carried;
                }
                if (!returning && !pend_h) {
                    // push: the register frame goes to shared memory, the child becomes the register frame
                    SQ_CHECK(sp >= 0 && sp < n_frames, 15, sp, n_frames);
                    SQ_FR(sp, 0) = moves | ((uint32_t)kidx << 25);
                    SQ_FR(sp, 1) = ((uint32_t)value & 0xffffu) | ((uint32_t)alpha << 16);
                    SQ_FR(sp, 2) = ((uint32_t)beta & 0xffffu) | ((uint32_t)carried << 16);
                    if (kDelayed) { SQ_FR(sp, 3) = fe1 | ((uint32_t)fpc << 25); SQ_FR(sp, 4) = fe2; }
                    x = xs; o = os; emp = emp_after;
                    sp++;
                    moves = emp;
                    value = side > 0 ? R.sentinel : -R.sentinel;      // the child is the other side
                    carried = d;
                    if (kDelayed) { fe1 = w.e1; fe2 = w.e2; fpc = cell; }
                    fresh = true;
                }
            }
        }
        }   // !served
        // hand `ret` to frame sp; a cutoff there hands that frame's value further up
        while (returning) {
            if (sp < 0) {
                atomicAdd(&R.node_count[group], evals);
                ab_complete(R, unit, ret);
                if (QUEUED) atomicSub(ab_cnt(R, 7), 1u);
                break;
            }
            const int s = (sp & 1) ? -base_side : base_side;
            if (restore) {
                // a pushed child came back: restore the frame and take its move off the board
                const uint32_t w0 = SQ_FR(sp, 0), w1 = SQ_FR(sp, 1), w2 = SQ_FR(sp, 2);
                moves = w0 & kFull25;
                value = (int)(short)(w1 & 0xffffu); alpha = (int)(short)(w1 >> 16);
                beta = (int)(short)(w2 & 0xffffu); carried = (int)(short)(w2 >> 16);
                if (kDelayed) {
                    const uint32_t w3 = SQ_FR(sp, 3);
                    fe1 = w3 & kFull25; fpc = (int)(w3 >> 25); fe2 = SQ_FR(sp, 4);
                }
                const int kidx = (int)(w0 >> 25);
                const uint32_t bit = 1u << (kOrdered ? T.order[kidx] : kidx);
                if (s > 0) x &= ~bit; else o &= ~bit;
                emp |= 1u << kidx;
                cur_kbit = 1u << kidx;
            }
            if (s > 0) { value = max(value, ret); alpha = max(alpha, value); }
            else { value = min(value, ret); beta = min(beta, value); }
            if (sp == 0) ab_bounds(R, unit, alpha, beta);        // siblings may have tightened the window
            if (beta <= alpha) {
                if (R.killers && cur_kbit) {                                                     // this child cut the node off
                    const uint32_t k1 = SQ_FR(sp, AbFrameWords<DIALECT>::saved);
                    if (k1 != cur_kbit) { SQ_FR(sp, AbFrameWords<DIALECT>::saved + 1) = k1; SQ_FR(sp, AbFrameWords<DIALECT>::saved) = cur_kbit; }
                }
                ret = value; sp--; restore = true; cur_kbit = 0;
            } else returning = false;
        }
    }
    if (QUEUED) { atomicAdd(ab_cnt(R, 10), (unsigned)(busy_turns >> 10)); atomicAdd(ab_cnt(R, 11), (unsigned)(idle_turns >> 10)); }
#undef SQ_FR
}

// ---- host orchestration ---------------------------------------------------------------
template <int DIALECT>
inline cudaError_t ab_configure_dialect(int max_smem) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(ab_dfs_kernel<DIALECT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(ab_dfs_kernel<DIALECT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

inline cudaError_t ab_configure(int* blocks_per_sm) {
    int b1 = 0;
    const int max_smem = (kAbMaxDepth + 2) * 7 * kAbThreads * 4;      // 5 saved words + two killer words
    cudaError_t e;
    if ((e = ab_configure_dialect<1>(max_smem)) != cudaSuccess) return e;
    if ((e = ab_configure_dialect<2>(max_smem)) != cudaSuccess) return e;
    if ((e = ab_configure_dialect<3>(max_smem)) != cudaSuccess) return e;
    if ((e = ab_configure_dialect<5>(max_smem)) != cudaSuccess) return e;
    if ((e = ab_configure_dialect<6>(max_smem)) != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, ab_dfs_kernel<1, false>, kAbThreads, 10 * 7 * kAbThreads * 4);
    if (e != cudaSuccess) return e;
    *blocks_per_sm = b1 < 1 ? 1 : b1;
    return cudaSuccess;
}

// Evaluator calls a unit may spend in one pass before it is split (SQ_AB_BUDGET overrides).
inline unsigned long long ab_budget() {
    const char* env = getenv("SQ_AB_BUDGET");
    if (env && *env) return strtoull(env, nullptr, 10);
    return 0;   // 0: chosen per pass from the number of units
}

// Static pre-split: plies to expand breadth-first under a level-0 node before the first pass
// (skips passes that would certainly abort).  Override with SQ_AB_SPLIT=<0..3>.
inline int ab_choose_split(int empties, int rem) {
    const char* env = getenv("SQ_AB_SPLIT");
    int split = (env && *env) ? atoi(env) : 0;     // default 0: the adaptive passes do the splitting
    if (split < 0) split = 0;
    if (split > rem - 1) split = rem - 1 < 0 ? 0 : rem - 1;
    if (split > empties - 1) split = empties - 1 < 0 ? 0 : empties - 1;
    if (split > 3) split = 3;
    return split;
}

struct AbWork {   // device allocations of one chunk
    AbNode* nodes = nullptr; int* value = nullptr; int* pending = nullptr; int* units = nullptr;
    unsigned int* counters = nullptr; int8_t* split = nullptr;
    int* aborted = nullptr; int* kids = nullptr; int* held = nullptr; int* released = nullptr;
};

inline void ab_free(AbWork& w, cudaStream_t st) {
    if (w.nodes) cudaFreeAsync(w.nodes, st);
    if (w.value) cudaFreeAsync(w.value, st);
    if (w.pending) cudaFreeAsync(w.pending, st);
    if (w.units) cudaFreeAsync(w.units, st);
    if (w.counters) cudaFreeAsync(w.counters, st);
    if (w.split) cudaFreeAsync(w.split, st);
    if (w.aborted) cudaFreeAsync(w.aborted, st);
    if (w.kids) cudaFreeAsync(w.kids, st);
    if (w.held) cudaFreeAsync(w.held, st);
    if (w.released) cudaFreeAsync(w.released, st);
    w = AbWork();
}

#define SQ_AB_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ab_free(w, st); return e_; } } while (0)

// Node pool: starts at what the level-0 nodes need (at least kAbPoolStart slots) and doubles on demand up to
// the ceiling (ab_pool_nodes()); the arrays move, the node indices stay valid.
inline cudaError_t ab_alloc(AbWork& w, long long capacity, cudaStream_t st) {
    cudaError_t e;
    if ((e = cudaMallocAsync(&w.nodes, sizeof(AbNode) * capacity, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.value, sizeof(int) * capacity, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.pending, sizeof(int) * capacity, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.units, sizeof(int) * capacity * 2, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.aborted, sizeof(int) * capacity, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.kids, sizeof(int) * capacity, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.held, sizeof(int) * capacity, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&w.released, sizeof(int) * capacity, st)) != cudaSuccess) return e;
    return cudaSuccess;
}

inline cudaError_t ab_grow(AbWork& w, long long used, int n_aborted, int n_units, long long new_capacity, cudaStream_t st) {
    AbWork g;
    cudaError_t e = ab_alloc(g, new_capacity, st);
    if (e != cudaSuccess) { ab_free(g, st); return e; }
    cudaMemcpyAsync(g.nodes, w.nodes, sizeof(AbNode) * used, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(g.value, w.value, sizeof(int) * used, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(g.pending, w.pending, sizeof(int) * used, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(g.kids, w.kids, sizeof(int) * used, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(g.held, w.held, sizeof(int) * used, cudaMemcpyDeviceToDevice, st);
    if (n_units > 0) cudaMemcpyAsync(g.units, w.units, sizeof(int) * n_units, cudaMemcpyDeviceToDevice, st);
    e = cudaMemcpyAsync(g.aborted, w.aborted, sizeof(int) * n_aborted, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { ab_free(g, st); return e; }
    g.counters = w.counters; g.split = w.split;
    w.counters = nullptr; w.split = nullptr;
    ab_free(w, st);
    w = g;
    return cudaSuccess;
}

constexpr long long kAbPoolStart = 1ll << 20;

// the pool size the last search of this process ended with: the next one starts there instead of growing again
inline std::atomic<long long>& ab_hint() { static std::atomic<long long> h{0}; return h; }
inline void ab_hint_store(long long capacity) { ab_hint().store(capacity); }

template <int DIALECT>
cudaError_t ab_run_chunk(AbRun R, AbWork& w, const void* d_in, bool roots, int n, const int8_t* h_split,
                         long long capacity, long long ceiling, int grid_cap, int min_ply, cudaStream_t st, uint64_t* launches) {
    SQ_AB_CK(ab_alloc(w, capacity, st));
    SQ_AB_CK(cudaMallocAsync(&w.counters, sizeof(unsigned int) * 8, st));
    SQ_AB_CK(cudaMallocAsync(&w.split, n, st));
    {
        const unsigned int init8[8] = {0, 0, 0, 0, 0, 0, 127, 0};
        SQ_AB_CK(cudaMemcpyAsync(w.counters, init8, sizeof init8, cudaMemcpyHostToDevice, st));
    }
    SQ_AB_CK(cudaMemcpyAsync(w.split, h_split, n, cudaMemcpyHostToDevice, st));
    auto bind = [&] {
        R.nodes = w.nodes; R.value = w.value; R.pending = w.pending; R.units = w.units; R.aborted = w.aborted;
        R.kids = w.kids; R.held = w.held; R.released = w.released;
        R.counters = w.counters; R.cstride = 1; R.node_capacity = (int)capacity; R.queue_capacity = (int)(2 * capacity);
    };
    bind();
    if (roots) {
        int threads = n * 25;
        ab_roots_kernel<DIALECT><<<(threads + kAbThreads - 1) / kAbThreads, kAbThreads, 0, st>>>(
            R, (const sq_board*)d_in, w.split, n);
    } else {
        ab_subtree_nodes_kernel<<<(n + kAbThreads - 1) / kAbThreads, kAbThreads, 0, st>>>(
            R, (const sq_subtree*)d_in, w.split, n);
    }
    (*launches)++;
    SQ_AB_CK(cudaGetLastError());
    unsigned int h_counters[7] = {0, 0, 0, 0, 0, 0, 127};
    int lo = 0;
    for (int level = 0; level < 3; level++) {        // static pre-split (usually none)
        SQ_AB_CK(cudaMemcpyAsync(h_counters, w.counters, sizeof h_counters, cudaMemcpyDeviceToHost, st));
        SQ_AB_CK(cudaStreamSynchronize(st));
        int hi = (int)h_counters[0];
        if (hi <= lo) break;
        ab_expand_kernel<DIALECT><<<(hi - lo + kAbThreads - 1) / kAbThreads, kAbThreads, 0, st>>>(R, nullptr, lo, hi - lo);
        (*launches)++;
        SQ_AB_CK(cudaGetLastError());
        lo = hi;
    }
    // Adaptive passes: every unit gets `budget` evaluator calls; the ones that need more are split one ply
    // deeper.  Of a split unit's children only the ELDEST becomes a unit of the next pass; the younger ones wait
    // until it has finished (ab_complete sets them free) and then start with its value as their bound -- young
    // brothers wait, at pass granularity.  Small subtrees keep sequential pruning, big ones are spread over the
    // machine one generation at a time, with bounds that are nearly as good as the sequential search's.
    const bool trace = getenv("SQ_AB_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what, int a, int b) {
        if (!trace) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ab] %-8s %9d %9d  %8.3f ms\n", what, a, b, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    lap("setup", n, (int)capacity);
    { const char* e = getenv("SQ_AB_YBW"); const int k = e && *e ? atoi(e) : 0; if (k >= 1) R.ybw_below_ply = min_ply + k; }   // SQ_AB_YBW=all: everywhere; =k: the top k plies
    const unsigned long long fixed_budget = ab_budget();
    const bool deep_few = R.max_depth - min_ply >= 10 && (long long)n * (roots ? 25 : 1) <= 4096;
    bool unlimited = false;
    int passes = 0;
    for (;;) {
        SQ_AB_CK(cudaMemcpyAsync(h_counters, w.counters, sizeof h_counters, cudaMemcpyDeviceToHost, st));
        SQ_AB_CK(cudaStreamSynchronize(st));
        if (h_counters[4]) { ab_free(w, st); return cudaErrorMemoryAllocation; }     // a node did not fit: never a silent short tree
        int n_units = (int)h_counters[1];
        const int n_released = (int)h_counters[5];
        if (n_released > 0) {                          // younger brothers set free during the last pass / expansion
            SQ_AB_CK(cudaMemcpyAsync(w.units + n_units, w.released, sizeof(int) * n_released, cudaMemcpyDeviceToDevice, st));
            n_units += n_released;
        }
        if (n_units == 0) break;
        // (units re-queued without a budget are not in the ply count: size for the shallowest possible unit then)
        const int unit_ply = (unlimited || (int)h_counters[6] > 126) ? min_ply : (int)h_counters[6];
        const unsigned int zero7[7] = {h_counters[0], 0, 0, 0, 0, 0, 127};  // n_nodes stays; units, queue head, aborted, released, min ply restart
        SQ_AB_CK(cudaMemcpyAsync(w.counters, zero7, sizeof zero7, cudaMemcpyHostToDevice, st));
        // few units: split early (most lanes are idle anyway); many: let lanes run longer
        // Deep searches from a few level-0 nodes split until the pool is full, and their last pass has to finish whatever is
        // left without a budget: the finer the units are by then, the shorter that pass's tail -- a small fixed budget
        // (measured: opening2 -d 12 3.36 s -> 1.98 s, opening3 -d 12 2.26 s -> 1.46 s; profiles/r02_alphabeta_modes_timing.txt).
        unsigned long long budget = fixed_budget ? fixed_budget : deep_few ? 64 : n_units < (1 << 15) ? 256 : n_units < (1 << 18) ? 1024 : 4096;
        if (unlimited) budget = ~0ull;
        // frames a unit can push: one per ply from its own down to max_depth - 1
        int n_frames = R.max_depth - unit_ply + 1;
        if (n_frames > kAbMaxDepth + 2) n_frames = kAbMaxDepth + 2;
        if (n_frames < 2) n_frames = 2;
        long long blocks = ((long long)n_units + kAbThreads - 1) / kAbThreads;
        if (blocks > grid_cap) blocks = grid_cap;
        ab_dfs_kernel<DIALECT, false><<<(unsigned)blocks, kAbThreads, (size_t)n_frames * AbFrameWords<DIALECT>::n * kAbThreads * 4, st>>>(
            R, n_units, budget, n_frames);
        (*launches)++;
        passes++;
        SQ_AB_CK(cudaGetLastError());
        SQ_AB_CK(cudaMemcpyAsync(h_counters, w.counters, sizeof h_counters, cudaMemcpyDeviceToHost, st));
        SQ_AB_CK(cudaStreamSynchronize(st));
        const int n_aborted = (int)h_counters[3];
        lap("dfs", n_units, n_aborted);
        if (n_aborted == 0) continue;                 // nothing to split; units set free by this pass are picked up at the top
        const long long want = (long long)h_counters[0] + 25ll * n_aborted;
        if (want > capacity && capacity < ceiling) {      // grow the pool: the arrays move, indices stay valid
            long long bigger = capacity;
            while (bigger < want && bigger < ceiling) bigger *= 2;
            if (bigger > ceiling) bigger = ceiling;
            // the units list is empty here (consumed by the pass); released[] is re-read from the old arrays below
            int* old_released = w.released; w.released = nullptr;
            if (ab_grow(w, (long long)h_counters[0], n_aborted, 0, bigger, st) == cudaSuccess) {
                if (h_counters[5]) cudaMemcpyAsync(w.released, old_released, sizeof(int) * h_counters[5], cudaMemcpyDeviceToDevice, st);
                cudaFreeAsync(old_released, st);
                capacity = bigger; bind(); lap("grow", (int)(capacity >> 10), 0);
            } else { cudaGetLastError(); w.released = old_released; }   // no memory for a bigger pool: carry on with what there is
        }
        if (want > capacity) {
            // no room to split further: finish the stragglers without a budget (they become units again)
            SQ_AB_CK(cudaMemcpyAsync(w.units, w.aborted, sizeof(int) * n_aborted, cudaMemcpyDeviceToDevice, st));
            const unsigned int nu = (unsigned)n_aborted;
            SQ_AB_CK(cudaMemcpyAsync(w.counters + 1, &nu, sizeof nu, cudaMemcpyHostToDevice, st));
            unlimited = true;
            continue;
        }
        // each split goes at least one ply deeper than the shallowest unit so far; with brothers waiting the deepest
        // chain grows by one ply per split, which only makes this frame bound generous
        ab_expand_kernel<DIALECT><<<(n_aborted + kAbThreads - 1) / kAbThreads, kAbThreads, 0, st>>>(R, w.aborted, 0, n_aborted);
        (*launches)++;
        SQ_AB_CK(cudaGetLastError());
    }
    if (trace) fprintf(stderr, "[ab] %d passes, pool %lld of %lld slots\n", passes, (long long)h_counters[0], capacity);
    ab_hint_store(capacity);
    ab_free(w, st);
    return cudaSuccess;
}

// Queue-driven search of one chunk: the root kernel makes the level-0 nodes and queues them, ONE persistent launch of
// ab_dfs_kernel<., true> searches them to the end (splitting on the device as lanes run dry), the host only reads
// the flags afterwards.  The node pool cannot grow while the kernel runs: it is sized from the request and from what
// the last search of this process needed; if it does fill up, units simply stop being split (correct, less parallel)
// and the next search starts with a bigger pool.
constexpr long long kAbQueuePoolStart = 8ll << 20;

template <int DIALECT>
cudaError_t ab_run_queue(AbRun R, AbWork& w, const void* d_in, bool roots, int n, const int8_t* h_split,
                         long long capacity, int sm_count, int min_ply, cudaStream_t st, uint64_t* launches) {
    (void)h_split;
    SQ_AB_CK(ab_alloc(w, capacity, st));
    SQ_AB_CK(cudaMallocAsync(&w.counters, sizeof(unsigned int) * 16 * 32, st));
    SQ_AB_CK(cudaMallocAsync(&w.split, n, st));
    SQ_AB_CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned int) * 16 * 32, st));
    SQ_AB_CK(cudaMemsetAsync(w.split, 0, n, st));
    SQ_AB_CK(cudaMemsetAsync(w.units, 0xff, sizeof(int) * capacity * 2, st));      // every queue slot: empty
    R.nodes = w.nodes; R.value = w.value; R.pending = w.pending; R.units = w.units; R.aborted = w.aborted;
    R.kids = w.kids; R.held = w.held; R.released = w.released;
    R.counters = w.counters; R.cstride = 32; R.node_capacity = (int)capacity; R.queue_capacity = (int)(2 * capacity);
    R.queued = 1;
    R.ybw = getenv("SQ_AB_NO_YBW") == nullptr;
    { const char* e = getenv("SQ_AB_MIN_BUDGET"); R.min_budget = e && *e ? (unsigned)atoi(e) : 384u; }
    R.gen_scaling = getenv("SQ_AB_GEN_SCALING") != nullptr;
    R.stall_cycles = 8000000000ll;                  // ~4 s without any working lane's heartbeat: a stuck queue, not a long search
    if (roots) {
        int threads = n * 25;
        ab_roots_kernel<DIALECT><<<(threads + kAbThreads - 1) / kAbThreads, kAbThreads, 0, st>>>(R, (const sq_board*)d_in, w.split, n);
    } else {
        ab_subtree_nodes_kernel<<<(n + kAbThreads - 1) / kAbThreads, kAbThreads, 0, st>>>(R, (const sq_subtree*)d_in, w.split, n);
    }
    (*launches)++;
    SQ_AB_CK(cudaGetLastError());
    int n_frames = R.max_depth - min_ply + 1;       // units sit at any ply from the level-0 nodes down
    if (n_frames > kAbMaxDepth + 2) n_frames = kAbMaxDepth + 2;
    if (n_frames < 2) n_frames = 2;
    const size_t smem = (size_t)n_frames * AbFrameWords<DIALECT>::n * kAbThreads * 4;
    int per_sm = 1;
    SQ_AB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ab_dfs_kernel<DIALECT, true>, kAbThreads, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)sm_count * per_sm;             // the whole machine, every block resident: lanes that find no unit
                                                                 // wait for the ones the working lanes split off
    if (R.max_depth - min_ply <= 3) {                            // shallow searches have no work to spread
        const long long most = ((long long)n * 25 + kAbThreads - 1) / kAbThreads;
        if (blocks > most) blocks = most < 1 ? 1 : most;
    }
    const bool trace = getenv("SQ_AB_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    ab_dfs_kernel<DIALECT, true><<<(unsigned)blocks, kAbThreads, smem, st>>>(R, 0, 0ull, n_frames);
    (*launches)++;
    SQ_AB_CK(cudaGetLastError());
    unsigned int h_padded[16 * 32], h_counters[16];
    SQ_AB_CK(cudaMemcpyAsync(h_padded, w.counters, sizeof h_padded, cudaMemcpyDeviceToHost, st));
    SQ_AB_CK(cudaStreamSynchronize(st));
    for (int k = 0; k < 16; k++) h_counters[k] = h_padded[k * 32];
    if (trace)
        fprintf(stderr, "[ab] queue: %d level-0 inputs, %u nodes of %lld, %u units queued, %u tickets drawn, %u alive at the end, %u blocks, %.3f ms; "
                "%u split attempts, lane turns busy %.3g idle %.3g\n",
                n, h_counters[0], capacity, h_counters[1], h_counters[2], h_counters[7], (unsigned)blocks,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), h_counters[12],
                1024.0 * h_counters[10], 1024.0 * h_counters[11]);
    ab_free(w, st);
    if (h_counters[8] || h_counters[7]) return cudaErrorLaunchTimeout;   // the watchdog fired / units left over: never a silent wrong answer
    // remember what this search needed: twice the nodes it used, so that the next one has room to split freely
    ab_hint_store(std::max<long long>(2ll * h_counters[0], kAbQueuePoolStart));
    return cudaSuccess;
}

constexpr long long kAbChunkNodesDefault = 64ll << 20;   // ceiling of node slots per chunk: level-0 nodes + what adaptive splitting may draw (~2.5 GB when all used)

inline long long ab_pool_nodes() {      // SQ_AB_POOL_NODES shrinks the pool (tests of the no-room path)
    const char* env = getenv("SQ_AB_POOL_NODES");
    long long v = (env && *env) ? atoll(env) : kAbChunkNodesDefault;
    return v < 4096 ? 4096 : v;
}

inline cudaError_t ab_dispatch(int dialect, AbRun R, AbWork& w, const void* d_in, bool roots, int n,
                               const int8_t* h_split, long long needed, int grid_cap, int min_ply,
                               cudaStream_t st, uint64_t* launches) {
    // ceiling: SQ_AB_POOL_NODES (default 64 Mi slots, ~2.5 GB), never below what the level-0 nodes need;
    // the run starts with max(needed, kAbPoolStart) slots and grows towards the ceiling only if it splits that far
    const long long ceiling = std::max(needed + 16, ab_pool_nodes());
    if (getenv("SQ_AB_QUEUE") && !getenv("SQ_AB_PASSES") && !getenv("SQ_AB_BUDGET") && !getenv("SQ_AB_SPLIT")) {
        // queue-driven search: opt-in (SQ_AB_QUEUE=1).  Measured in round 2 (profiles/r02_alphabeta_queue_driven.md): bit-exact,
        // an eighth of the evaluator calls on deep searches when brothers wait, but the lanes stand idle most of the time
        // (units reach them one generation at a time), so the pass-driven search below is still the faster one and the default.
        const long long start_q = R.max_depth - min_ply >= 9 ? 4 * kAbQueuePoolStart : kAbQueuePoolStart;     // deep searches split a lot
        const long long cap_q = std::min(ceiling, std::max(std::max(2 * needed + 16, start_q), ab_hint().load()));
        int dev_id = 0, sm_count = 148;
        if (cudaGetDevice(&dev_id) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev_id) != cudaSuccess) return cudaGetLastError();
        switch (dialect) {
        case 1: return ab_run_queue<1>(R, w, d_in, roots, n, h_split, cap_q, sm_count, min_ply, st, launches);
        case 2: return ab_run_queue<2>(R, w, d_in, roots, n, h_split, cap_q, sm_count, min_ply, st, launches);
        case 5: return ab_run_queue<5>(R, w, d_in, roots, n, h_split, cap_q, sm_count, min_ply, st, launches);
        case 6: return ab_run_queue<6>(R, w, d_in, roots, n, h_split, cap_q, sm_count, min_ply, st, launches);
        default: return ab_run_queue<3>(R, w, d_in, roots, n, h_split, cap_q, sm_count, min_ply, st, launches);
        }
    }
    const long long capacity = std::min(ceiling, std::max(std::max(needed + 16, kAbPoolStart), ab_hint().load()));
    switch (dialect) {
    case 1: return ab_run_chunk<1>(R, w, d_in, roots, n, h_split, capacity, ceiling, grid_cap, min_ply, st, launches);
    case 2: return ab_run_chunk<2>(R, w, d_in, roots, n, h_split, capacity, ceiling, grid_cap, min_ply, st, launches);
    case 5: return ab_run_chunk<5>(R, w, d_in, roots, n, h_split, capacity, ceiling, grid_cap, min_ply, st, launches);
    case 6: return ab_run_chunk<6>(R, w, d_in, roots, n, h_split, capacity, ceiling, grid_cap, min_ply, st, launches);
    default: return ab_run_chunk<3>(R, w, d_in, roots, n, h_split, capacity, ceiling, grid_cap, min_ply, st, launches);
    }
}

inline void ab_fill_run(AbRun& R, int dialect, int max_depth, const int8_t scores[25], long long max_abs_carried = 0) {
    memset(&R, 0, sizeof R);
    R.max_depth = max_depth;       // squavathr's own decrement (squavathr.go:239) is applied by ab_root_search
    const bool opening = dialect == 3 || dialect == 4;
    R.sentinel = dialect == 4 ? kWin : 2 * kWin;
    R.win_lo = opening ? -kWin : -2 * kWin;
    R.win_hi = opening ? kWin : 2 * kWin;
    memcpy(R.scores, scores, 25);
    // The shortcuts assume a win (>= 10000 - max_depth in magnitude) outranks every heuristic
    // value a leaf can return.  One evaluation contributes at most 8 quads x 30, the bias and
    // (dialects 2 and 5) three -100 terms; dialects 2-4 sum it over the plies, dialects 1 and
    // 5 add two consecutive evaluations; a caller-supplied carried value comes on top.
    int max_score = 0;
    for (int c = 0; c < 25; c++) max_score = std::max(max_score, std::abs((int)scores[c]));
    const long long per_eval = 240 + max_score + ((dialect == 2 || dialect == 5) ? 300 : 0);
    const bool cumulative = dialect == 2 || dialect == 3 || dialect == 4 || dialect == 6;
    const long long worst = (cumulative ? (long long)(max_depth + 1) * per_eval : 2 * per_eval) + max_abs_carried;
    R.shortcuts = worst < kWin - max_depth - 1 && !getenv("SQ_AB_NO_SHORTCUTS");
    build_horizon_tables(scores, R.hz);
    R.fast_horizon = R.shortcuts && R.hz.ok && !getenv("SQ_AB_NO_FAST_HORIZON");
    // measured (profiles/r02_alphabeta_ybw_passes.md): with a barrier between passes the brothers' wait costs more
    // time than the pruning it buys except in very deep searches, so it is opt-in for the pass-driven search
    R.ybw = getenv("SQ_AB_YBW") != nullptr;
    R.ybw_below_ply = 127;
    R.killers = getenv("SQ_AB_NO_KILLERS") == nullptr;
    R.two_killers = getenv("SQ_AB_ONE_KILLER") == nullptr;
    { const char* e = getenv("SQ_AB_ORDER"); if (R.killers && e && *e) R.killers = 1 + atoi(e); }   // 1: + blocking moves after the killer, 2: before it
    { const char* e = getenv("SQ_AB_DEFER"); R.defer_min = e && *e ? atoi(e) : 0; }
}

// worst-case node count below one level-0 node with e empties and `split` expanded plies
inline long long ab_node_bound(int e, int split) {
    long long total = 1, layer = 1;
    for (int k = 0; k < split; k++) { layer *= (e - k > 0 ? e - k : 0); total += layer; }
    return total;
}

// positions: h_pos (host copy, for sizing) and d_pos (device).  d_values [n][25], d_nodes [n].
inline cudaError_t ab_root_search(const sq_board* h_pos, const sq_board* d_pos, long long n, int dialect,
                                  int max_depth, const int8_t scores[25], int32_t* d_values,
                                  unsigned long long* d_nodes, int grid_cap, cudaStream_t st,
                                  uint64_t* launches) {
    AbRun R; AbWork w;
    ab_fill_run(R, dialect, dialect == 2 ? max_depth - 1 : max_depth, scores);   // squavathr.go:239
    cudaError_t e = cudaMemsetAsync(d_nodes, 0, sizeof(unsigned long long) * n, st);
    if (e != cudaSuccess) return e;
    std::vector<int8_t> split((size_t)n);
    const long long kAbChunkNodes = ab_pool_nodes();
    long long lo = 0;
    while (lo < n) {
        long long cap = 0, hi = lo;
        while (hi < n) {
            int empt = 25 - __builtin_popcount((h_pos[hi].x | h_pos[hi].o) & kFull25);
            int sp = ab_choose_split(empt - 1, R.max_depth - 1);
            long long need = (long long)empt * ab_node_bound(empt - 1, sp);
            if (hi > lo && cap + need > kAbChunkNodes / 16) break;
            split[hi] = (int8_t)sp; cap += need; hi++;
        }
        AbRun Rc = R;
        Rc.out = d_values + lo * 25;
        Rc.node_count = d_nodes + lo;
        e = ab_dispatch(dialect, Rc, w, d_pos + lo, true, (int)(hi - lo), split.data() + lo, cap, grid_cap, 1, st, launches);
        if (e != cudaSuccess) return e;
        lo = hi;
    }
    return cudaSuccess;
}

inline cudaError_t ab_subtree_search(const sq_subtree* h_st, const sq_subtree* d_st, long long n, int dialect,
                                     int max_depth, const int8_t scores[25], int32_t* d_value,
                                     unsigned long long* d_nodes, int grid_cap, cudaStream_t st,
                                     uint64_t* launches) {
    AbRun R; AbWork w;
    long long max_carried = 0;
    for (long long q = 0; q < n; q++) max_carried = std::max(max_carried, std::llabs((long long)h_st[q].carried));
    ab_fill_run(R, dialect, max_depth, scores, max_carried);
    cudaError_t e = cudaMemsetAsync(d_nodes, 0, sizeof(unsigned long long) * n, st);
    if (e != cudaSuccess) return e;
    std::vector<int8_t> split((size_t)n);
    const long long kAbChunkNodes = ab_pool_nodes();
    long long lo = 0;
    while (lo < n) {
        long long cap = 0, hi = lo;
        while (hi < n) {
            int empt = 25 - __builtin_popcount((h_st[hi].b.x | h_st[hi].b.o) & kFull25);
            int sp = ab_choose_split(empt, R.max_depth - h_st[hi].ply);
            long long need = ab_node_bound(empt, sp);
            if (hi > lo && cap + need > kAbChunkNodes / 16) break;
            split[hi] = (int8_t)sp; cap += need; hi++;
        }
        AbRun Rc = R;
        Rc.out = d_value + lo;
        Rc.node_count = d_nodes + lo;
        int min_ply = 127;
        for (long long q = lo; q < hi; q++) if (h_st[q].ply < min_ply) min_ply = h_st[q].ply;
        e = ab_dispatch(dialect, Rc, w, d_st + lo, false, (int)(hi - lo), split.data() + lo, cap, grid_cap, min_ply, st, launches);
        if (e != cudaSuccess) return e;
        lo = hi;
    }
    return cudaSuccess;
}

}  // namespace sq
