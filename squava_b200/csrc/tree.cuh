// tree.cuh -- the MCTS tree of squavam's UCT (src/mcts/mcts.go:140-200) kept in HBM.
//
// configs[2] of BASELINE.json is 10^9 one-playout iterations through ONE tree.  On the host that is a chain of
// dependent DRAM misses per walk (squava_b200/host/fast_tree.cpp: 6*10^6 walks/s on 16 cores); here the tree never
// leaves the device: a batch is three launches on one stream -- select (thousands of walks at once, every level a
// handful of L2/HBM reads), the playout kernel on the leaves the walks wrote, update -- and nothing crosses PCIe
// until the decision is read back.  Semantics are the host arena tree's throughput mode (fast_tree.cpp):
//   * a walk adds its visits on the way down (virtual loss, seen by every later step of every walk), its wins when the
//     batch's playouts are back;
//   * select: first maximum of UCB1 = w/v + UCTK*sqrt(f*ln(N)/v) over the children that are set up (mcts.go:218-238);
//     expand: one random untried move, claimed by an atomic AND (mcts.go:166-171); a node waiting for its first rollout
//     is a leaf for everybody who reaches it;
//   * a child block (one slot per empty cell of the parent, the child's move is its slot's rank among the empty cells)
//     is taken from a bump pointer by the first expansion of the parent; a block that lost the race is not used.
// The functions below are __host__ __device__: tests/native/tree_check.cu runs the same code sequentially on the
// host against a CPU rollout.
#pragma once
#include <cstdint>
#include "../../include/squava_b200.h"

#if defined(__CUDA_ARCH__)
#define SQT_DEVICE 1
#else
#define SQT_DEVICE 0
#endif

namespace sqt {

constexpr uint32_t kFull = 0x1ffffffu;
constexpr uint32_t kPending = 1u << 31;        // in `untried`: the node waits for its first rollout
constexpr int kPathLen = 27;                   // root + 25 moves + 1

struct Tree {
    uint32_t* visits;
    uint32_t* wins;
    uint32_t* first_child;                     // 0: no child block yet
    uint32_t* untried;                         // mask of the moves not expanded yet | kPending
    unsigned long long* next_free;             // bump pointer (slot 0 is never handed out: it means "none")
    unsigned long long cap;
    uint32_t* overflow;                        // set when the arena ran out: the search goes on without growing
    uint32_t* arrivals;                        // per node: walks of the current batch that stand on it (synchronised levels only)
};

struct Walks {                                 // what the select step leaves for the playout and update steps
    sq_board* boards;
    int8_t* to_move;
    uint32_t* leaf;
    uint8_t* fresh;                            // the leaf was made by this walk
    uint8_t* depth;                            // nodes on the path
    uint32_t* paths;                           // [walk][kPathLen]
    // synchronised levels: the walk's place among the walks on its node, and what it decided to do next
    uint32_t* rank;                            // kNoRank: holds no place
    uint8_t* act;                              // kActNone / kActDescend / kActExpand / kActDone
    uint8_t* arg;                              // child slot / cell
    int8_t* pjm;                               // player who just moved at the walk's node
};
constexpr uint32_t kNoRank = 0xffffffffu;
constexpr uint8_t kActNone = 0, kActDescend = 1, kActExpand = 2, kActDone = 3;

// ---- memory operations: L2-coherent on the device, plain on the host (sequential emulation) ----
__host__ __device__ inline uint32_t t_load(const uint32_t* p) {
#if SQT_DEVICE
    return __ldcg(p);
#else
    return *p;
#endif
}
__host__ __device__ inline void t_store(uint32_t* p, uint32_t v) {
#if SQT_DEVICE
    __stcg(p, v);
#else
    *p = v;
#endif
}
__host__ __device__ inline uint32_t t_add(uint32_t* p, uint32_t v) {
#if SQT_DEVICE
    return atomicAdd(p, v);
#else
    const uint32_t old = *p; *p = old + v; return old;
#endif
}
__host__ __device__ inline uint32_t t_and(uint32_t* p, uint32_t v) {
#if SQT_DEVICE
    return atomicAnd(p, v);
#else
    const uint32_t old = *p; *p = old & v; return old;
#endif
}
__host__ __device__ inline uint32_t t_or(uint32_t* p, uint32_t v) {
#if SQT_DEVICE
    return atomicOr(p, v);
#else
    const uint32_t old = *p; *p = old | v; return old;
#endif
}
__host__ __device__ inline uint32_t t_cas(uint32_t* p, uint32_t expect, uint32_t v) {
#if SQT_DEVICE
    return atomicCAS(p, expect, v);
#else
    const uint32_t old = *p; if (old == expect) *p = v; return old;
#endif
}
__host__ __device__ inline unsigned long long t_add64(unsigned long long* p, unsigned long long v) {
#if SQT_DEVICE
    return atomicAdd(p, v);
#else
    const unsigned long long old = *p; *p = old + v; return old;
#endif
}
__host__ __device__ inline void t_fence() {
#if SQT_DEVICE
    __threadfence();
#endif
}
__host__ __device__ inline int t_popc(uint32_t v) {
#if SQT_DEVICE
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
__host__ __device__ inline int t_nth_set_bit(uint32_t m, int n) {
    for (int i = 0; i < n; i++) m &= m - 1u;
#if SQT_DEVICE
    return __ffs(m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
__host__ __device__ inline uint64_t t_mix(uint64_t z) {           // splitmix64: which untried move a walk expands
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// Expand move `bit` of node idx (its claim already made): the child's slot, set up and counted.  0: the arena is full.
__host__ __device__ inline uint32_t tree_make_child(const Tree& T, uint32_t idx, uint32_t emp, uint32_t bit, uint32_t P) {
    const int slots = t_popc(emp);
    uint32_t fc = t_load(&T.first_child[idx]);
    if (!fc) {
        const unsigned long long base = t_add64(T.next_free, (unsigned long long)slots);
        if (base + (unsigned long long)slots >= T.cap) {      // arena full: hand the move back
            t_store(T.overflow, 1u);
            t_or(&T.untried[idx], bit);
            return 0u;
        }
        const uint32_t old = t_cas(&T.first_child[idx], 0u, (uint32_t)base);
        fc = old ? old : (uint32_t)base;                     // lost the race: that block is simply not used
    }
    const uint32_t child = fc + (uint32_t)t_popc(emp & (bit - 1u));
    t_store(&T.untried[child], kPending);
    t_fence();
    t_store(&T.visits[child], P);                            // nobody looks at a child before its visits are non-zero
    return child;
}

// The rest of a walk, from node idx (board x/o, pjm just moved, `depth` nodes on the path so far) to its leaf
// (mcts.go:153-171): every step sees the tree as it is at that moment.
__host__ __device__ inline void tree_walk_from(const Tree& T, uint32_t idx, uint32_t x, uint32_t o, int pjm, int depth, bool fresh, uint32_t P,
                                               float c2, uint64_t key, long long k, const Walks& W) {
    uint32_t* path = W.paths + (size_t)k * kPathLen;
    int waited = 0;
    while (!fresh) {
        const uint32_t u = t_load(&T.untried[idx]);
        if (u & kPending) break;                                 // waits for its first rollout: a leaf
        const uint32_t um = u & kFull;
        const uint32_t emp = kFull & ~(x | o);
        if (um) {                                                // expand
            const int cnt = t_popc(um);
            const int m = t_nth_set_bit(um, (int)(t_mix(key ^ ((uint64_t)k << 8) ^ (uint64_t)depth) % (uint64_t)cnt));
            const uint32_t bit = 1u << m;
            if (!(t_and(&T.untried[idx], ~bit) & bit)) continue; // another walk took it: look again (at most 25 times)
            const uint32_t child = tree_make_child(T, idx, emp, bit, P);
            if (!child) break;                                   // arena full: roll out from here
            pjm = -pjm;
            if (pjm > 0) x |= bit; else o |= bit;
            idx = child;
            path[depth++] = child;
            fresh = true;
            break;
        }
        const uint32_t fc = t_load(&T.first_child[idx]);
        if (!fc) break;                                          // a finished game: no moves
        const int slots = t_popc(emp);
        uint32_t seen = t_load(&T.visits[idx]);
        if (seen < (uint32_t)slots * P + 1u) seen = (uint32_t)slots * P + 1u;    // every child was made by a walk through the node
#if SQT_DEVICE
        const float lg = c2 * __logf((float)seen);
#else
        const float lg = c2 * __builtin_logf((float)seen);
#endif
        int best = -1;
        float best_score = -1.f;
        for (int j = 0; j < slots; j++) {
            const uint32_t v = t_load(&T.visits[fc + (uint32_t)j]);
            if (!v) continue;                                    // claimed, still being set up
            const float w = (float)t_load(&T.wins[fc + (uint32_t)j]);
            const float inv = 1.f / (float)v;
#if SQT_DEVICE
            const float s = w * inv + sqrtf(lg * inv);
#else
            const float s = w * inv + __builtin_sqrtf(lg * inv);
#endif
            if (s > best_score) { best_score = s; best = j; }
        }
        if (best < 0) {                                          // every child is being set up by another walk right now
            if (++waited < 64) {
#if SQT_DEVICE
                __nanosleep(100);
#endif
                continue;
            }
            break;                                               // roll out from here rather than wait
        }
        idx = fc + (uint32_t)best;
        t_add(&T.visits[idx], P);
        const uint32_t bit = 1u << t_nth_set_bit(emp, best);
        pjm = -pjm;
        if (pjm > 0) x |= bit; else o |= bit;
        path[depth++] = idx;
    }
    W.boards[k].x = x; W.boards[k].o = o;
    W.to_move[k] = (int8_t)-pjm;
    W.leaf[k] = idx;
    W.fresh[k] = fresh ? 1 : 0;
    W.depth[k] = (uint8_t)depth;
}

// One walk from the root to a leaf.  pjm: the player who just moved at the root (+1 X, -1 O).
// c2 = UCTK^2 * f (the exploration term is sqrt(c2 * ln N / v)).
__host__ __device__ inline void tree_walk(const Tree& T, uint32_t root, uint32_t rx, uint32_t ro, int root_pjm, uint32_t P, float c2,
                                          uint64_t key, long long k, const Walks& W, bool add_root_visit) {
    if (add_root_visit) t_add(&T.visits[root], P);
    W.paths[(size_t)k * kPathLen] = root;
    tree_walk_from(T, root, rx, ro, root_pjm, 1, false, P, c2, key, k, W);
}

// ---- synchronised levels ---------------------------------------------------------------------------------------------------
// Thousands of walks that reach a node in the same instant see the same statistics and would all take the same child: the
// virtual loss of the others is not there yet.  For the first levels of the tree -- where a batch's walks still travel
// together -- the descent therefore goes level by level, two steps per level over all walks of the batch:
//   arrive:  every walk carries out what it decided (descend: add its visits to the child; expand: make the child), then
//            takes a place (rank) among the walks standing on its node;
//   decide:  nothing in the tree changes during this step.  The K walks on a node read the same numbers and split
//            themselves among the children exactly as K sequential selections with virtual loss would: the first
//            popc(untried) ranks expand one untried move each; the others share out the ready children so that all
//            children that get walks end at the same UCB1 ("water level", found by bisection; the number of visits that
//            brings a child down to a level is a closed form), and a walk's rank picks its child from the running sums.
// Below the synchronised levels the walks have spread out and finish with tree_walk_from.

// arrive.  level 0: the walk starts at the root.
__host__ __device__ inline void tree_arrive(const Tree& T, int level, uint32_t root, uint32_t rx, uint32_t ro, int root_pjm, uint32_t P,
                                            long long k, const Walks& W, bool take_place) {
    uint32_t* path = W.paths + (size_t)k * kPathLen;
    if (level == 0) {
        W.boards[k].x = rx; W.boards[k].o = ro; W.pjm[k] = (int8_t)root_pjm; W.leaf[k] = root; W.depth[k] = 1; W.fresh[k] = 0;
        W.act[k] = kActNone; W.rank[k] = kNoRank;
        path[0] = root;
    } else {
        const uint32_t node = W.leaf[k];
        if (W.rank[k] == 0u) t_store(&T.arrivals[node], 0u);     // the first of the node's walks clears its counter for the next batch
        W.rank[k] = kNoRank;
        const uint8_t act = W.act[k];
        if (act == kActDone) return;
        uint32_t x = W.boards[k].x, o = W.boards[k].o;
        const uint32_t emp = kFull & ~(x | o);
        int pjm = W.pjm[k];
        if (act == kActDescend) {
            const uint32_t child = t_load(&T.first_child[node]) + (uint32_t)W.arg[k];
            t_add(&T.visits[child], P);
            const uint32_t bit = 1u << t_nth_set_bit(emp, (int)W.arg[k]);
            pjm = -pjm;
            if (pjm > 0) x |= bit; else o |= bit;
            path[W.depth[k]] = child;
            W.depth[k]++;
            W.leaf[k] = child;
        } else if (act == kActExpand) {
            const uint32_t bit = 1u << W.arg[k];
            W.act[k] = kActDone;
            if (!(t_and(&T.untried[node], ~bit) & bit)) return;  // (cannot happen: ranks hold different moves) -- roll out from the node
            const uint32_t child = tree_make_child(T, node, emp, bit, P);
            if (!child) return;                                  // arena full: roll out from the node
            pjm = -pjm;
            if (pjm > 0) x |= bit; else o |= bit;
            path[W.depth[k]] = child;
            W.depth[k]++;
            W.leaf[k] = child;
            W.fresh[k] = 1;
            W.boards[k].x = x; W.boards[k].o = o; W.pjm[k] = (int8_t)pjm;
            return;
        }
        W.boards[k].x = x; W.boards[k].o = o; W.pjm[k] = (int8_t)pjm;
        W.act[k] = kActNone;
    }
    if (take_place) W.rank[k] = t_add(&T.arrivals[W.leaf[k]], 1u);
}

// The decide step's arithmetic.  Single precision: the launch list (profiles/r02_gpu_tree_launches.json) had this step at
// 2.0 of the 2.6 ms of a batch in double precision (236 registers, FP64 square roots and divisions in a 40-step bisection
// that every walk on a node repeats).  What single precision costs is a child's share of a batch being off by a few
// visits in 65536 when the counts reach 10^8..10^9; every walk on a node still computes the SAME numbers, which is what
// makes the ranks partition the children.  -DSQT_REAL=double brings the old arithmetic back.
#ifndef SQT_REAL
#define SQT_REAL float
#endif
typedef SQT_REAL treal;
__host__ __device__ inline treal t_sqrt(treal a) {
#if SQT_DEVICE
    return (treal)sqrt(a);                       // (overload resolution picks sqrtf for float)
#else
    return sizeof(treal) == 4 ? (treal)__builtin_sqrtf((float)a) : (treal)__builtin_sqrt((double)a);
#endif
}
__host__ __device__ inline treal t_log(treal a) {
#if SQT_DEVICE
    return (treal)log(a);
#else
    return sizeof(treal) == 4 ? (treal)__builtin_logf((float)a) : (treal)__builtin_log((double)a);
#endif
}

// visits that bring a child (w wins, v visits) down to UCB1 level lambda; s = sqrt(c2 * ln N)
__host__ __device__ inline treal tree_visits_to_level(treal w, treal v, treal s, treal lambda) {
    // w/v' + s/sqrt(v') = lambda, t = 1/sqrt(v'):  w t^2 + s t - lambda = 0
    treal t;
    if (w > (treal)0) t = (-s + t_sqrt(s * s + (treal)4 * w * lambda)) / ((treal)2 * w);
    else t = lambda / s;
    const treal vp = (treal)1 / (t * t);
    return vp > v ? vp - v : (treal)0;
}

// decide.  Nothing in the tree is written here.
__host__ __device__ inline void tree_decide(const Tree& T, uint32_t P, float c2, uint64_t key, long long k, const Walks& W) {
    if (W.act[k] == kActDone || W.rank[k] == kNoRank) return;
    const uint32_t idx = W.leaf[k];
    const uint32_t rank = W.rank[k], K = t_load(&T.arrivals[idx]);
    const uint32_t u = t_load(&T.untried[idx]);
    if (u & kPending) { W.act[k] = kActDone; return; }          // waits for its first rollout: a leaf for all of them
    const uint32_t um = u & kFull;
    const uint32_t emp = kFull & ~(W.boards[k].x | W.boards[k].o);
    const uint32_t cnt = (uint32_t)t_popc(um);
    if (rank < cnt) {                                            // one untried move each, in an order that changes from batch to batch
        const uint32_t rot = (uint32_t)(t_mix(key ^ (uint64_t)idx) % (uint64_t)cnt);
        W.act[k] = kActExpand;
        W.arg[k] = (uint8_t)t_nth_set_bit(um, (int)((rank + rot) % cnt));
        return;
    }
    const uint32_t fc = t_load(&T.first_child[idx]);
    if (!fc) { W.act[k] = kActDone; return; }                   // a finished game, or no child is ready yet: roll out from here
    const int slots = t_popc(emp);
    treal v[25], w[25];
    int n_ready = 0, best = -1;
    treal top = (treal)-1;
    uint32_t seen = t_load(&T.visits[idx]);
    if (seen < (uint32_t)slots * P + 1u) seen = (uint32_t)slots * P + 1u;
    const treal s = t_sqrt((treal)c2 * t_log((treal)seen));
    for (int j = 0; j < slots; j++) {
        v[j] = (treal)t_load(&T.visits[fc + (uint32_t)j]);
        w[j] = (treal)t_load(&T.wins[fc + (uint32_t)j]);
        if (v[j] > (treal)0) {
            n_ready++;
            const treal f = w[j] / v[j] + s / t_sqrt(v[j]);
            if (f > top) { top = f; best = j; }
        }
    }
    if (!n_ready) { W.act[k] = kActDone; return; }
    const uint32_t Kp = K - (cnt < K ? cnt : K);                 // walks that select (this one among them)
    const uint32_t r = rank - cnt;
    W.act[k] = kActDescend;
    if (Kp <= 1u) { W.arg[k] = (uint8_t)best; return; }          // alone: the plain argmax
    // the level that Kp * P more visits fill the children up to
    const treal want = (treal)Kp * (treal)P;
    treal lo = (treal)0, hi = top;                               // total(hi) = 0, total(lo -> 0) -> infinity
    const int steps = sizeof(treal) == 4 ? 26 : 40;
    for (int it = 0; it < steps; it++) {
        const treal mid = (treal)0.5 * (lo + hi);
        treal total = (treal)0;
        for (int j = 0; j < slots; j++) if (v[j] > (treal)0) total += tree_visits_to_level(w[j], v[j], s, mid);
        if (total > want) lo = mid; else hi = mid;
    }
    const treal level = (treal)0.5 * (lo + hi);
    treal total = (treal)0;
    for (int j = 0; j < slots; j++) if (v[j] > (treal)0) total += tree_visits_to_level(w[j], v[j], s, level);
    if (!(total > (treal)0)) { W.arg[k] = (uint8_t)best; return; }
    // the walk's rank picks its child from the running sums (scaled so that they end at Kp exactly)
    const treal mine = ((treal)r + (treal)0.5) * total / (treal)Kp;
    treal run = (treal)0;
    int pick = best;
    for (int j = 0; j < slots; j++) {
        if (!(v[j] > (treal)0)) continue;
        run += tree_visits_to_level(w[j], v[j], s, level);
        if (mine < run) { pick = j; break; }
    }
    W.arg[k] = (uint8_t)pick;
}

// after the last synchronised level: carry out what was decided, then finish the walk on its own
__host__ __device__ inline void tree_finish(const Tree& T, int levels, uint32_t root, uint32_t rx, uint32_t ro, int root_pjm, uint32_t P, float c2,
                                            uint64_t key, long long k, const Walks& W) {
    if (levels == 0) { tree_walk(T, root, rx, ro, root_pjm, P, c2, key, k, W, false); return; }
    tree_arrive(T, levels, root, rx, ro, root_pjm, P, k, W, false);
    if (W.act[k] == kActDone) { W.to_move[k] = (int8_t)-W.pjm[k]; return; }     // its leaf record is what the arrive steps left
    tree_walk_from(T, W.leaf[k], W.boards[k].x, W.boards[k].o, W.pjm[k], W.depth[k], false, P, c2, key, k, W);
}

// Update (mcts.go:190-192, 268-271) of one walk.  counts: {X wins, O wins, draws, plies} of the leaf's playouts.
// The node at depth i of a path was moved INTO by the root's player-just-moved when i is even.  Returns what the root
// gets (the caller adds it: one atomic per warp instead of one per walk).
__host__ __device__ inline uint32_t tree_update(const Tree& T, int root_pjm, const unsigned long long* counts, long long k, const Walks& W) {
    const uint32_t* path = W.paths + (size_t)k * kPathLen;
    const int depth = W.depth[k];
    const uint32_t xw = (uint32_t)counts[0], ow = (uint32_t)counts[1];
    if (W.fresh[k]) {
        const uint32_t emp = kFull & ~(W.boards[k].x | W.boards[k].o);
        t_store(&T.untried[W.leaf[k]], counts[3] != 0 ? emp : 0u);          // no ply played: the game was over at the leaf
    }
    for (int i = 1; i < depth; i++) {
        const bool x_moved_in = (i & 1) ? root_pjm < 0 : root_pjm > 0;
        const uint32_t add = x_moved_in ? xw : ow;
        if (add) t_add(&T.wins[path[i]], add);
    }
    return root_pjm > 0 ? xw : ow;
}

}  // namespace sqt
