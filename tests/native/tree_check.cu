// The HBM tree's walk and update (squava_b200/csrc/tree.cuh) run sequentially on the host against a CPU rollout
// (uniform random legal moves; four in a row wins, three loses, four first): the books must balance at every node,
// and the search must find a move that wins at once.
// usage: tree_check ITERATIONS BATCH PLAYOUTS_PER_LEAF SEED X O PJM [SYNC_LEVELS]   -> prints what tests/test_tree_native.py reads
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../squava_b200/csrc/tree.cuh"
#include "../../squava_b200/csrc/sq_tables.h"

static sq::ScanTables S;
static uint64_t rng_state;
static uint32_t rnd() { rng_state = sqt::t_mix(rng_state); return (uint32_t)(rng_state >> 32); }

static bool has(uint32_t own, const uint32_t* masks, int n) { for (int i = 0; i < n; i++) if ((own & masks[i]) == masks[i]) return true; return false; }

// one random playout from (x, o), `to_move` (+1 X) to move: -> +1 X wins, -1 O wins, 0 draw; *plies = plies played
static int rollout(uint32_t x, uint32_t o, int to_move, int* plies) {
    *plies = 0;
    // a position on which the game is over plays no ply: four wins, three loses, four first (what the playout kernel reports)
    if (has(x, S.ab_quads, 28)) return 1;
    if (has(o, S.ab_quads, 28)) return -1;
    if (has(x, S.ab_triplets, 48)) return -1;
    if (has(o, S.ab_triplets, 48)) return 1;
    for (;;) {
        const uint32_t emp = sqt::kFull & ~(x | o);
        if (!emp) return 0;
        const int m = sqt::t_nth_set_bit(emp, (int)(rnd() % (uint32_t)__builtin_popcount(emp)));
        uint32_t& own = to_move > 0 ? x : o;
        own |= 1u << m;
        (*plies)++;
        if (has(own, S.ab_quads, 28)) return to_move;
        if (has(own, S.ab_triplets, 48)) return -to_move;
        to_move = -to_move;
    }
}

int main(int argc, char** argv) {
    if (argc < 8) return 2;
    const long long iterations = atoll(argv[1]), batch = atoll(argv[2]);
    const uint32_t P = (uint32_t)atoi(argv[3]);
    rng_state = strtoull(argv[4], nullptr, 10);
    const uint32_t rx = (uint32_t)strtoul(argv[5], nullptr, 0), ro = (uint32_t)strtoul(argv[6], nullptr, 0);
    const int pjm = atoi(argv[7]);
    const int sync_levels = argc > 8 ? atoi(argv[8]) : 0;
    sq::build_scan_tables(S);
    const unsigned long long cap = (unsigned long long)(26.0 * (double)(iterations / P + 1)) + 1024;
    std::vector<uint32_t> visits(cap, 0), wins(cap, 0), first_child(cap, 0), untried(cap, 0), arrivals(cap, 0);
    unsigned long long next_free = 2;
    uint32_t overflow = 0;
    sqt::Tree T{visits.data(), wins.data(), first_child.data(), untried.data(), &next_free, cap, &overflow, arrivals.data()};
    const uint32_t root = 1;
    // a root on which the game is already over has no moves
    const bool over = has(rx, S.ab_quads, 28) || has(ro, S.ab_quads, 28) || has(rx, S.ab_triplets, 48) || has(ro, S.ab_triplets, 48);
    untried[root] = over ? 0u : (sqt::kFull & ~(rx | ro));
    std::vector<sq_board> boards((size_t)batch);
    std::vector<int8_t> to_move((size_t)batch);
    std::vector<uint32_t> leaf((size_t)batch), paths((size_t)batch * sqt::kPathLen);
    std::vector<uint8_t> fresh((size_t)batch), depth((size_t)batch);
    std::vector<uint32_t> rank((size_t)batch);
    std::vector<uint8_t> act((size_t)batch), arg((size_t)batch);
    std::vector<int8_t> wpjm((size_t)batch);
    sqt::Walks W{boards.data(), to_move.data(), leaf.data(), fresh.data(), depth.data(), paths.data(), rank.data(), act.data(), arg.data(), wpjm.data()};
    std::vector<unsigned long long> counts((size_t)batch * 4);
    long long done = 0, batches = 0;
    while (done < iterations) {
        const long long remaining = iterations - done;
        uint32_t p = P;
        if ((long long)p > remaining) p = (uint32_t)remaining;
        long long B = batch;
        if (batches < 16 && (256ll << batches) < B) B = 256ll << batches;       // the first batches grow with the tree: 256, 512, ...
        if (B * (long long)p > remaining) B = remaining / p;
        if (B < 1) B = 1;
        const uint64_t key = sqt::t_mix((uint64_t)batches);
        if (sync_levels == 0) for (long long k = 0; k < B; k++) sqt::tree_walk(T, root, rx, ro, pjm, p, 2.0f, key, k, W, true);
        else {
            visits[root] += (uint32_t)(B * (long long)p);
            for (int level = 0; level < sync_levels; level++) {
                for (long long k = 0; k < B; k++) sqt::tree_arrive(T, level, root, rx, ro, pjm, p, k, W, true);
                for (long long k = 0; k < B; k++) sqt::tree_decide(T, p, 2.0f, key, k, W);
            }
            for (long long k = 0; k < B; k++) sqt::tree_finish(T, sync_levels, root, rx, ro, pjm, p, 2.0f, key, k, W);
        }
        for (long long k = 0; k < B; k++) {
            unsigned long long* c = &counts[(size_t)k * 4];
            c[0] = c[1] = c[2] = c[3] = 0;
            for (uint32_t r = 0; r < p; r++) {
                int plies = 0;
                const int w = rollout(boards[(size_t)k].x, boards[(size_t)k].o, to_move[(size_t)k], &plies);
                // a leaf on which the game is over plays no ply: the playout kernel reports the winner of the position
                if (w > 0) c[0]++; else if (w < 0) c[1]++; else c[2]++;
                c[3] += (unsigned long long)plies;
            }
        }
        uint32_t root_wins = 0;
        for (long long k = 0; k < B; k++) root_wins += sqt::tree_update(T, pjm, &counts[(size_t)k * 4], k, W);
        wins[root] += root_wins;
        done += B * (long long)p;
        batches++;
    }
    // the books: every walk is counted once at every node of its path
    unsigned long long bad = 0, pending = 0, nodes = 0, root_gap = 0;
    for (unsigned long long i = 0; i < next_free; i++) if (arrivals[i]) bad++;          // every counter was cleared
    for (unsigned long long i = 1; i < next_free; i++) {
        if (!visits[i]) continue;
        nodes++;
        if (wins[i] > visits[i]) bad++;
        if (untried[i] & sqt::kPending) pending++;
        const uint32_t fc = first_child[i];
        if (!fc) continue;
        // the children's visits add up to the node's, less the walks that ended AT the node (its own first rollouts)
        unsigned long long sum = 0;
        // (the number of slots of the block is not stored: the root's is known, the others are checked through the root's total below)
        if (i == root) {
            const int slots = __builtin_popcount(sqt::kFull & ~(rx | ro));
            for (int k = 0; k < slots; k++) sum += visits[fc + (uint32_t)k];
            if (sum > visits[root]) bad++;
            root_gap = visits[root] - sum;                 // walks that rolled out from the root itself (no child was ready)
        }
    }
    const uint32_t fc = first_child[root];
    const uint32_t emp = sqt::kFull & ~(rx | ro);
    const int slots = __builtin_popcount(emp);
    std::printf("root_visits %u iterations %lld batches %lld nodes %llu used %llu bad %llu pending %llu overflow %u root_gap %llu\n", visits[root], done,
                batches, nodes, next_free, bad, pending, overflow, root_gap);
    uint32_t e = emp;
    for (int k = 0; k < slots && fc; k++, e &= e - 1u) std::printf("child %d %u %u\n", __builtin_ctz(e), visits[fc + (uint32_t)k], wins[fc + (uint32_t)k]);
    return 0;
}
