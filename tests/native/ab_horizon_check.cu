// Host-side check of squava_b200/csrc/ab_horizon.cuh (TEST INFRASTRUCTURE): the bit-parallel value of a horizon-1
// node against a per-child walk over the line tables, on random boards.  Compiled with nvcc, run on the CPU.
//   usage: ab_horizon_check <cases> <seed>     prints "ok <cases> <nonzero counts> <terminal nodes>" or the first mismatch
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../squava_b200/csrc/alphabeta.cuh"

using namespace sq;

static AbTables T;

static int count_three_of_four(uint32_t own, uint32_t opp, int c, uint32_t not_hole) {
    int n = 0;
    for (int k = 0; k < T.n_quads[c]; k++) {
        uint32_t q = T.quads[c][k];
        if (__builtin_popcount(own & q) == 3 && !(opp & q)) { uint32_t hole = q & ~own; if (!(hole & not_hole)) n++; }
    }
    return n;
}
static bool has_four(uint32_t own, int c) { for (int k = 0; k < T.n_quads[c]; k++) if ((own & T.quads[c][k]) == T.quads[c][k]) return true; return false; }
static bool has_three(uint32_t own, int c) { for (int k = 0; k < T.n_triplets[c]; k++) if ((own & T.triplets[c][k]) == T.triplets[c][k]) return true; return false; }

int main(int argc, char** argv) {
    long cases = argc > 1 ? atol(argv[1]) : 200000;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 1);
    build_ab_tables(T);
    long nonzero = 0, terminal = 0;
    for (long it = 0; it < cases; it++) {
        int8_t scores[25];
        const int mode = (int)(rng() % 3);          // default table, small random, wide (span <= 29) random
        static const int8_t def[25] = {3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3};
        for (int c = 0; c < 25; c++) scores[c] = mode == 0 ? def[c] : mode == 1 ? (int8_t)((int)(rng() % 11) - 5) : (int8_t)((int)(rng() % 30) - 17);
        HorizonTables H;
        build_horizon_tables(scores, H);
        if (!H.ok) { printf("tables not ok\n"); return 1; }
        uint32_t own = 0, opp = 0;
        const double fill = (rng() % 1000) / 1000.0;
        for (int c = 0; c < 25; c++) {
            double r = (rng() % 100000) / 100000.0;
            if (r < fill * 0.5) own |= 1u << c; else if (r < fill) opp |= 1u << c;
        }
        const uint32_t emp = kFull25 & ~(own | opp);
        if (__builtin_popcount(emp) < 2) { it--; continue; }
        uint32_t pe1 = 0, pe2 = 0;
        for (int c = 0; c < 25; c++) if ((emp >> c & 1) && rng() % 4 == 0) { pe1 |= 1u << c; if (rng() % 3 == 0) pe2 |= 1u << c; }
        const int term = 10000 - (int)(rng() % 20), konst_d = -30 * (__builtin_popcount(pe1) + __builtin_popcount(pe2)) - (int)(rng() % 9) + (int)(rng() % 200) - 100;
        const int konst_n = (int)(rng() % 600) - 300;
        const uint32_t f0 = emp & (0u - emp), f1 = (emp ^ f0) & (0u - (emp ^ f0));
        // reference: per child
        int ref_d = -0x40000000, ref_n = -0x40000000; bool any_win = false;
        for (int c = 0; c < 25; c++) {
            if (!(emp >> c & 1)) continue;
            const uint32_t bit = 1u << c, o2 = own | bit;
            int vd, vn;
            if (has_four(o2, c)) { vd = vn = term; any_win = true; }
            else if (has_three(o2, c)) vd = vn = -term;
            else {
                const uint32_t j0 = bit == f0 ? f1 : f0;
                const int cnt_d = count_three_of_four(o2, opp, c, j0) + ((pe1 & bit) ? 1 : 0) + ((pe2 & bit) ? 1 : 0);
                const int cnt_n = count_three_of_four(o2, opp, c, 0);
                vd = 30 * cnt_d + scores[c] + konst_d;
                vn = 30 * cnt_n + scores[c] + konst_n;
                nonzero += cnt_n > 0;
            }
            ref_d = vd > ref_d ? vd : ref_d; ref_n = vn > ref_n ? vn : ref_n;
        }
        terminal += any_win;
        const int cf0 = __builtin_ctz(f0);
        const int got_d = horizon_node_value<true>(own, opp, emp, term, konst_d, pe1, pe2, scores[cf0], H);
        const int got_n = horizon_node_value<false>(own, opp, emp, term, konst_n, 0, 0, 0, H);
        if (got_d != ref_d || got_n != ref_n) {
            printf("MISMATCH own %07x opp %07x pe1 %07x pe2 %07x term %d: delayed %d vs %d, entry %d vs %d (mode %d)\n", own, opp, pe1, pe2, term, got_d, ref_d, got_n, ref_n, mode);
            return 1;
        }
    }
    printf("ok %ld %ld %ld\n", cases, nonzero, terminal);
    return 0;
}
