// Host-side pieces of the NegaScout split driver (squava_b200/csrc/negascout.cuh: ns_reorder_host, ns_move_ends_game),
// compiled for the host and printed for tests/test_negascout_host_pieces.py to compare with the Python oracle.
// stdin: lines "x o" (25-bit masks).  stdout per line: n, the MAXIMIZER's order, the MINIMIZER's order, then for every
// empty cell c (ascending) "c:ab" with a = 1 if X's mark on c ends the game, b = 1 if O's does.
#include <cstdio>
#include "../../squava_b200/csrc/negascout.cuh"

int main() {
    sq::NsTables t;
    sq::build_ns_tables(t);
    unsigned x, o;
    while (std::scanf("%u %u", &x, &o) == 2) {
        uint8_t mx[25], mn[25];
        const int n = sq::ns_reorder_host(t, x, o, mx, mn);
        std::printf("%d", n);
        for (int i = 0; i < n; i++) std::printf(" %d", mx[i]);
        for (int i = 0; i < n; i++) std::printf(" %d", mn[i]);
        for (int c = 0; c < 25; c++)
            if (!(((x | o) >> c) & 1u)) std::printf(" %d:%d%d", c, (int)sq::ns_move_ends_game(t, x, c), (int)sq::ns_move_ends_game(t, o, c));
        std::printf("\n");
    }
    return 0;
}
