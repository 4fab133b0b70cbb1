// sq_tables.h -- host-side construction of the scan-order line tables.
//
// The fast paths never read a table (lines are found with shifts); the tables
// only decide WHICH owned line a reference scan meets first on boards where
// both colours own one.  Order sources:
//   mcts order: src/mcts/mcts.go:404-446 (important cells 2,7,10,11,12,13,14,17,22,
//               each with its own list of quads / triplets)
//   ab order:   src/alphabeta/alphabeta.go:249-328 (flat lists)
#pragma once
#include <cstdint>
#include "sq_device.cuh"

namespace sq {

// A line = cells start, start+step, ... ; encoded start*16 + (step+8).
constexpr uint16_t L(int start, int step) { return (uint16_t)(start * 16 + step + 8); }

// mcts.go scan order, flattened over the nine important cells.
constexpr uint16_t kMctsQuadOrder[28] = {
    L(0, 1),  L(1, 1),  L(2, 5),                                  // cell 2
    L(5, 1),  L(6, 1),  L(7, 5),                                  // cell 7
    L(0, 5),  L(5, 5),                                            // cell 10
    L(1, 5),  L(6, 5),  L(3, 4),  L(5, 6),                        // cell 11
    L(10, 1), L(11, 1), L(4, 4),  L(8, 4),  L(0, 6),  L(6, 6),    // cell 12
    L(3, 5),  L(8, 5),  L(1, 6),  L(9, 4),                        // cell 13
    L(4, 5),  L(9, 5),                                            // cell 14
    L(15, 1), L(16, 1),                                           // cell 17
    L(20, 1), L(21, 1),                                           // cell 22
};
constexpr uint16_t kMctsTripletOrder[72] = {
    L(0, 1),  L(1, 1),  L(2, 1),  L(2, 5),  L(2, 4),  L(14, -6),                                  // 2
    L(5, 1),  L(6, 1),  L(7, 1),  L(2, 5),  L(7, 5),  L(3, 4),  L(7, 4),  L(1, 6),  L(7, 6),      // 7
    L(10, 1), L(0, 5),  L(5, 5),  L(10, 5), L(2, 4),  L(10, 6),                                   // 10
    L(10, 1), L(11, 1), L(1, 5),  L(6, 5),  L(11, 5), L(3, 4),  L(7, 4),  L(5, 6),  L(11, 6),     // 11
    L(10, 1), L(11, 1), L(12, 1), L(2, 5),  L(7, 5),  L(12, 5), L(0, 6),  L(6, 6),  L(12, 6),
    L(4, 4),  L(8, 4),  L(12, 4),                                                                 // 12
    L(11, 1), L(12, 1), L(3, 5),  L(8, 5),  L(13, 5), L(1, 6),  L(7, 6),  L(21, -4), L(17, -4),   // 13
    L(12, 1), L(4, 5),  L(9, 5),  L(14, 5), L(22, -4), L(14, -6),                                 // 14
    L(15, 1), L(16, 1), L(17, 1), L(7, 5),  L(12, 5), L(5, 6),  L(11, 6), L(21, -4), L(17, -4),   // 17
    L(20, 1), L(21, 1), L(22, 1), L(12, 5), L(10, 6), L(22, -4),                                  // 22
};

inline uint32_t line_mask(uint16_t code, int len) {
    int start = code / 16, step = code % 16 - 8;
    uint32_t m = 0;
    for (int i = 0; i < len; i++) m |= 1u << (start + i * step);
    return m;
}

inline void build_scan_tables(ScanTables& t) {
    for (int i = 0; i < 28; i++) t.mcts_quads[i] = line_mask(kMctsQuadOrder[i], 4);
    int n = 0;
    for (int i = 0; i < 72; i++) {
        uint32_t m = line_mask(kMctsTripletOrder[i], 3);
        bool seen = false;
        for (int j = 0; j < n; j++) seen |= (t.mcts_triplets[j] == m);
        if (!seen && n < 48) t.mcts_triplets[n++] = m;
    }
    // alphabeta.go lists quads by first cell with x outermost, triplets with y
    // outermost; directions from a first cell: (x+1), (y+1), (x-1,y+1), (x+1,y+1).
    const int dx[4] = {1, 0, -1, 1}, dy[4] = {0, 1, 1, 1};
    auto ok = [](int x, int y) { return x >= 0 && x < 5 && y >= 0 && y < 5; };
    int nq = 0, nt = 0;
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++)
            for (int d = 0; d < 4; d++)
                if (ok(x + 3 * dx[d], y + 3 * dy[d])) {
                    uint32_t m = 0;
                    for (int i = 0; i < 4; i++) m |= 1u << (5 * (x + i * dx[d]) + y + i * dy[d]);
                    t.ab_quads[nq++] = m;
                }
    for (int y = 0; y < 5; y++)
        for (int x = 0; x < 5; x++)
            for (int d = 0; d < 4; d++)
                if (ok(x + 2 * dx[d], y + 2 * dy[d])) {
                    uint32_t m = 0;
                    for (int i = 0; i < 3; i++) m |= 1u << (5 * (x + i * dx[d]) + y + i * dy[d]);
                    t.ab_triplets[nt++] = m;
                }
}

}  // namespace sq
