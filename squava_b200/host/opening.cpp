// opening2 / opening3 -- drop-ins for the reference's deep opening evaluations
// (opening2.go:62-71, opening3.go:66-128).  The independent subtrees (6, or the 85
// symmetry-unique first-move/reply pairs) go to the GPU as ONE sq_alphabeta_subtrees batch;
// output lines keep the reference's format (per-subtree elapsed is the batch time, since
// the subtrees run concurrently).  Built twice: -DOPENING=2 and -DOPENING=3.
#include <chrono>
#include <cstdio>
#include <cstring>

#include "squava_host.hpp"

using namespace squava;

#ifndef OPENING
#define OPENING 3
#endif

// opening3.go:366-372 / opening2.go:307-313: the default bias table with [2][4] = 2
static const int8_t scores[25] = {3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 2, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3};
// opening3.go:456-463 / opening2.go:262-269
static const int firstMoves[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};

// opening3.go:379-451: image of (x,y) under the k-th symmetry
[[maybe_unused]] static void image(int k, int x, int y, int& a, int& b) {
    switch (k) {
    case 0: a = x; b = y; break;              // identity
    case 1: a = y; b = 4 - x; break;          // rotate 90 CCW
    case 2: a = 4 - x; b = 4 - y; break;      // rotate 180
    case 3: a = 4 - y; b = x; break;          // rotate 90 CW
    case 4: a = 4 - x; b = y; break;          // reflect across horizontal axis
    case 5: a = x; b = 4 - y; break;          // reflect across vertical axis
    case 6: a = 4 - y; b = 4 - x; break;      // reflect across anti-diagonal
    default: a = y; b = x; break;             // reflect across main diagonal
    }
}

int main(int argc, char** argv) {
    int maxDepth = OPENING == 2 ? 9 : 10, gpus = 0;
    std::string bookFile;
    Flags fl;
    fl.Int("d", &maxDepth, "maximum lookahead depth");
#if OPENING == 3
    fl.String("book", &bookFile, "also write an opening-book table (all 600 first-move/reply pairs, by symmetry) to this file");
#endif
    fl.Int("gpus", &gpus, "number of GPUs (default 1)");
    fl.Parse(argc, argv);
    init_devices(gpus, 0);
    std::vector<sq_subtree> st;
#if OPENING == 2
    for (auto& cell : firstMoves) {             // opening2.go:62-71
        sq_subtree t; std::memset(&t, 0, sizeof t);
        t.b.x = 1u << (5 * cell[0] + cell[1]);
        t.last_cell = (int8_t)(5 * cell[0] + cell[1]); t.ply = 1; t.side_to_move = MINIMIZER; t.carried = 0;
        st.push_back(t);
    }
    std::vector<int32_t> val(st.size()); std::vector<uint64_t> nodes(st.size());
    auto t0 = std::chrono::steady_clock::now();
    must(sq_alphabeta_subtrees(st.data(), (int64_t)st.size(), SQ_AB_OPENING2, maxDepth, scores, val.data(), nodes.data()), "sq_alphabeta_subtrees");
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t k = 0; k < st.size(); k++) {
        int i = firstMoves[k][0], j = firstMoves[k][1];
        std::printf("<%d,%d>\t%d (%d) [%llu]\t%s\n", i, j, val[k], scores[5 * i + j], (unsigned long long)nodes[k], go_duration(el).c_str());
    }
#else
    int computedMoves[6][5][5]; std::memset(computedMoves, 0, sizeof computedMoves);
    for (int k = 0; k < 6; k++) computedMoves[k][firstMoves[k][0]][firstMoves[k][1]] = MAXIMIZER;
    struct Row { int first, x, y, p, q; };
    std::vector<Row> rows;
    int dup[6] = {0, 0, 0, 0, 0, 0};
    for (int fi = 0; fi < 6; fi++) {            // opening3.go:66-128
        int x = firstMoves[fi][0], y = firstMoves[fi][1];
        for (int p = 0; p < 5; p++)
            for (int q = 0; q < 5; q++) {
                if (p == x && q == y) continue;
                bool matched = false;
                for (int m = 0; m < 8 && !matched; m++) {
                    int a, b, c, d;
                    image(m, x, y, a, b); image(m, p, q, c, d);
                    for (int bi = 0; bi < 6; bi++)
                        if (computedMoves[bi][a][b] == MAXIMIZER && computedMoves[bi][c][d] == MINIMIZER) { matched = true; break; }
                }
                if (matched) { dup[fi]++; continue; }
                computedMoves[fi][p][q] = MINIMIZER;
                rows.push_back({fi, x, y, p, q});
                sq_subtree t; std::memset(&t, 0, sizeof t);
                t.b.x = 1u << (5 * x + y); t.b.o = 1u << (5 * p + q);
                t.last_cell = (int8_t)(5 * p + q); t.ply = 2; t.side_to_move = MAXIMIZER; t.carried = scores[5 * x + y];
                st.push_back(t);
            }
    }
    std::vector<int32_t> val(st.size()); std::vector<uint64_t> nodes(st.size());
    auto t0 = std::chrono::steady_clock::now();
    must(sq_alphabeta_subtrees(st.data(), (int64_t)st.size(), SQ_AB_OPENING3, maxDepth, scores, val.data(), nodes.data()), "sq_alphabeta_subtrees");
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    size_t k = 0;
    for (int fi = 0; fi < 6; fi++) {
        unsigned long long sum = 0; int uniq = 0;
        for (; k < rows.size() && rows[k].first == fi; k++) {
            const Row& r = rows[k];
            std::printf("<%d,%d>:<%d,%d>\t%d (%d) [%llu]\t%s\n", r.x, r.y, r.p, r.q, val[k], scores[5 * r.x + r.y],
                        (unsigned long long)nodes[k], go_duration(el).c_str());
            sum += nodes[k]; uniq++;
        }
        std::printf("Leaf nodes visited: %llu\nUnique moves computed: %d\nDuplicated moves: %d\n\n", sum, uniq, dup[fi]);
    }
    // the loop the reference leaves open.  The 85 computed pairs, carried through the 8
    // symmetries the program itself uses to skip duplicates (opening3.go:379-451), give a value for
    // every (first move, reply) pair: a table a -book reader (squavathr -book) can play from.
    // Line format: "x1 y1 x2 y2 value depth", first mover's view (MAXIMIZER), 600 lines.
    if (!bookFile.empty()) {
        int table[25][25]; bool have[25][25];
        std::memset(have, 0, sizeof have);
        for (size_t r = 0; r < rows.size(); r++)
            for (int m = 0; m < 8; m++) {
                int a, b, c, d;
                image(m, rows[r].x, rows[r].y, a, b); image(m, rows[r].p, rows[r].q, c, d);
                if (!have[5 * a + b][5 * c + d]) { have[5 * a + b][5 * c + d] = true; table[5 * a + b][5 * c + d] = val[r]; }
            }
        FILE* fp = std::fopen(bookFile.c_str(), "w");
        if (!fp) { std::fprintf(stderr, "open %s: cannot write\n", bookFile.c_str()); return 1; }
        int missing = 0;
        for (int f = 0; f < 25; f++)
            for (int r = 0; r < 25; r++) {
                if (f == r) continue;
                if (!have[f][r]) { missing++; continue; }
                std::fprintf(fp, "%d %d %d %d %d %d\n", f / 5, f % 5, r / 5, r % 5, table[f][r], maxDepth);
            }
        std::fclose(fp);
        if (missing) std::fprintf(stderr, "book: %d pairs not covered by the symmetries\n", missing);
    }
#endif
    sq_shutdown();
    return 0;
}
