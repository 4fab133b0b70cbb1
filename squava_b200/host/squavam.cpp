// squavam -- drop-in for the reference's stand-alone MCTS program (squavam.go): same
// flags (-i, -u, -C), same stdin protocol ("x y" per move, EOF exits 0), same output lines.
// The UCT loop's rollout block (squavam.go:128-147) runs on the GPU in leaf-parallel
// batches; extra flags: -batch, -ppl, -gpus, -seed.
#include <chrono>
#include <cmath>
#include <cstdio>

#include "squava_host.hpp"

using namespace squava;
using namespace squava::mcts;

static bool end_of_game(const GameState& s) {      // GetMoves' second result, squavam.go:246-284
    sq_board b = s.board.raw(); uint8_t f = 0;
    must(sq_classify(&b, 1, &f, nullptr, nullptr), "sq_classify");
    return f != 0;
}

static int readMove(const Board& bd) {             // squavam.go:345-372
    for (;;) {
        std::printf("Enter move (N M): ");
        std::fflush(stdout);
        int x, y;
        int n = std::scanf("%d %d", &x, &y);
        if (n == EOF) std::exit(0);
        if (n != 2) { std::printf("Failed to read: unexpected input\n"); std::exit(1); }
        if (x < 0 || x > 4 || y < 0 || y > 4) { std::printf("Choose two numbers between 0 and 4, try again\n"); continue; }
        if (bd.at(5 * x + y) == UNSET) return 5 * x + y;
        std::printf("Cell (%d, %d) already occupied, try again\n", x, y);
    }
}

int main(int argc, char** argv) {
    int iterMax = 10000, batch = 256, ppl = 1, gpus = 0, seed = -1, threads = 0;
    bool fast = false, stats = false, rootParallel = false, gpuTree = false;
    int walkers = 0, syncLevels = -1;
    double uctk = 1.00;
    bool computerFirst = false;
    Flags fl;
    fl.Int("i", &iterMax, "maximum iterations");
    fl.Float("u", &uctk, "UCTK explore/exploit coefficient");
    fl.Bool("C", &computerFirst, "Computer takes first move");
    fl.Int("batch", &batch, "leaves per GPU batch (1 = the reference's strictly sequential UCT)");
    fl.Int("ppl", &ppl, "playouts per leaf");
    fl.Bool("fast", &fast, "arena tree with multi-threaded select/update (for 10^8..10^9 iterations); no tree reuse");
    fl.Int("threads", &threads, "with -fast: host threads (default: all)");
    fl.Bool("stats", &stats, "with -fast: print iterations, nodes and the time spent selecting / on the GPU / updating");
    fl.Bool("root-parallel", &rootParallel, "with -fast -gpus N: N independent trees, one per GPU, root statistics merged once per decision (squavam2.go's N searchers without the shared tree)");
    fl.Bool("gpu-tree", &gpuTree, "the whole tree in GPU memory (sq_mcts_decide): select, playouts and update are device kernels, nothing but the decision comes back; for 10^8..10^9 iterations; no tree reuse");
    fl.Int("walkers", &walkers, "with -gpu-tree: walks of a batch in flight at once below the synchronised levels (default 8192)");
    fl.Int("sync-levels", &syncLevels, "with -gpu-tree: levels below the root on which the walks of a batch share out the children by counts (default 4; 0: none)");
    fl.Int("gpus", &gpus, "number of GPUs (default 1)");
    fl.Int("seed", &seed, "seed for reproducible play (default: clock, as the reference)");
    fl.Parse(argc, argv);
    if (seed >= 0) seed_host_rng((uint64_t)seed);
    init_devices(gpus, (uint64_t)host_rng()());

    GameState state;
    state.playerJustMoved = MINIMIZER;
    if (computerFirst) state.playerJustMoved = MAXIMIZER;
    Options opt; opt.batch = batch; opt.playouts_per_leaf = ppl; opt.ucb_factor2 = false;   // squavam.go:197-199
    std::unique_ptr<Node> movesNode;
    while (!end_of_game(state)) {
        int m;
        std::printf("%s\n", state.board.str().c_str());
        if (state.playerJustMoved == MAXIMIZER) {
            auto start = std::chrono::steady_clock::now();
            if (gpuTree) {
                sq_board rb{state.board.x, state.board.o};
                sq_mcts_result r;
                const int rc = sq_mcts_decide(0, &rb, state.playerJustMoved, iterMax, uctk, 0, batch < 1024 ? 65536 : batch, ppl, walkers, syncLevels, 0,
                                              (uint64_t)host_rng()() << 20, (uint64_t)host_rng()(), &r);
                if (rc != SQ_OK) { std::fprintf(stderr, "sq_mcts_decide: %s\n", sq_mcts_last_error()); std::exit(2); }
                // the move returned is argmax UCB1 over the root's children (mcts.go:196-198)
                const double lg = std::log(r.root_visits > 1. ? r.root_visits : 1.);
                double bestscore = 1e-6;
                m = -1;
                for (int i = 0; i < r.n_children; i++) {
                    const double sc = r.child_wins[i] / (r.child_visits[i] + 1e-6) + uctk * std::sqrt(lg / (r.child_visits[i] + 1e-6));
                    if (sc > bestscore) { bestscore = sc; m = r.child_move[i]; }
                }
                if (m < 0) { std::fprintf(stderr, "Node has %d children\n", r.n_children); std::exit(2); }
                double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
                std::printf("My move: %d %d %s\n", m / 5, m % 5, go_duration(el).c_str());
                if (stats)
                    std::printf("gpu tree: %lld iterations, %lld batches, %llu node slots of %llu%s; device %.3fs (%.3g iterations/s)\n",
                                (long long)r.iterations, (long long)r.batches, (unsigned long long)r.nodes, (unsigned long long)r.arena_slots,
                                r.arena_full ? " (FULL)" : "", r.device_ms * 1e-3, (double)r.iterations / (r.device_ms * 1e-3 + 1e-12));
                state.DoMove(m);
                movesNode.reset();
                continue;
            }
            if (fast) {
                FastOptions fo; fo.batch = batch; fo.playouts_per_leaf = ppl; fo.ucb_factor2 = false; fo.threads = threads;
                FastResult r;
                if (rootParallel) {
                    RootParallelResult rp = root_parallel_uct(state, iterMax, uctk, fo, sq_device_count());
                    r = rp.merged;
                    if (stats)
                        for (size_t i = 0; i < rp.searcher.size(); i++)
                            std::printf("searcher %zu: %lld iterations, best %d %d, select %.3fs gpu %.3fs update %.3fs\n", i, rp.searcher[i].iterations,
                                        rp.searcher[i].best_move / 5, rp.searcher[i].best_move % 5, rp.searcher[i].select_seconds,
                                        rp.searcher[i].gpu_seconds, rp.searcher[i].update_seconds);
                } else r = fast_uct(state, iterMax, uctk, fo);
                m = r.best_move;
                double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
                std::printf("My move: %d %d %s\n", m / 5, m % 5, go_duration(el).c_str());
                if (stats)
                    std::printf("fast tree: %lld iterations, %lld batches, %llu node slots, %d threads; select %.3fs (%.3g/s) gpu %.3fs update %.3fs; %lld MB in 2 MB pages\n",
                                r.iterations, r.batches, (unsigned long long)r.nodes, r.threads, r.select_seconds,
                                (double)r.iterations / ppl / (r.select_seconds > 0 ? r.select_seconds : 1e-9), r.gpu_seconds, r.update_seconds,
                                r.huge_page_kb / 1024);
                state.DoMove(m);
                movesNode.reset();
                continue;
            }
            Node* best = UCT(state, iterMax, uctk, movesNode, opt, nullptr, nullptr);
            m = best->move;
            std::unique_ptr<Node> keep;
            for (auto& c : movesNode->childNodes) if (c.get() == best) keep = std::move(c);
            keep->parentNode = nullptr;
            movesNode = std::move(keep);
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
            std::printf("My move: %d %d %s\n", m / 5, m % 5, go_duration(el).c_str());
        } else {
            m = readMove(state.board);
            if (movesNode) {     // pick out the child of movesNode corresponding to m (squavam.go:70-78)
                std::unique_ptr<Node> keep;
                for (auto& c : movesNode->childNodes) if (c->move == m) { keep = std::move(c); break; }
                if (keep) { keep->parentNode = nullptr; movesNode = std::move(keep); }
                else movesNode.reset();
            }
        }
        state.DoMove(m);
    }
    std::printf("%s\n", state.board.str().c_str());
    sq_board b = state.board.raw(); int8_t w = 0;
    must(sq_classify(&b, 1, nullptr, &w, nullptr), "sq_classify");
    char last = "O_X"[state.playerJustMoved + 1];
    if (w == state.playerJustMoved) std::printf("Player %c wins!\n", last);       // GetResult == 1.0
    else std::printf("Player %c loses!\n", last);                                // 0.0, draws included (squavam.go:88-93)
    sq_shutdown();
    return 0;
}
