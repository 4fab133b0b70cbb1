// host_capi.cpp -- a C entry point onto the host-side UCT driver (squava_host.cpp) with
// injectable callbacks, so the TREE logic (select / expand / backprop / UCB1, mcts.go:123-271)
// can be compared with the reference's on a CPU box: the rollout, the terminal test and
// rand.Intn come from the caller.  With null callbacks it runs on the GPU like squavam does.
#include <chrono>

#include "squava_host.hpp"

using namespace squava;
using namespace squava::mcts;

extern "C" {

typedef int (*sqh_intn_fn)(int n, void* ud);
typedef int (*sqh_terminal_fn)(uint32_t x, uint32_t o, void* ud);
typedef void (*sqh_playouts_fn)(const sq_board* leaves, const int8_t* to_move, int64_t n, uint64_t per_leaf,
                                uint64_t stream, uint64_t (*out)[4], void* ud);

// Runs UCT(rootstate, itermax, UCTK, nil).  Root children are reported in creation order.
int sqh_uct(uint32_t x, uint32_t o, int player_just_moved, int itermax, double uctk, int factor2, int batch,
            int playouts_per_leaf, sqh_intn_fn intn, sqh_terminal_fn terminal, sqh_playouts_fn playouts, void* ud,
            int* n_children, int moves[25], double visits[25], double wins[25], int* best_move, int* value,
            int* leaves) {
    GameState st;
    st.board.x = x; st.board.o = o; st.playerJustMoved = player_just_moved;
    Hooks hk;
    if (intn) hk.intn = [=](int n) { return intn(n, ud); };
    if (terminal) hk.terminal = [=](const Board& b) { return terminal(b.x, b.o, ud) != 0; };
    if (playouts) hk.playouts = [=](const sq_board* l, const int8_t* t, int64_t n, uint64_t p, uint64_t s, uint64_t (*out)[4]) {
        playouts(l, t, n, p, s, out, ud);
    };
    Options opt;
    opt.batch = batch; opt.playouts_per_leaf = playouts_per_leaf; opt.ucb_factor2 = factor2 != 0;
    opt.hooks = &hk;
    std::unique_ptr<Node> root;
    int lv = 0, val = 0;
    Node* best = UCT(st, itermax, uctk, root, opt, &lv, &val);
    int k = 0;
    for (auto& c : root->childNodes) {
        if (k < 25) { moves[k] = c->move; visits[k] = c->visits; wins[k] = c->wins; }
        k++;
    }
    *n_children = k; *best_move = best->move; *value = val; *leaves = lv;
    return 0;
}

// The same through the arena tree of fast_tree.cpp (one thread when callbacks are given).
int sqh_fast_uct(uint32_t x, uint32_t o, int player_just_moved, long long itermax, double uctk, int factor2, int batch,
                 int playouts_per_leaf, int threads, sqh_intn_fn intn, sqh_terminal_fn terminal, sqh_playouts_fn playouts,
                 void* ud, int* n_children, int moves[25], double visits[25], double wins[25], int* best_move, int* value,
                 double* root_visits, double seconds[3], unsigned long long* nodes) {
    GameState st;
    st.board.x = x; st.board.o = o; st.playerJustMoved = player_just_moved;
    Hooks hk;
    if (intn) hk.intn = [=](int n) { return intn(n, ud); };
    if (terminal) hk.terminal = [=](const Board& b) { return terminal(b.x, b.o, ud) != 0; };
    if (playouts) hk.playouts = [=](const sq_board* l, const int8_t* t, int64_t n, uint64_t p, uint64_t s, uint64_t (*out)[4]) {
        playouts(l, t, n, p, s, out, ud);
    };
    FastOptions opt;
    opt.batch = batch; opt.playouts_per_leaf = playouts_per_leaf; opt.ucb_factor2 = factor2 != 0; opt.threads = threads;
    if (intn || terminal || playouts) opt.hooks = &hk;
    FastResult r = fast_uct(st, itermax, uctk, opt);
    for (int k = 0; k < r.n_children; k++) { moves[k] = r.child_move[k]; visits[k] = r.child_visits[k]; wins[k] = r.child_wins[k]; }
    *n_children = r.n_children; *best_move = r.best_move; *value = r.value;
    if (root_visits) *root_visits = r.root_visits;
    if (seconds) { seconds[0] = r.select_seconds; seconds[1] = r.gpu_seconds; seconds[2] = r.update_seconds; }
    if (nodes) *nodes = r.nodes;
    return 0;
}

// Root-parallel: n_searchers independent arena trees (searcher i on device slot i % devices; with callbacks all on
// the callbacks), statistics merged by move.  device_slot >= 0 pins a ONE-searcher call to that device and
// stream_base separates the random streams of searchers living in different processes (torchrun ranks).
int sqh_root_parallel_uct(uint32_t x, uint32_t o, int player_just_moved, long long itermax, double uctk, int factor2, int batch,
                          int playouts_per_leaf, int threads, int n_searchers, int device_slot, unsigned long long stream_base,
                          unsigned long long rng_seed, sqh_terminal_fn terminal, sqh_playouts_fn playouts, void* ud,
                          double visits[25], double wins[25], int* best_move, int* value, double* root_visits, double* seconds) {
    GameState st;
    st.board.x = x; st.board.o = o; st.playerJustMoved = player_just_moved;
    Hooks hk;
    if (terminal) hk.terminal = [=](const Board& b) { return terminal(b.x, b.o, ud) != 0; };
    if (playouts) hk.playouts = [=](const sq_board* l, const int8_t* t, int64_t n, uint64_t p, uint64_t s, uint64_t (*out)[4]) {
        playouts(l, t, n, p, s, out, ud);
    };
    FastOptions opt;
    opt.batch = batch; opt.playouts_per_leaf = playouts_per_leaf; opt.ucb_factor2 = factor2 != 0; opt.threads = threads;
    opt.stream_base = stream_base; opt.rng_seed = rng_seed;
    if (terminal || playouts) opt.hooks = &hk;
    FastResult m;
    double secs;
    if (n_searchers <= 1 && device_slot >= 0) {
        opt.device_slot = device_slot;
        auto t0 = std::chrono::steady_clock::now();
        m = fast_uct(st, itermax, uctk, opt);
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } else {
        RootParallelResult r = root_parallel_uct(st, itermax, uctk, opt, n_searchers);
        m = r.merged; secs = r.seconds;
    }
    for (int c = 0; c < 25; c++) { visits[c] = 0; wins[c] = 0; }
    for (int k = 0; k < m.n_children; k++) { visits[m.child_move[k]] = m.child_visits[k]; wins[m.child_move[k]] = m.child_wins[k]; }
    *best_move = m.best_move; *value = m.value;
    if (root_visits) *root_visits = m.root_visits;
    if (seconds) *seconds = secs;
    return 0;
}

// Drives src/abbook's player (abbook.go:93-131) against a scripted opponent, the way the golden
// vectors were made.  role 0: the engine moves first; role 1: the opponent does.  With
// book_only != 0 the run stops before the first move that needs the search (no GPU needed).
// Returns the number of engine moves written; -99 if the book hit its os.Exit(99) path.
int sqh_abbook_script(int role, int pick, int max_depth, const int* script, int n_script, int book_only,
                      int moves[], int values[], int from_book[], int max_out, int board[25], int* still_in_book) {
    abbook::AlphaBetaBook p(true, max_depth);
    p.SetScores(false);
    p.SetBookPick(pick);
    int k = 0;
    auto engine = [&]() -> bool {
        const bool in_book = p.bookInProgress();
        if (book_only && !in_book) return false;
        Move mv = p.ChooseMove();
        if (book_only && mv.leafCount != 0) return false;
        if (k < max_out) { moves[k] = 5 * mv.x + mv.y; values[k] = mv.value; from_book[k] = in_book; }
        k++;
        return true;
    };
    bool go = true;
    if (role == 0) {
        for (int i = 0; i <= n_script && go; i++) {
            go = engine();
            if (!go || i == n_script || p.board().at(script[i]) != UNSET) break;
            p.MakeMove(script[i] / 5, script[i] % 5, MINIMIZER);
        }
    } else {
        for (int i = 0; i < n_script && go; i++) {
            if (p.board().at(script[i]) != UNSET) break;
            p.MakeMove(script[i] / 5, script[i] % 5, MINIMIZER);
            go = engine();
        }
    }
    for (int c = 0; c < 25; c++) board[c] = p.board().at(c);
    *still_in_book = p.bookInProgress();
    return k;
}

}  // extern "C"
