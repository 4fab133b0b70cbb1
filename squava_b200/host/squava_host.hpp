// squava_host.hpp -- host-side mirror of the reference's engine surface, in C++
// (the reference is Go; this image has no Go toolchain).
//
// Same names, argument meaning and failure behaviour as the reference:
//   Player            playoff5.go:21-29 (identical copy sqv.go:23-31)
//   mcts::MCTS        src/mcts/mcts.go:32-90  (New, SetUCTK, SetIterations, ...)
//   alphabeta::AlphaBeta  src/alphabeta/alphabeta.go:20-117
//   MoveKeeper        src/movekeeper/movekeeper.go:10-52
// Everything that is hot in the reference (rollouts + terminal test, the
// alpha-beta recursion) is a call into libsquava_b200.so (include/squava_b200.h);
// what stays here is what the north star leaves on the host: the MCTS tree
// (select / expand / backprop, UCB1, tree reuse) and move bookkeeping.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <functional>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/squava_b200.h"

namespace squava {

constexpr int MAXIMIZER = 1, MINIMIZER = -1, UNSET = 0;
constexpr int WIN = 10000, LOSS = -10000;

// The reference never returns errors: failure is panic or os.Exit.  Same here.
inline void must(int status, const char* what) {
    if (status != SQ_OK) {
        std::fprintf(stderr, "%s: %s (%s)\n", what, sq_strerror(status), sq_last_error());
        std::exit(2);
    }
}

struct Board {            // [25]int / [5][5]int of the reference as two masks
    uint32_t x = 0, o = 0;
    int at(int c) const { return (x >> c) & 1u ? MAXIMIZER : (o >> c) & 1u ? MINIMIZER : UNSET; }
    void set(int c, int player) {
        x &= ~(1u << c); o &= ~(1u << c);
        if (player == MAXIMIZER) x |= 1u << c; else if (player == MINIMIZER) o |= 1u << c;
    }
    uint32_t empty() const { return 0x1FFFFFFu & ~(x | o); }
    sq_board raw() const { return sq_board{x, o}; }
    std::string str() const;        // GameState.String(), squavam.go:327-339
    std::string str_grid() const;   // PrintBoard(), alphabeta.go:226-247
};

struct Move { int x, y, value, leafCount; };

class Player {             // playoff5.go:21-29
public:
    virtual ~Player() {}
    virtual std::string Name() = 0;
    virtual void MakeMove(int x, int y, int player) = 0;
    virtual void SetDepth(int moveCounter) = 0;
    virtual Move ChooseMove() = 0;   // x,y coords of move, value, leaf node count
    virtual void PrintBoard() = 0;
    virtual void SetScores(bool randomize) = 0;
    virtual int FindWinner() = 0;
};

// Engines whose ChooseMove is "maybe a book move, else ONE library search" expose the two halves, so
// that a harness can advance many games in lockstep with one batched library call per ply
// (playoff5 -lockstep).  ChooseMove() itself is BookMove / Request / search / Finish.
constexpr int kSearchNegaScout = 100;
struct SearchRequest { sq_board board; int kind; int depth; const int8_t* scores; };   // kind: SQ_AB_* or kSearchNegaScout
struct SearchResult {
    int32_t values[25]; int8_t visit_order[25]; int32_t best_value; uint32_t best_mask; int8_t first_best; uint64_t nodes;
};
class BatchSearchable {
public:
    virtual ~BatchSearchable() {}
    virtual bool BookMove(Move*) { return false; }     // true: a move was made without searching
    virtual SearchRequest Request() = 0;
    virtual Move Finish(const SearchResult& r) = 0;    // movekeeper + MakeMove
};
// one library call per (kind, depth, bias table) group; results in request order
void search_batch(const std::vector<SearchRequest>& req, std::vector<SearchResult>& res);

// src/movekeeper/movekeeper.go
class MoveKeeper {
public:
    MoveKeeper(int max, bool deterministic, std::mt19937_64* rng) : max_(max), det_(deterministic), rng_(rng) {}
    void SetMove(int a, int b, int value);
    void ChooseMove(int& x, int& y, int& value);
    int count() const { return next_; }
private:
    int moves_[25][2]; int next_ = 0; int max_; bool det_; std::mt19937_64* rng_;
};

std::mt19937_64& host_rng();           // rand.Seed(time.Now()...) unless SQUAVA_SEED is set
void seed_host_rng(uint64_t seed);
std::string go_duration(double seconds);  // time.Duration.String()

// ---- MCTS ------------------------------------------------------------------------
namespace mcts {

struct GameState {         // mcts.go:9-13 (cachedResults is only a memo there)
    int playerJustMoved = MINIMIZER;
    Board board;
    void DoMove(int move) { playerJustMoved = -playerJustMoved; board.set(move, playerJustMoved); }  // :292-295
};

struct Node {              // mcts.go:15-23
    int move = -1;
    Node* parentNode = nullptr;
    std::vector<std::unique_ptr<Node>> childNodes;
    double wins = 0, visits = 0;
    std::vector<int> untriedMoves;
    int playerJustMoved = 0;
};

// What UCT needs from outside the tree.  Default (null members): the GPU library and the host
// RNG.  Tests substitute deterministic callbacks to compare the tree logic with the reference's.
struct Hooks {
    std::function<int(int)> intn;                                  // rand.Intn
    std::function<bool(const Board&)> terminal;                    // GetMoves' second result
    std::function<void(const sq_board*, const int8_t*, int64_t, uint64_t, uint64_t, uint64_t (*)[4])> playouts;
};

struct Options {
    int batch = 256;                 // leaves per GPU call (leaf-parallel); 1 = the reference's sequence
    int playouts_per_leaf = 1;       // rollouts per expanded leaf
    bool ucb_factor2 = true;         // src/mcts (mcts.go:248-250) has 2*ln; squavam.go:197-199 does not
    const Hooks* hooks = nullptr;
};

// UCT(rootstate, itermax, UCTK, rootnode), squavam.go:98 / mcts.go:123.  Returns the chosen
// child (owned by *root); leaves = rollouts performed; value = int(1000*UCB1) as mcts.go:198.
Node* UCT(const GameState& rootstate, int itermax, double UCTK, std::unique_ptr<Node>& root,
          const Options& opt, int* leaves, int* value);

// The same search on a tree built for 10^9 iterations and many host threads (fast_tree.cpp): 24-byte
// nodes in an arena, contiguous child blocks, atomic counts, three phases per batch (parallel
// select / one sq_playouts call / parallel update).  One decision, no tree reuse.
struct FastOptions {
    int batch = 65536;               // leaves per GPU call
    int playouts_per_leaf = 1;
    bool ucb_factor2 = true;
    int threads = 0;                 // selecting / updating host threads; 0 = all hardware threads
    uint64_t max_nodes = 0;          // arena slots (virtual memory); 0 = sized from the iteration count
    bool exact_ucb = false;          // UCB1 in double, reference tie rule (always on under a caller-supplied rand.Intn)
    const Hooks* hooks = nullptr;    // tests; a supplied intn forces one thread
    int device_slot = -1;            // >= 0: every GPU call of this search on that one device (root-parallel: one tree per device)
    uint64_t stream_base = 0;        // added to the random-stream ids of this search: searchers of one decision must not share streams
    uint64_t rng_seed = 0;           // != 0: seeds this search's host RNGs (else they are drawn from the process-wide host RNG)
};
struct FastResult {
    int best_move, value;            // argmax UCB1 over the root's children, int(1000*UCB1) (mcts.go:196-198)
    long long iterations, batches;
    long long huge_page_kb;           // AnonHugePages of the process when the search ended (-1: not readable): did the arena get 2 MB pages?
    int n_children, child_move[25];  // root children in creation order
    double child_visits[25], child_wins[25], root_visits;
    double select_seconds, gpu_seconds, update_seconds;
    uint64_t nodes;                  // arena slots handed out
    int threads;
};
FastResult fast_uct(const GameState& rootstate, long long itermax, double UCTK, const FastOptions& opt);

// Root-parallel MCTS (the shape of squavam2.go:93-112 -- N searchers, ONE decision -- without its shared, locked
// tree): n_searchers independent arena trees are grown from the same root, searcher i on device slot
// i % sq_device_count() with its own share of the host threads, its own host RNG and its own random streams,
// itermax / n_searchers iterations each.  The root children's {visits, wins} are then summed move by move and the
// move returned is argmax UCB1 over the merged statistics (mcts.go:196-198,218-238).  One process drives all the
// devices, so the merge is a sum in host memory; under torchrun (one process per GPU, bench.py) the same 2 x 25
// vector is merged with one NCCL all-reduce instead.
struct RootParallelResult {
    FastResult merged;               // child_* merged by move (ascending move order), best_move / value from the merged statistics
    std::vector<FastResult> searcher;
    double seconds;
};
RootParallelResult root_parallel_uct(const GameState& rootstate, long long itermax, double UCTK, const FastOptions& opt,
                                     int n_searchers);

class MCTS : public Player {           // mcts.go:32-90
public:
    MCTS(bool deterministic, int maxdepth);
    std::string Name() override { return "MCTS"; }
    void SetUCTK(double k) { UCTK_ = k; }
    void SetIterations(int n) { iterations_ = n; }
    void SetBatch(int leaves, int playouts_per_leaf) { opt_.batch = leaves; opt_.playouts_per_leaf = playouts_per_leaf; }
    void MakeMove(int x, int y, int player) override;
    void SetDepth(int) override {}
    Move ChooseMove() override;
    void PrintBoard() override;
    void SetScores(bool) override {}
    int FindWinner() override;
private:
    GameState game_;
    int iterations_ = 500000;
    std::unique_ptr<Node> movesNode_;
    double UCTK_ = 1.0;
    Options opt_;
};

}  // namespace mcts

// ---- alpha-beta ---------------------------------------------------------------------
namespace alphabeta {

class AlphaBeta : public Player, public BatchSearchable {      // alphabeta.go:20-117
public:
    AlphaBeta(bool deterministic, int maxdepth);
    std::string Name() override { return name_; }
    void MakeMove(int x, int y, int player) override { bd_.set(5 * x + y, player); }
    void SetDepth(int moveCounter) override;
    Move ChooseMove() override;
    void PrintBoard() override;
    void SetScores(bool randomize) override;
    int FindWinner() override;
    void SetAvoid() { name_ = "A/B+Avoid"; dialect_ = SQ_AB_ALPHABETA_AVOID; }   // alphabeta.go:473-475
    SearchRequest Request() override { return SearchRequest{bd_.raw(), dialect_, maxDepth_, scores_}; }
    Move Finish(const SearchResult& r) override;
    const Board& board() const { return bd_; }
protected:
    Board bd_;
    std::string name_ = "AlphaBeta";
    int maxDepth_;
    bool deterministic_;
    int dialect_ = SQ_AB_ALPHABETA;
    int8_t* scores_;         // the package-level bias table this engine reads and SetScores writes
};
extern int8_t scores[25];    // package-level, shared by all instances (alphabeta.go:330)

}  // namespace alphabeta

// ---- alpha-beta with the "triangle" opening book ---------------------------------------
namespace abbook {

// src/abbook/abbook.go:21-131,448-575.  The book is a small state machine on the host; once it
// is exhausted the engine is src/alphabeta's search with boardValue+delta handed down
// (dialect SQ_AB_ABBOOK) and its own SetDepth table and bias table.
class AlphaBetaBook : public alphabeta::AlphaBeta {
public:
    AlphaBetaBook(bool deterministic, int maxdepth);
    void MakeMove(int x, int y, int player) override { moveCount_++; bd_.set(5 * x + y, player); }   // :75-78
    void SetDepth(int moveCounter) override;
    Move ChooseMove() override;
    bool BookMove(Move* mv) override;
    bool bookInProgress() const { return bookInProgress_; }
    void SetBookPick(int k) { pick_ = k; }      // tests: stands in for rand.Intn(4) (abbook.go:526); -1 = random
private:
    enum State { FIRST, DIAGONAL, CORNER, OTHERCORNER, OTHERDIAGONAL, BLOCKDIAG1, BLOCKDIAG2, LAST };
    void bookDefend();
    void bookStart();
    int moveCount_ = 0, state_ = FIRST, c_x_ = 0, c_y_ = 0, firstX_ = 0, firstY_ = 0;
    bool bookInProgress_ = true;
    int pick_ = -1;
};
extern int8_t scores[25];    // abbook.go:379

}  // namespace abbook

// ---- NegaScout ---------------------------------------------------------------------------
namespace negascout {

extern int8_t scores[25];    // negascout.go:369

// src/negascout/negascout.go:20-122.  The whole ChooseMove search is one sq_negascout_root call.
class NegaScout : public Player, public BatchSearchable {
public:
    NegaScout(bool deterministic, int maxdepth) : maxDepth_(maxdepth), deterministic_(deterministic) {}
    std::string Name() override { return "NegaScout"; }
    void MakeMove(int x, int y, int player) override { bd_.set(5 * x + y, player); }
    void SetDepth(int moveCounter) override;
    Move ChooseMove() override;
    void PrintBoard() override;
    void SetScores(bool randomize) override;
    int FindWinner() override;
    SearchRequest Request() override { return SearchRequest{bd_.raw(), kSearchNegaScout, maxDepth_, scores}; }
    Move Finish(const SearchResult& r) override;
private:
    Board bd_;
    int maxDepth_;
    bool deterministic_;
};

}  // namespace negascout

// ---- tiny Go-style flag parser ("-i 100", "-i=100", "-C", "-C=false") ----------------
class Flags {
public:
    void Int(const char* name, int* v, const char* help);
    void Float(const char* name, double* v, const char* help);
    void Bool(const char* name, bool* v, const char* help);
    void String(const char* name, std::string* v, const char* help);
    void Parse(int argc, char** argv);
private:
    struct F { std::string name, help; int kind; void* ptr; };
    std::vector<F> flags_;
    void usage(const char* prog);
};

// common start-up: sq_init on the devices named by -gpus / SQUAVA_GPUS (default device 0)
void init_devices(int n_gpus, uint64_t seed);

}  // namespace squava
