// squava_host.cpp -- see squava_host.hpp.  Host-side engine surface over the C ABI.
#include "squava_host.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

namespace squava {

// ---- Board printing ----------------------------------------------------------------
std::string Board::str() const {            // squavam.go:327-339
    std::string s = "   0 1 2 3 4\n";
    for (int i = 0; i < 25; i++) {
        if (i % 5 == 0) { s += char('0' + i / 5); s += "  "; }
        s += "O_X"[at(i) + 1]; s += ' ';
        if (i % 5 == 4) s += '\n';
    }
    return s;
}
std::string Board::str_grid() const {       // alphabeta.go:226-247
    std::string s = "   0 1 2 3 4\n";
    for (int i = 0; i < 5; i++) {
        s += char('0' + i); s += "  ";
        for (int j = 0; j < 5; j++) { s += "O_X"[at(5 * i + j) + 1]; s += ' '; }
        s += '\n';
    }
    s += '\n';
    return s;
}

// ---- host RNG ------------------------------------------------------------------------
static std::mt19937_64 g_rng;
static bool g_rng_seeded = false;
void seed_host_rng(uint64_t seed) { g_rng.seed(seed); g_rng_seeded = true; }
std::mt19937_64& host_rng() {
    if (!g_rng_seeded) {
        const char* s = std::getenv("SQUAVA_SEED");
        uint64_t seed = s ? std::strtoull(s, nullptr, 0)
                          : (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();  // squavam.go:48
        seed_host_rng(seed);
    }
    return g_rng;
}
static int intn(std::mt19937_64& r, int n) { return (int)std::uniform_int_distribution<int>(0, n - 1)(r); }

std::string go_duration(double s) {       // time.Duration.String()
    char b[64];
    if (s < 1e-6) std::snprintf(b, sizeof b, "%.0fns", s * 1e9);
    else if (s < 1e-3) std::snprintf(b, sizeof b, "%.3fµs", s * 1e6);
    else if (s < 1.0) std::snprintf(b, sizeof b, "%.6fms", s * 1e3);
    else if (s < 60.0) std::snprintf(b, sizeof b, "%.9fs", s);
    else {
        int m = (int)(s / 60.0);
        std::snprintf(b, sizeof b, "%dm%.9fs", m, s - 60.0 * m);
    }
    return b;
}

// ---- MoveKeeper (movekeeper.go:26-52) ---------------------------------------------------
void MoveKeeper::SetMove(int a, int b, int value) {
    if (value >= max_) {
        if (value > max_) { max_ = value; next_ = 0; }
        moves_[next_][0] = a; moves_[next_][1] = b; next_++;
    }
}
void MoveKeeper::ChooseMove(int& x, int& y, int& value) {
    if (next_ == 0) { x = -1; y = -1; value = 0; return; }   // cat got the game
    int r = det_ ? 0 : intn(*rng_, next_);
    x = moves_[r][0]; y = moves_[r][1]; value = max_;
}

// ---- flags ------------------------------------------------------------------------------
void Flags::Int(const char* n, int* v, const char* h) { flags_.push_back({n, h, 0, v}); }
void Flags::Float(const char* n, double* v, const char* h) { flags_.push_back({n, h, 1, v}); }
void Flags::Bool(const char* n, bool* v, const char* h) { flags_.push_back({n, h, 2, v}); }
void Flags::String(const char* n, std::string* v, const char* h) { flags_.push_back({n, h, 3, v}); }
void Flags::usage(const char* prog) {
    std::fprintf(stderr, "Usage of %s:\n", prog);
    for (auto& f : flags_) std::fprintf(stderr, "  -%s\n    \t%s\n", f.name.c_str(), f.help.c_str());
}
void Flags::Parse(int argc, char** argv) {
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.size() < 2 || a[0] != '-') break;
        a = a.substr(a[1] == '-' ? 2 : 1);
        std::string val; bool has_val = false;
        auto eq = a.find('=');
        if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        F* f = nullptr;
        for (auto& c : flags_) if (c.name == a) f = &c;
        if (!f) {
            if (a != "h" && a != "help") std::fprintf(stderr, "flag provided but not defined: -%s\n", a.c_str());
            usage(argv[0]); std::exit(2);
        }
        if (f->kind == 2) {
            *(bool*)f->ptr = !has_val || (val != "false" && val != "0" && val != "f" && val != "F" && val != "FALSE");
            continue;
        }
        if (!has_val) {
            if (i + 1 >= argc) { std::fprintf(stderr, "flag needs an argument: -%s\n", a.c_str()); usage(argv[0]); std::exit(2); }
            val = argv[++i];
        }
        if (f->kind == 0) *(int*)f->ptr = std::atoi(val.c_str());
        else if (f->kind == 1) *(double*)f->ptr = std::atof(val.c_str());
        else *(std::string*)f->ptr = val;
    }
}

void init_devices(int n_gpus, uint64_t seed) {
    const char* e = std::getenv("SQUAVA_GPUS");
    if (n_gpus <= 0 && e) n_gpus = std::atoi(e);
    if (n_gpus <= 0) n_gpus = 1;
    std::vector<int> ids(n_gpus);
    for (int i = 0; i < n_gpus; i++) ids[i] = i;
    must(sq_init(ids.data(), n_gpus, seed), "sq_init");
}

// ---- MCTS ---------------------------------------------------------------------------------
namespace mcts {

static const double kEps = std::numeric_limits<double>::denorm_min();   // math.SmallestNonzeroFloat64

static double UCB1(const Node* p, double UCTK, bool factor2) {          // mcts.go:248-250 / squavam.go:197-199
    double lg = std::log(p->parentNode->visits);
    return p->wins / (p->visits + kEps) + UCTK * std::sqrt((factor2 ? 2. : 1.) * lg / (p->visits + kEps));
}

static Node* bestMove(Node* p, double UCTK, bool factor2) {             // mcts.go:218-238
    double bestscore = kEps;
    Node* best = nullptr;
    // same expression as UCB1() per child; ln(parent visits) is the same for all of them
    const double lg = (factor2 ? 2. : 1.) * std::log(p->visits);
    for (auto& c : p->childNodes) {
        double u = c->wins / (c->visits + kEps) + UCTK * std::sqrt(lg / (c->visits + kEps));
        if (u > bestscore) { bestscore = u; best = c.get(); }
    }
    return best;
}

static std::vector<int> empties_of(const Board& b) {
    std::vector<int> m;
    for (int i = 0; i < 25; i++) if (b.at(i) == UNSET) m.push_back(i);
    return m;
}

// NewNode (mcts.go:201-208).  untriedMoves needs GetMoves' terminal test; that test is the
// GPU's: for a node created inside a batch it is read off the batch result (a terminal leaf
// plays zero plies), so the node is `pending` until its batch returns.
struct Pending { Node* node; Board board; int to_move; };

Node* UCT(const GameState& rootstate, int itermax, double UCTK, std::unique_ptr<Node>& root,
          const Options& opt, int* leaves, int* value) {
    static uint64_t stream_id = 0;
    const Hooks* hk = opt.hooks;
    auto rand_intn = [&](int n) { return hk && hk->intn ? hk->intn(n) : intn(host_rng(), n); };
    if (!root) {
        root.reset(new Node());
        root->move = -1;
        root->playerJustMoved = rootstate.playerJustMoved;
        bool terminal;
        if (hk && hk->terminal) terminal = hk->terminal(rootstate.board);
        else {
            uint8_t flags = 0; sq_board rb = rootstate.board.raw();
            must(sq_classify(&rb, 1, &flags, nullptr, nullptr), "sq_classify");
            terminal = flags != 0;
        }
        if (!terminal) root->untriedMoves = empties_of(rootstate.board);
    } else {
        root->playerJustMoved = rootstate.playerJustMoved;     // mcts.go:149-151
    }
    int done = 0;
    std::vector<Pending> batch;
    std::vector<sq_board> boards;
    std::vector<int8_t> to_move;
    std::vector<uint64_t> out;
    std::vector<char> is_new;
    while (done < itermax) {
        int P = opt.playouts_per_leaf < 1 ? 1 : opt.playouts_per_leaf;
        int remaining = itermax - done;
        if (P > remaining) P = remaining;
        int B = opt.batch < 1 ? 1 : opt.batch;
        if ((long long)B * P > remaining) B = remaining / P;
        batch.clear(); is_new.clear();
        for (int k = 0; k < B; k++) {
            Node* node = root.get();
            GameState state = rootstate;                      // Clone, mcts.go:156
            // select (mcts.go:158-161); a node still waiting for its first batch is a leaf
            while (node->untriedMoves.empty() && !node->childNodes.empty()) {
                Node* nx = bestMove(node, UCTK, opt.ucb_factor2);
                if (!nx) {
                    std::fprintf(stderr, "Node has %zu children\nplayer just moved %d\n", node->childNodes.size(), node->playerJustMoved);
                    std::exit(2);                             // the reference dereferences nil here
                }
                node = nx;
                state.DoMove(node->move);
            }
            bool fresh = false;
            if (!node->untriedMoves.empty()) {                // expand (mcts.go:166-171)
                int idx = rand_intn((int)node->untriedMoves.size());
                int m = node->untriedMoves[idx];
                state.DoMove(m);
                node->untriedMoves.erase(node->untriedMoves.begin() + idx);
                Node* c = new Node();
                c->move = m; c->parentNode = node; c->playerJustMoved = state.playerJustMoved;
                node->childNodes.emplace_back(c);
                node = c;
                fresh = true;
            }
            for (Node* n = node; n; n = n->parentNode) n->visits += P;   // counted now (virtual loss), wins on return
            batch.push_back({node, state.board, -state.playerJustMoved});
            is_new.push_back(fresh);
        }
        boards.resize(B); to_move.resize(B); out.assign((size_t)B * 4, 0);
        for (int k = 0; k < B; k++) { boards[k] = batch[k].board.raw(); to_move[k] = (int8_t)batch[k].to_move; }
        if (hk && hk->playouts) hk->playouts(boards.data(), to_move.data(), B, (uint64_t)P, stream_id++, (uint64_t(*)[4])out.data());
        else must(sq_playouts(boards.data(), to_move.data(), B, (uint64_t)P, stream_id++, (uint64_t(*)[4])out.data()), "sq_playouts");
        for (int k = 0; k < B; k++) {
            const uint64_t* r = &out[(size_t)k * 4];
            if (is_new[k] && r[3] != 0) batch[k].node->untriedMoves = empties_of(batch[k].board);  // non-terminal
            for (Node* n = batch[k].node; n; n = n->parentNode)           // Update, mcts.go:190-192,268-271
                n->wins += (double)(n->playerJustMoved == MAXIMIZER ? r[0] : r[1]);
        }
        done += B * P;
    }
    if (leaves) *leaves = done;
    Node* choice = bestMove(root.get(), UCTK, opt.ucb_factor2);
    if (!choice) { std::fprintf(stderr, "Node has %zu children\n", root->childNodes.size()); std::exit(2); }
    if (value) *value = (int)(1000. * UCB1(choice, UCTK, opt.ucb_factor2));
    return choice;
}

// detach `child` from *root and make it the new root (movesNode.parentNode = nil)
static void descend(std::unique_ptr<Node>& root, Node* child) {
    std::unique_ptr<Node> keep;
    for (auto& c : root->childNodes) if (c.get() == child) keep = std::move(c);
    keep->parentNode = nullptr;
    root = std::move(keep);
}

MCTS::MCTS(bool, int) { game_.playerJustMoved = -1; }           // mcts.go:39-41,273-277

void MCTS::MakeMove(int x, int y, int player) {                // mcts.go:55-59,297-308
    int m = 5 * x + y;
    game_.board.set(m, player);
    game_.playerJustMoved = player;
    if (movesNode_) {
        Node* hit = nullptr;
        for (auto& c : movesNode_->childNodes) if (c->move == m) { hit = c.get(); break; }
        if (hit) descend(movesNode_, hit);
        else movesNode_.reset();      // deviation: the reference keeps the stale tree
    }
}

Move MCTS::ChooseMove() {                                       // mcts.go:66-82
    int leaves = 0, value = 0;
    Node* best = UCT(game_, iterations_, UCTK_, movesNode_, opt_, &leaves, &value);
    int move = best->move;
    descend(movesNode_, best);
    game_.DoMove(move);
    return Move{move / 5, move % 5, value, leaves};
}

void MCTS::PrintBoard() { std::printf("%s", game_.board.str().c_str()); }

int MCTS::FindWinner() {                                        // mcts.go:92-121
    sq_board b = game_.board.raw(); int8_t w = 0;
    must(sq_classify(&b, 1, nullptr, &w, nullptr), "sq_classify");
    return w;
}

}  // namespace mcts

// ---- alpha-beta ---------------------------------------------------------------------------
// One library call per (kind, depth, bias table) group of requests.
void search_batch(const std::vector<SearchRequest>& req, std::vector<SearchResult>& res) {
    const size_t n = req.size();
    res.assign(n, SearchResult());
    std::vector<char> done(n, 0);
    for (size_t i = 0; i < n; i++) {
        if (done[i]) continue;
        std::vector<size_t> grp;
        for (size_t j = i; j < n; j++)
            if (!done[j] && req[j].kind == req[i].kind && req[j].depth == req[i].depth && req[j].scores == req[i].scores) {
                grp.push_back(j); done[j] = 1;
            }
        const size_t m = grp.size();
        std::vector<sq_board> b(m);
        std::vector<int32_t> values(m * 25), bv(m);
        std::vector<uint32_t> bm(m);
        std::vector<uint64_t> nodes(m);
        std::vector<int8_t> order(m * 25, -1), first(m, -1);
        for (size_t k = 0; k < m; k++) b[k] = req[grp[k]].board;
        if (req[i].kind == kSearchNegaScout)
            must(sq_negascout_root(b.data(), (int64_t)m, req[i].depth, req[i].scores, (int32_t(*)[25])values.data(),
                                   (int8_t(*)[25])order.data(), bv.data(), bm.data(), first.data(), nodes.data()), "sq_negascout_root");
        else
            must(sq_alphabeta_root(b.data(), (int64_t)m, req[i].kind, req[i].depth, req[i].scores, (int32_t(*)[25])values.data(),
                                   bv.data(), bm.data(), nodes.data()), "sq_alphabeta_root");
        for (size_t k = 0; k < m; k++) {
            SearchResult& r = res[grp[k]];
            std::memcpy(r.values, &values[k * 25], sizeof r.values);
            std::memcpy(r.visit_order, &order[k * 25], sizeof r.visit_order);
            r.best_value = bv[k]; r.best_mask = bm[k]; r.first_best = first[k]; r.nodes = nodes[k];
        }
    }
}

namespace alphabeta {

int8_t scores[25] = {0};

AlphaBeta::AlphaBeta(bool deterministic, int maxdepth)
    : maxDepth_(maxdepth), deterministic_(deterministic), scores_(scores) {}

void AlphaBeta::SetDepth(int moveCounter) {       // alphabeta.go:79-89
    if (moveCounter < 4) maxDepth_ = 8;
    if (moveCounter > 3) maxDepth_ = 10;
    if (moveCounter > 10) maxDepth_ = 12;
}

Move AlphaBeta::ChooseMove() {                    // alphabeta.go:92-117
    std::vector<SearchRequest> req{Request()};
    std::vector<SearchResult> res;
    search_batch(req, res);
    return Finish(res[0]);
}

Move AlphaBeta::Finish(const SearchResult& r) {   // moves.SetMove in row-major order, ChooseMove, MakeMove (:97-116)
    MoveKeeper moves(2 * LOSS, deterministic_, &host_rng());
    for (int c = 0; c < 25; c++) if (r.values[c] != SQ_NO_MOVE) moves.SetMove(c / 5, c % 5, r.values[c]);
    int a, bb, v;
    moves.ChooseMove(a, bb, v);
    if (a < 0) { std::fprintf(stderr, "panic: index out of range [-1]\n"); std::exit(2); }   // the reference indexes bd[-1]
    MakeMove(a, bb, MAXIMIZER);
    return Move{a, bb, v, (int)r.nodes};
}

void AlphaBeta::PrintBoard() { std::printf("%s", bd_.str_grid().c_str()); }

void AlphaBeta::SetScores(bool randomize) {        // alphabeta.go:334-351
    if (randomize) {
        static const int vals[11] = {-5, -4, -3 - 2, -1, 0, 1, 2, 3, 4, 5, 0};
        for (int i = 0; i < 25; i++) scores_[i] = (int8_t)vals[intn(host_rng(), 11)];
    } else {
        static const int8_t d[25] = {3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3};
        std::memcpy(scores_, d, 25);
    }
}

int AlphaBeta::FindWinner() {                      // alphabeta.go:355-378
    sq_board b = bd_.raw(); int8_t w = 0;
    must(sq_classify(&b, 1, nullptr, nullptr, &w), "sq_classify");
    return w;
}

}  // namespace alphabeta

// ---- src/abbook ------------------------------------------------------------------------
namespace abbook {

int8_t scores[25];

AlphaBetaBook::AlphaBetaBook(bool deterministic, int maxdepth) : AlphaBeta(deterministic, maxdepth) {
    name_ = "A/B+Book";
    dialect_ = SQ_AB_ABBOOK;
    scores_ = abbook::scores;
}

void AlphaBetaBook::SetDepth(int moveCounter) {    // abbook.go:80-90
    if (moveCounter < 4) maxDepth_ = 6;
    if (moveCounter > 3) maxDepth_ = 8;
    if (moveCounter > 10) maxDepth_ = 10;
}

bool AlphaBetaBook::BookMove(Move* mv) {           // abbook.go:95-105
    if (!bookInProgress_) return false;
    if (moveCount_ % 2 == 0) bookStart(); else bookDefend();
    if (c_x_ == -1 || c_y_ == -1) return false;
    MakeMove(c_x_, c_y_, MAXIMIZER);
    *mv = Move{c_x_, c_y_, 0, 0};
    return true;
}

Move AlphaBetaBook::ChooseMove() {                 // abbook.go:93-131
    Move mv;
    if (BookMove(&mv)) return mv;
    return AlphaBeta::ChooseMove();                // the same loop with dialect 6; MakeMove is virtual
}

// Answer the opponent's first two moves on the cell a "3 away on both axes" from
// them (abbook.go:455-516).  Quirk kept: when no such cell exists, c_x/c_y keep their previous
// value (0,0 at first) and that stale cell is played.
void AlphaBetaBook::bookDefend() {
    int lastx = -1, lasty = -1;
    if (state_ == FIRST) {
        state_ = BLOCKDIAG1;
        for (int c = 0; c < 25 && lastx < 0; c++) if (bd_.at(c) != UNSET) { lastx = c / 5; lasty = c % 5; }
        firstX_ = lastx; firstY_ = lasty;
        for (int i = -3; i < 4; i += 6) for (int j = -3; j < 4; j += 6) {
            int a = lastx + i, b = lasty + j;
            if (a >= 0 && a <= 4 && b >= 0 && b <= 4) { c_x_ = a; c_y_ = b; return; }
        }
    } else if (state_ == BLOCKDIAG1) {
        state_ = LAST;
        bookInProgress_ = false;
        for (int c = 0; c < 25 && lastx < 0; c++)
            if (!(c / 5 == firstX_ && c % 5 == firstY_) && bd_.at(c) == MINIMIZER) { lastx = c / 5; lasty = c % 5; }
        for (int i = -3; i < 4; i += 6) for (int j = -3; j < 4; j += 6) {
            int a = lastx + i, b = lasty + j;
            if (a >= 0 && a <= 4 && b >= 0 && b <= 4 && bd_.at(5 * a + b) == UNSET) { c_x_ = a; c_y_ = b; return; }
        }
    }
}

// The "triangle" opening (abbook.go:518-575).
void AlphaBetaBook::bookStart() {
    static const int firstMoves[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
    bool looping = true;
    while (looping) {
        looping = false;
        switch (state_) {
        case FIRST: {
            const int* c = firstMoves[pick_ >= 0 ? pick_ % 4 : intn(host_rng(), 4)];
            c_x_ = c[0]; c_y_ = c[1];
            state_ = DIAGONAL;
            break;
        }
        case DIAGONAL:
            if (bd_.at(5 * (c_x_ + 3) + c_y_ + 3) == UNSET) { c_x_ += 3; c_y_ += 3; state_ = CORNER; }
            else { state_ = OTHERCORNER; looping = true; }
            break;
        case CORNER:
            if (bd_.at(5 * (c_x_ - 3) + c_y_) == UNSET) c_x_ -= 3;
            else if (bd_.at(5 * c_x_ + c_y_ - 3) == UNSET) c_y_ -= 3;
            state_ = LAST;
            bookInProgress_ = false;
            break;
        case OTHERCORNER:
            if (bd_.at(5 * c_x_ + c_y_ + 3) == UNSET) { c_y_ += 3; state_ = OTHERDIAGONAL; }
            else {
                std::printf("Unreachable state in OTHERCORNER\n");
                PrintBoard();
                std::exit(99);
            }
            break;
        case OTHERDIAGONAL:
            if (bd_.at(5 * (c_x_ + 3) + c_y_ - 3) == UNSET) { c_x_ += 3; c_y_ -= 3; }
            else { c_x_ = -1; c_y_ = -1; }
            state_ = LAST;
            bookInProgress_ = false;
            break;
        default:
            break;
        }
    }
}

}  // namespace abbook

// ---- src/negascout -----------------------------------------------------------------------
namespace negascout {

int8_t scores[25];

void NegaScout::SetDepth(int moveCounter) {        // negascout.go:69-79
    if (moveCounter < 4) maxDepth_ = 6;
    if (moveCounter > 3) maxDepth_ = 8;
    if (moveCounter > 10) maxDepth_ = 10;
}

Move NegaScout::ChooseMove() {                     // negascout.go:81-122
    std::vector<SearchRequest> req{Request()};
    std::vector<SearchResult> res;
    search_batch(req, res);
    return Finish(res[0]);
}

Move NegaScout::Finish(const SearchResult& r) {
    if (r.first_best < 0) { std::fprintf(stderr, "panic: index out of range [-1]\n"); std::exit(2); }   // MakeMove(-1, -1, ...)
    int cell = r.first_best;
    if (!deterministic_) {                         // movekeeper.go:46-49: a uniform member of the kept moves, in visit order
        int kept[25], nk = 0;
        for (int k = 0; k < 25 && r.visit_order[k] >= 0; k++) if ((r.best_mask >> r.visit_order[k]) & 1u) kept[nk++] = r.visit_order[k];
        cell = kept[intn(host_rng(), nk)];
    }
    MakeMove(cell / 5, cell % 5, MAXIMIZER);
    return Move{cell / 5, cell % 5, r.best_value, (int)r.nodes};
}

void NegaScout::PrintBoard() { std::printf("%s", bd_.str_grid().c_str()); }

void NegaScout::SetScores(bool randomize) {        // negascout.go:371-388
    if (randomize) {
        static const int vals[11] = {-5, -4, -3 - 2, -1, 0, 1, 2, 3, 4, 5, 0};
        for (int i = 0; i < 25; i++) scores[i] = (int8_t)vals[intn(host_rng(), 11)];
    } else {
        static const int8_t d[25] = {3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3};
        std::memcpy(scores, d, 25);
    }
}

int NegaScout::FindWinner() {                      // negascout.go:124-145: the flat-list scan of src/alphabeta
    sq_board b = bd_.raw(); int8_t w = 0;
    must(sq_classify(&b, 1, nullptr, nullptr, &w), "sq_classify");
    return w;
}

}  // namespace negascout
}  // namespace squava
