// playoff5 -- engine-vs-engine harness over the Player interface (playoff5.go:31-126),
// with all five engine types: A (alpha-beta), N (NegaScout), B (alpha-beta + opening book),
// G (alpha-beta + the "avoid" evaluator) and M (MCTS).
#include <chrono>
#include <cstdio>
#include <memory>
#include <vector>

#include "squava_host.hpp"

using namespace squava;

static Player* create(const std::string& t, int maxDepth, bool det) {     // playoff5.go:185-221
    if (t == "A" || t == "a") return new alphabeta::AlphaBeta(det, maxDepth);
    if (t == "N" || t == "n") return new negascout::NegaScout(det, maxDepth);
    if (t == "B" || t == "b") return new abbook::AlphaBetaBook(det, maxDepth);
    if (t == "G" || t == "g") { auto* p = new alphabeta::AlphaBeta(det, maxDepth); p->SetAvoid(); return p; }
    if (t == "M" || t == "m") return new mcts::MCTS(det, maxDepth);
    std::fprintf(stderr, "panic: runtime error: invalid memory address or nil pointer dereference (player type %s)\n", t.c_str());
    std::exit(2);                                   // the reference leaves the Player nil (playoff5.go:192-217)
}

// nonInteractiveGames, playoff5.go:128-183 -- quirks kept: `randomize` is passed where
// createPlayers expects `deterministic`, SetScores is never called (the bias table stays
// zero), MCTS players keep their defaults, values[][] is indexed by the move counter.
static void nonInteractiveGames(int gameCount, const std::string& firstType, const std::string& secondType,
                                bool randomize, int maxDepth) {
    for (int g = 0; g < gameCount; g++) {
        int moveCounter = 0;
        std::unique_ptr<Player> first(create(firstType, maxDepth, randomize)), second(create(secondType, maxDepth, randomize));
        std::printf("%d %s %s %d %s ", g, first->Name().c_str(), second->Name().c_str(), maxDepth, randomize ? "true" : "false");
        int moves[25][2] = {{0}}, values[25][2] = {{0}};
        int winner = 0;
        while (moveCounter < 25) {
            first->SetDepth(moveCounter);
            Move mv = first->ChooseMove();
            moves[moveCounter][0] = mv.x; moves[moveCounter][1] = mv.y;
            values[moveCounter][0] = mv.value;
            second->MakeMove(mv.x, mv.y, MINIMIZER);
            moveCounter++;
            winner = first->FindWinner();
            if (winner != 0 || moveCounter >= 25) break;
            second->SetDepth(moveCounter);
            mv = second->ChooseMove();
            moves[moveCounter][0] = mv.x; moves[moveCounter][1] = mv.y;
            values[moveCounter][1] = mv.value;
            first->MakeMove(mv.x, mv.y, MINIMIZER);
            moveCounter++;
            winner = -second->FindWinner();
            if (winner != 0) break;
        }
        std::printf("%d %d", moveCounter, winner);
        for (int i = 0; i < moveCounter; i++) {
            const char* marker[2] = {"", ""};
            for (int j = 0; j < 2; j++) {
                if (values[i][j] > 9000) marker[j] = "+";
                if (values[i][j] < -9000) marker[j] = "-";
            }
            std::printf(" %d%s,%d%s", moves[i][0], marker[0], moves[i][1], marker[1]);
        }
        std::printf("\n");
        std::fflush(stdout);
    }
}

// The same games, advanced in LOCKSTEP: at every ply the engines to move in all live games go
// through one batched library call per engine type (batched engine-vs-engine
// evaluation).  Game lines are identical in format; with -records the playoff4 record
// (playoff4.go:119-123: winner, move count, moves, tab-separated) is printed instead.
// MCTS players keep their own trees and are stepped one game at a time.
struct LockstepGame {
    std::unique_ptr<Player> p[2];
    int moves[25][2] = {{0}}, values[25][2] = {{0}};
    int moveCounter = 0, winner = 0;
    bool live = true;
};

static void print_game(int g, const LockstepGame& G, int maxDepth, bool randomize, bool records) {
    if (records) {
        std::printf("%s\t%d", G.winner == 1 ? "X" : G.winner == -1 ? "O" : "C", G.moveCounter);
        for (int i = 0; i < G.moveCounter; i++) std::printf("\t%d,%d", G.moves[i][0], G.moves[i][1]);
        std::printf("\n");
        return;
    }
    std::printf("%d %s %s %d %s ", g, G.p[0]->Name().c_str(), G.p[1]->Name().c_str(), maxDepth, randomize ? "true" : "false");
    std::printf("%d %d", G.moveCounter, G.winner);
    for (int i = 0; i < G.moveCounter; i++) {
        const char* marker[2] = {"", ""};
        for (int j = 0; j < 2; j++) {
            if (G.values[i][j] > 9000) marker[j] = "+";
            if (G.values[i][j] < -9000) marker[j] = "-";
        }
        std::printf(" %d%s,%d%s", G.moves[i][0], marker[0], G.moves[i][1], marker[1]);
    }
    std::printf("\n");
}

static void lockstepGames(int gameCount, const std::string& firstType, const std::string& secondType,
                          bool randomize, int maxDepth, bool records) {
    std::vector<LockstepGame> games((size_t)gameCount);
    for (auto& G : games) { G.p[0].reset(create(firstType, maxDepth, randomize)); G.p[1].reset(create(secondType, maxDepth, randomize)); }
    for (int ply = 0; ply < 25; ply++) {
        const int side = ply & 1;
        std::vector<SearchRequest> req;
        std::vector<LockstepGame*> who;
        std::vector<Move> made(games.size());
        std::vector<char> have(games.size(), 0);
        for (size_t g = 0; g < games.size(); g++) {
            LockstepGame& G = games[g];
            if (!G.live) continue;
            Player* me = G.p[side].get();
            me->SetDepth(G.moveCounter);
            auto* bs = dynamic_cast<BatchSearchable*>(me);
            if (!bs) { made[g] = me->ChooseMove(); have[g] = 1; continue; }       // MCTS
            if (bs->BookMove(&made[g])) { have[g] = 1; continue; }
            req.push_back(bs->Request());
            who.push_back(&G);
        }
        std::vector<SearchResult> res;
        search_batch(req, res);
        for (size_t k = 0; k < who.size(); k++) {
            const size_t g = (size_t)(who[k] - &games[0]);
            made[g] = dynamic_cast<BatchSearchable*>(who[k]->p[side].get())->Finish(res[k]);
            have[g] = 1;
        }
        bool any = false;
        for (size_t g = 0; g < games.size(); g++) {
            LockstepGame& G = games[g];
            if (!G.live || !have[g]) continue;
            const Move& mv = made[g];
            G.moves[G.moveCounter][0] = mv.x; G.moves[G.moveCounter][1] = mv.y;
            G.values[G.moveCounter][side] = mv.value;
            G.p[1 - side]->MakeMove(mv.x, mv.y, MINIMIZER);
            G.moveCounter++;
            G.winner = side == 0 ? G.p[0]->FindWinner() : -G.p[1]->FindWinner();
            if (G.winner != 0 || (side == 0 && G.moveCounter >= 25)) G.live = false;
            any |= G.live;
        }
        if (!any) break;
    }
    for (size_t g = 0; g < games.size(); g++) print_game((int)g, games[g], maxDepth, randomize, records);
    std::fflush(stdout);
}

int main(int argc, char** argv) {
    int maxDepth = 10, i1 = 500000, i2 = 500000, gpus = 0, seed = -1, games = 1;
    bool deterministic = false, randomize = false, lockstep = false, records = false;
    std::string firstType = "A", secondType = "M";
    double u1 = 0.50, u2 = 0.50;
    Flags fl;
    fl.Int("d", &maxDepth, "maximum lookahead depth (alpha/beta)");
    fl.Bool("D", &deterministic, "Play deterministically");
    fl.Bool("r", &randomize, "Randomize bias scores");
    fl.String("1", &firstType, "first player type, A: alphabeta, N: negascout, B: A/B+book opening, G: A/B+avoid bad positions, M: MCTS");
    fl.String("2", &secondType, "second player type, A: alphabeta, N: negascout, B: A/B+book opening, G: A/B+avoid bad positions, M: MCTS");
    fl.Int("n", &games, "play <number> games non-interactively");
    fl.Float("u1", &u1, "UCTK coefficient, player 1 (MCTS)");
    fl.Float("u2", &u2, "UCTK coefficient, player 2 (MCTS)");
    fl.Int("i1", &i1, "MCTS iterations, player 1");
    fl.Int("i2", &i2, "MCTS iterations, player 2");
    fl.Bool("lockstep", &lockstep, "with -n: advance all games together, one batched GPU call per ply");
    fl.Bool("records", &records, "with -lockstep: print playoff4-style game records (winner, moves, tab-separated)");
    fl.Int("gpus", &gpus, "number of GPUs (default 1)");
    fl.Int("seed", &seed, "seed for reproducible play");
    fl.Parse(argc, argv);
    if (seed >= 0) seed_host_rng((uint64_t)seed);
    init_devices(gpus, (uint64_t)host_rng()());
    if (games > 1) {                                  // playoff5.go:48-51
        if (lockstep) lockstepGames(games, firstType, secondType, randomize, maxDepth, records);
        else nonInteractiveGames(games, firstType, secondType, randomize, maxDepth);
        sq_shutdown();
        return 0;
    }
    std::unique_ptr<Player> first(create(firstType, maxDepth, deterministic)), second(create(secondType, maxDepth, deterministic));
    if (auto* m = dynamic_cast<mcts::MCTS*>(first.get())) { m->SetUCTK(u1); m->SetIterations(i1); }
    if (auto* m = dynamic_cast<mcts::MCTS*>(second.get())) { m->SetUCTK(u2); m->SetIterations(i2); }
    first->SetScores(randomize);
    second->SetScores(randomize);
    int winner = 0, moveCounter = 0;
    auto gameStart = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t) {
        return go_duration(std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count());
    };
    while (moveCounter < 25) {
        first->SetDepth(moveCounter);
        auto before = std::chrono::steady_clock::now();
        Move mv = first->ChooseMove();
        std::string et = since(before);
        second->MakeMove(mv.x, mv.y, MINIMIZER);
        moveCounter++;
        std::printf("X (%s) <%d,%d> (%d) [%d] %s\n", first->Name().c_str(), mv.x, mv.y, mv.value, mv.leafCount, et.c_str());
        winner = first->FindWinner();
        if (winner != 0 || moveCounter >= 25) break;
        second->SetDepth(moveCounter);
        before = std::chrono::steady_clock::now();
        mv = second->ChooseMove();
        et = since(before);
        first->MakeMove(mv.x, mv.y, MINIMIZER);
        moveCounter++;
        std::printf("O (%s) <%d,%d> (%d) [%d] %s\n", second->Name().c_str(), mv.x, mv.y, mv.value, mv.leafCount, et.c_str());
        first->PrintBoard();
        int winner1 = first->FindWinner();
        int winner2 = -second->FindWinner();
        if (winner1 != winner2) std::printf("Winner disagreement. First %d, second %d\n", winner1, winner2);
        if (winner2 != 0) { winner = winner2; break; }
        std::fflush(stdout);
    }
    std::string gameET = since(gameStart);
    if (winner == 1) std::printf("player 1 X (%s) wins, %s\n", first->Name().c_str(), gameET.c_str());
    else if (winner == -1) std::printf("player 2 O (%s) wins, %s\n", second->Name().c_str(), gameET.c_str());
    else std::printf("Cat wins\n");
    first->PrintBoard();
    sq_shutdown();
    return 0;
}
