// playoff5 -- engine-vs-engine harness over the Player interface (playoff5.go:31-126),
// with all five engine types: A (alpha-beta), N (NegaScout), B (alpha-beta + opening book),
// G (alpha-beta + the "avoid" evaluator) and M (MCTS).
#include <chrono>
#include <cstdio>

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

int main(int argc, char** argv) {
    int maxDepth = 10, i1 = 500000, i2 = 500000, gpus = 0, seed = -1, games = 1;
    bool deterministic = false, randomize = false;
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
    fl.Int("gpus", &gpus, "number of GPUs (default 1)");
    fl.Int("seed", &seed, "seed for reproducible play");
    fl.Parse(argc, argv);
    if (seed >= 0) seed_host_rng((uint64_t)seed);
    init_devices(gpus, (uint64_t)host_rng()());
    if (games > 1) {                                  // playoff5.go:48-51
        nonInteractiveGames(games, firstType, secondType, randomize, maxDepth);
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
