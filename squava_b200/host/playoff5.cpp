// playoff5 -- engine-vs-engine harness over the Player interface (playoff5.go:31-126),
// restricted to the two engines on the accelerated path: A (alpha-beta) and M (MCTS).
#include <chrono>
#include <cstdio>

#include "squava_host.hpp"

using namespace squava;

static Player* create(const std::string& t, int maxDepth, bool det) {     // playoff5.go:185-221
    if (t == "A" || t == "a") return new alphabeta::AlphaBeta(det, maxDepth);
    if (t == "M" || t == "m") return new mcts::MCTS(det, maxDepth);
    std::fprintf(stderr, "player type %s is outside the accelerated path (A and M only)\n", t.c_str());
    std::exit(2);
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
    fl.String("1", &firstType, "first player type, A: alphabeta, M: MCTS");
    fl.String("2", &secondType, "second player type, A: alphabeta, M: MCTS");
    fl.Int("n", &games, "play <number> games non-interactively (only 1 supported)");
    fl.Float("u1", &u1, "UCTK coefficient, player 1 (MCTS)");
    fl.Float("u2", &u2, "UCTK coefficient, player 2 (MCTS)");
    fl.Int("i1", &i1, "MCTS iterations, player 1");
    fl.Int("i2", &i2, "MCTS iterations, player 2");
    fl.Int("gpus", &gpus, "number of GPUs (default 1)");
    fl.Int("seed", &seed, "seed for reproducible play");
    fl.Parse(argc, argv);
    if (seed >= 0) seed_host_rng((uint64_t)seed);
    init_devices(gpus, (uint64_t)host_rng()());
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
