// squavathr -- drop-in for the reference's alpha-beta program with a root-move worker pool
// (squavathr.go).  chooseMove's toDo/finished round trip (squavathr.go:236-269,448-466)
// becomes ONE sq_alphabeta_root call (dialect 2); flags keep their names.  -N (goroutine
// count) is accepted and ignored; -B plays the reference's hard-coded "triangle" book on the
// host (squavathr.go:93-106,655-810) before the search takes over.
#include <chrono>
#include <cstdio>

#include "squava_host.hpp"

using namespace squava;

static int8_t scores[25];

static void setScores(bool randomize) {          // squavathr.go:612-630
    if (randomize) {
        static const int vals[11] = {-5, -4, -3 - 2, -1, 0, 1, 2, 3, 4, 5, 0};
        for (int i = 0; i < 25; i++) scores[i] = (int8_t)vals[std::uniform_int_distribution<int>(0, 10)(host_rng())];
    } else {
        static const int8_t d[25] = {3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3};
        for (int i = 0; i < 25; i++) scores[i] = d[i];
    }
}

static int setDepth(int moveCounter, int endGameDepth) {   // squavathr.go:183-198
    int maxDepth = 0;
    if (moveCounter < 4) maxDepth = 10;
    if (moveCounter > 3) maxDepth = 12;
    if (moveCounter > 10) maxDepth = endGameDepth;
    return maxDepth;
}

static int findWinner(const Board& bd) {                   // squavathr.go:271-294
    sq_board b = bd.raw(); int8_t w = 0;
    must(sq_classify(&b, 1, nullptr, nullptr, &w), "sq_classify");
    return w;
}

// deltaValue(100, &bd, 0, l, m, 0)'s stopRecursing for a move just made (squavathr.go:133,158):
// true exactly when the mover owns a line, which after a legal move means the board has one.
static bool game_over_after_move(const Board& bd) {
    sq_board b = bd.raw(); uint8_t f = 0;
    must(sq_classify(&b, 1, &f, nullptr, nullptr), "sq_classify");
    return (f & (SQ_X4 | SQ_O4 | SQ_X3 | SQ_O3)) != 0;
}

static void readMove(const Board& bd, bool printit, int& x, int& y) {     // squavathr.go:559-610
    for (;;) {
        if (printit) { std::printf("Your move: "); std::fflush(stdout); }
        int n = std::scanf("%d %d", &x, &y);
        if (n == EOF) std::exit(0);
        if (n != 2) { std::printf("Give two numbers\n"); int c; while ((c = std::getchar()) != '\n' && c != EOF) {} continue; }
        if (x < 0 || x > 4 || y < 0 || y > 4) { if (printit) std::printf("Choose two numbers between 0 and 4, try again\n"); continue; }
        if (bd.at(5 * x + y) == UNSET) return;
        if (printit) std::printf("Cell (%d, %d) already occupied, try again\n", x, y);
    }
}

// ---- -B: the opening book, squavathr.go:647-810 -------------------------------------------
enum BookState { FIRST, DIAGONAL, CORNER, OTHERCORNER, OTHERDIAGONAL, LAST };

static void bookReply(Board& bd, int cX, int cY, int& moveCount) {   // the "My move / Your move" step
    std::printf("My move: %d %d\n", cX, cY);
    std::printf("%s", bd.str_grid().c_str());
    int l, m;
    readMove(bd, true, l, m);
    bd.set(5 * l + m, MINIMIZER);
    moveCount++;
}

// Defend against the triangle: answer the human's first two moves three cells away on both
// axes (squavathr.go:662-733).  Returns the number of moves made by both sides.
static int bookDefend(Board& bd, int firstX, int firstY) {
    int moveCount = 0, cX = 0, cY = 0;
    // FIRST
    bool found = false;
    for (int i = -3; i < 4 && !found; i += 6) for (int j = -3; j < 4 && !found; j += 6) {
        int a = firstX + i, b = firstY + j;
        if (a >= 0 && a <= 4 && b >= 0 && b <= 4) { cX = a; cY = b; bd.set(5 * a + b, MAXIMIZER); moveCount++; found = true; }
    }
    bookReply(bd, cX, cY, moveCount);     // printed even when no cell was found (cX,cY = 0,0), as the reference does
    // DIAGONAL
    int lastx = 0, lasty = 0;
    for (int c = 0; c < 25; c++)
        if (!(c / 5 == firstX && c % 5 == firstY) && bd.at(c) == MINIMIZER) { lastx = c / 5; lasty = c % 5; break; }
    cX = -1; cY = -1;
    for (int i = -3; i < 4 && cX < 0; i += 6) for (int j = -3; j < 4 && cX < 0; j += 6) {
        int a = lastx + i, b = lasty + j;
        if (a >= 0 && a <= 4 && b >= 0 && b <= 4 && bd.at(5 * a + b) == UNSET) { bd.set(5 * a + b, MAXIMIZER); moveCount++; cX = a; cY = b; }
    }
    if (cX != -1 && cY != -1) {           // else: the human is not playing the triangle
        std::printf("My move: %d %d\n", cX, cY);
        std::printf("%s", bd.str_grid().c_str());
    }
    return moveCount;
}

// Play the triangle (squavathr.go:735-810).
static int bookStart(Board& bd) {
    static const int firstMoves[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
    int state = FIRST, moveCount = 0, cX = 0, cY = 0;
    auto empty = [&](int x, int y) { return bd.at(5 * x + y) == UNSET; };
    auto mark = [&]() { bd.set(5 * cX + cY, MAXIMIZER); moveCount++; };
    while (state != LAST) {
        switch (state) {
        case FIRST: {
            const char* pin = std::getenv("SQUAVA_BOOK_PICK");     // tests: stands in for rand.Intn(4), :746
            const int* c = firstMoves[pin && *pin ? std::atoi(pin) & 3 : std::uniform_int_distribution<int>(0, 3)(host_rng())];
            cX = c[0]; cY = c[1]; mark();
            state = DIAGONAL;
            break;
        }
        case DIAGONAL:
            if (empty(cX + 3, cY + 3)) { cX += 3; cY += 3; mark(); state = CORNER; }
            else state = OTHERCORNER;
            break;
        case CORNER:
            if (empty(cX - 3, cY)) { cX -= 3; mark(); }
            else if (empty(cX, cY - 3)) { cY -= 3; mark(); }
            state = LAST;
            break;
        case OTHERCORNER:
            if (empty(cX, cY + 3)) { cY += 3; mark(); state = OTHERDIAGONAL; }
            else {
                std::printf("Unreachable state in OTHERCORNER\n");
                std::printf("%s", bd.str_grid().c_str());
                std::exit(99);
            }
            break;
        case OTHERDIAGONAL:
            if (empty(cX + 3, cY - 3)) { cX += 3; cY -= 3; mark(); }
            state = LAST;
            break;
        }
        if (moveCount % 2 == 1) bookReply(bd, cX, cY, moveCount);
    }
    return moveCount;
}

// ---- -book FILE: play the first computer move from a table written by `opening3 -book` -----------
// Lines "x1 y1 x2 y2 value depth", values from the first mover's point of view.  Moving first, the
// computer takes the first move whose worst reply is best; answering a first move, the reply that is
// worst for the first mover.  Ties: the first in row-major order (or a random one without -D).
static bool bookTableMove(const std::string& file, const Board& bd, int moveCounter, bool deterministic, int& bx, int& by, int& bv) {
    FILE* fp = std::fopen(file.c_str(), "r");
    if (!fp) { std::fprintf(stderr, "open %s: no such file or directory\n", file.c_str()); std::exit(1); }
    static int table[25][25]; static bool have[25][25];
    int a, b, c, d, v, depth;
    while (std::fscanf(fp, "%d %d %d %d %d %d", &a, &b, &c, &d, &v, &depth) == 6)
        if (a >= 0 && a < 5 && b >= 0 && b < 5 && c >= 0 && c < 5 && d >= 0 && d < 5) { table[5 * a + b][5 * c + d] = v; have[5 * a + b][5 * c + d] = true; }
    std::fclose(fp);
    MoveKeeper moves(3 * LOSS, deterministic, &host_rng());
    if (moveCounter == 0) {
        for (int f = 0; f < 25; f++) {
            int worst = 3 * WIN; bool any = false;
            for (int r = 0; r < 25; r++) if (have[f][r]) { any = true; if (table[f][r] < worst) worst = table[f][r]; }
            if (any) moves.SetMove(f / 5, f % 5, worst);
        }
    } else if (moveCounter == 1) {
        int f = -1;
        for (int cc = 0; cc < 25; cc++) if (bd.at(cc) == MINIMIZER) f = cc;
        if (f < 0) return false;
        for (int r = 0; r < 25; r++) if (have[f][r]) moves.SetMove(r / 5, r % 5, -table[f][r]);
    } else {
        return false;
    }
    moves.ChooseMove(bx, by, bv);
    return bx >= 0;
}

int main(int argc, char** argv) {
    bool humanFirst = true, computerFirst = false, deterministic = false, noPrint = false, randomize = false, useBook = false;
    int maxDepthFlag = 15, threads = 0, gpus = 0, seed = -1;
    std::string firstMove, bookFile;
    Flags fl;
    fl.Bool("H", &humanFirst, "Human takes first move");
    fl.Bool("C", &computerFirst, "Computer takes first move");
    fl.Int("d", &maxDepthFlag, "maximum lookahead depth");
    fl.Bool("D", &deterministic, "Play deterministically");
    fl.Bool("n", &noPrint, "Don't print board, just emit moves");
    fl.String("M", &firstMove, "Tell computer to make this first move (x,y)");
    fl.Int("N", &threads, "Use this many threads (ignored: the root moves are one GPU batch)");
    fl.Bool("r", &randomize, "Randomize bias scores");
    fl.Bool("B", &useBook, "Use book start or defense");
    fl.String("book", &bookFile, "play the computer's first move from this table (written by opening3 -book)");
    fl.Int("gpus", &gpus, "number of GPUs (default 1)");
    fl.Int("seed", &seed, "seed for reproducible play");
    fl.Parse(argc, argv);
    if (seed >= 0) seed_host_rng((uint64_t)seed);
    bool printBoard = !noPrint;
    init_devices(gpus, 0);
    setScores(randomize);
    if (computerFirst) humanFirst = false;
    int moveCounter = 0;
    Board bd;
    if (useBook) {                                  // squavathr.go:93-106
        std::printf("Using opening book\n");
        firstMove.clear();
        if (humanFirst) {
            int l, m;
            readMove(bd, printBoard, l, m);
            bd.set(5 * l + m, MINIMIZER);
            moveCounter = 1 + bookDefend(bd, l, m);
            if (moveCounter % 2 == 1) humanFirst = false;
        } else {
            moveCounter += bookStart(bd);
        }
    }
    if (!firstMove.empty()) {
        int x1, y1;
        if (std::sscanf(firstMove.c_str(), "%d,%d", &x1, &y1) != 2)
            std::printf("Wrong move input. Format N,M, where N and M numbers 0 - 4 \n");
        else {
            std::printf("My move: %d %d\n", x1, y1);
            humanFirst = true;
            bd.set(5 * x1 + y1, MAXIMIZER);
            std::printf("%s", bd.str_grid().c_str());
        }
    }
    bool endOfGame = false;
    while (!endOfGame) {
        int maxDepth = setDepth(moveCounter, maxDepthFlag);
        if (humanFirst) {
            int l, m;
            readMove(bd, printBoard, l, m);
            bd.set(5 * l + m, MINIMIZER);
            endOfGame = game_over_after_move(bd);
            moveCounter++;
        }
        if (endOfGame) break;
        humanFirst = true;
        if (!bookFile.empty() && moveCounter < 2) {
            int bx, by, bv;
            if (bookTableMove(bookFile, bd, moveCounter, deterministic, bx, by, bv)) {
                bd.set(5 * bx + by, MAXIMIZER);
                moveCounter++;
                if (printBoard) { std::printf("My move: %d %d (%d) [book]\n", bx, by, bv); std::printf("%s", bd.str_grid().c_str()); }
                else std::printf("%d %d\n", bx, by);
                std::fflush(stdout);
                continue;
            }
        }
        auto start = std::chrono::steady_clock::now();
        // chooseMove(&bd, deterministic, maxDepth), squavathr.go:236-269
        sq_board b = bd.raw();
        int32_t values[1][25]; int32_t bv = 0; uint32_t bm = 0; uint64_t nodes = 0;
        must(sq_alphabeta_root(&b, 1, SQ_AB_SQUAVATHR, maxDepth, scores, values, &bv, &bm, &nodes), "sq_alphabeta_root");
        MoveKeeper moves(2 * LOSS, deterministic, &host_rng());
        for (int c = 0; c < 25; c++) if (values[0][c] != SQ_NO_MOVE) moves.SetMove(c / 5, c % 5, values[0][c]);
        int a, bb, score;
        moves.ChooseMove(a, bb, score);
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        if (a < 0) break;    // cat gets the game
        bd.set(5 * a + bb, MAXIMIZER);
        moveCounter++;
        if (printBoard) {
            std::printf("My move: %d %d (%d) [%llu] %s\n", a, bb, score, (unsigned long long)nodes, go_duration(el).c_str());
            std::printf("%s", bd.str_grid().c_str());
        } else {
            std::printf("%d %d\n", a, bb);
        }
        std::fflush(stdout);
        endOfGame = game_over_after_move(bd);
    }
    if (printBoard) {
        int w = findWinner(bd);
        std::printf("%s", w == MAXIMIZER ? "\nX wins\n" : w == MINIMIZER ? "\nO wins\n" : "\nCat wins\n");
        std::printf("%s", bd.str_grid().c_str());
    }
    sq_shutdown();
    return 0;
}
