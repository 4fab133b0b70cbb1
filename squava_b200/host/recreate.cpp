// recreate -- replays a recorded game move by move (recreate.go:10-96): the file named on the
// command line holds moves "n,m" one per line, or all on one line separated by blanks (a
// playoff5 -n line or a playoff4 record works: fields that are not moves are reported on stderr
// and skipped, value markers "+" / "-" after a digit are dropped).  Enter steps to the next move.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static std::vector<std::string> split(const std::string& s, char sep) {     // bytes.Split
    std::vector<std::string> out;
    size_t at = 0;
    for (;;) {
        size_t k = s.find(sep, at);
        if (k == std::string::npos) { out.push_back(s.substr(at)); return out; }
        out.push_back(s.substr(at, k - at));
        at = k + 1;
    }
}

static bool finagleMove(const std::string& move, int moveCounter, int& n, int& m) {   // recreate.go:65-96
    std::vector<std::string> f = split(move, ',');
    if (f.size() != 2) { std::fprintf(stderr, "Move %d, \"%s\", problem\n", moveCounter, move.c_str()); return false; }
    if (f[0].size() == 2) f[0] = f[0].substr(0, 1);
    if (f[1].size() == 2) f[1] = f[1].substr(0, 1);
    if (f[0].empty() || f[1].empty()) { std::fprintf(stderr, "panic: runtime error: index out of range [0] with length 0\n"); std::exit(2); }
    if (f[0][0] < '0' || f[0][0] > '4') { std::fprintf(stderr, "Move %d, \"%s\", problem with 1st field\n", moveCounter, move.c_str()); return false; }
    n = f[0][0] - '0';
    if (f[1][0] < '0' || f[1][0] > '4') { std::fprintf(stderr, "Move %d, \"%s\", problem with 2nd field\n", moveCounter, move.c_str()); return false; }
    m = f[1][0] - '0';
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "panic: runtime error: index out of range [1] with length 1\n"); return 2; }
    const char markers[2] = {'X', 'O'};
    char board[5][5];
    for (auto& row : board) for (char& c : row) c = '_';
    FILE* fp = std::fopen(argv[1], "rb");
    if (!fp) { std::fprintf(stderr, "open %s: no such file or directory\n", argv[1]); return 1; }     // log.Fatal
    std::string buf;
    for (int c; (c = std::fgetc(fp)) != EOF;) buf.push_back((char)c);
    std::fclose(fp);
    int lines = 0;
    for (char c : buf) lines += c == '\n';
    std::vector<std::string> moves = split(buf, lines > 1 ? '\n' : ' ');
    int moveCounter = 0;
    for (const std::string& move : moves) {
        moveCounter++;
        if (moveCounter > 25) break;
        int n, m;
        if (!finagleMove(move, moveCounter, n, m)) continue;
        std::printf("%c move %d,%d\n", markers[moveCounter % 2], n, m);
        board[n][m] = markers[moveCounter % 2];
        for (auto& row : board) { for (char c : row) std::printf("%c ", c); std::printf("\n"); }
        std::printf("\n");
        std::fflush(stdout);
        int c = std::getchar();                       // fmt.Scanf("\n")
        if (c == EOF) std::fprintf(stderr, "EOF\n");
        else if (c != '\n') { std::fprintf(stderr, "unexpected newline\n"); while (c != '\n' && c != EOF) c = std::getchar(); }
    }
    return 0;
}
