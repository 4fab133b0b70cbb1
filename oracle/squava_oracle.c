/*
 * squava_oracle.c -- see squava_oracle.h.  TEST INFRASTRUCTURE, NOT PRODUCT.
 * PARITY PINNED (see the header): against output of the reference's squava2.py, of its Go
 * files executed through a mechanical in-memory translation, and of its squava.html likewise.
 *
 * Every function names the reference lines it restates.  Array board + table
 * walks on purpose (see header).
 */
#include "squava_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define WIN 10000
#define LOSS (-10000)
#define MAXIMIZER 1
#define MINIMIZER (-1)
#define UNSET 0

/* ------------------------------------------------------------------------ */
/* Tables                                                                    */
/* ------------------------------------------------------------------------ */

/* A line is written (start cell, step); its cells are start + k*step.  Steps:
 * +1 along a row (y+1), +5 down a column (x+1), +6 / +4 the two diagonals,
 * negative = the reference lists the cells in descending order.             */
typedef struct { signed char start, step; } line_t;

/* src/mcts/mcts.go:404 -- middle row and middle column */
static const int k_important[9] = {2, 7, 10, 11, 12, 13, 14, 17, 22};

/* src/mcts/mcts.go:409-425: quads grouped under the important cell they hold,
 * in the reference's order.  -1 start terminates a group.                   */
static const line_t k_mcts_quads[9][7] = {
    /* 2  */ {{0, 1}, {1, 1}, {2, 5}, {-1, 0}},
    /* 7  */ {{5, 1}, {6, 1}, {7, 5}, {-1, 0}},
    /* 10 */ {{0, 5}, {5, 5}, {-1, 0}},
    /* 11 */ {{1, 5}, {6, 5}, {3, 4}, {5, 6}, {-1, 0}},
    /* 12 */ {{10, 1}, {11, 1}, {4, 4}, {8, 4}, {0, 6}, {6, 6}, {-1, 0}},
    /* 13 */ {{3, 5}, {8, 5}, {1, 6}, {9, 4}, {-1, 0}},
    /* 14 */ {{4, 5}, {9, 5}, {-1, 0}},
    /* 17 */ {{15, 1}, {16, 1}, {-1, 0}},
    /* 22 */ {{20, 1}, {21, 1}, {-1, 0}},
};
/* src/mcts/mcts.go:430-446: 72 triplet entries (48 distinct lines) */
static const line_t k_mcts_trips[9][13] = {
    /* 2  */ {{0, 1}, {1, 1}, {2, 1}, {2, 5}, {2, 4}, {14, -6}, {-1, 0}},
    /* 7  */ {{5, 1}, {6, 1}, {7, 1}, {2, 5}, {7, 5}, {3, 4}, {7, 4}, {1, 6}, {7, 6}, {-1, 0}},
    /* 10 */ {{10, 1}, {0, 5}, {5, 5}, {10, 5}, {2, 4}, {10, 6}, {-1, 0}},
    /* 11 */ {{10, 1}, {11, 1}, {1, 5}, {6, 5}, {11, 5}, {3, 4}, {7, 4}, {5, 6}, {11, 6}, {-1, 0}},
    /* 12 */ {{10, 1}, {11, 1}, {12, 1}, {2, 5}, {7, 5}, {12, 5}, {0, 6}, {6, 6}, {12, 6},
              {4, 4}, {8, 4}, {12, 4}, {-1, 0}},
    /* 13 */ {{11, 1}, {12, 1}, {3, 5}, {8, 5}, {13, 5}, {1, 6}, {7, 6}, {21, -4}, {17, -4}, {-1, 0}},
    /* 14 */ {{12, 1}, {4, 5}, {9, 5}, {14, 5}, {22, -4}, {14, -6}, {-1, 0}},
    /* 17 */ {{15, 1}, {16, 1}, {17, 1}, {7, 5}, {12, 5}, {5, 6}, {11, 6}, {21, -4}, {17, -4}, {-1, 0}},
    /* 22 */ {{20, 1}, {21, 1}, {22, 1}, {12, 5}, {10, 6}, {22, -4}, {-1, 0}},
};

/* opening3.go:328-361 (identical opening2.go:272-305): child order of AB-3 */
static const int k_ordered[25] = {6, 8, 16, 18, 12, 7, 11, 13, 17, 1, 2, 3, 5,
                                  10, 15, 9, 14, 19, 21, 22, 23, 0, 4, 20, 24};
/* squavathr.go:305-311 / :297-302 (same lists alphabeta.go:457-472) */
static const int k_no2[4][3] = {{10, 6, 2}, {2, 8, 14}, {22, 18, 14}, {22, 16, 10}};
static const int k_no_middle2[4][4] = {{15, 11, 7, 3}, {5, 11, 17, 23}, {1, 7, 13, 19}, {9, 13, 17, 21}};

/* expanded forms, built once */
static int g_mq_n[25], g_mq[25][6][4];   /* mcts quads by important cell     */
static int g_mt_n[25], g_mt[25][12][3];  /* mcts triplets by important cell  */
static int g_aq[28][4], g_at[48][3];     /* alpha-beta flat lists            */
static int g_iq_n[25], g_iq[25][8];      /* indexedWinningQuads: indices into g_aq, table order */
static int g_it_n[25], g_it[25][12];     /* indexedLosingTriplets                              */
static int g_mat[8][25];                 /* opening3 Matrices, cell -> image cell */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static int in_board(int x, int y) { return x >= 0 && x < 5 && y >= 0 && y < 5; }

static void build_tables(void) {
    for (int k = 0; k < 9; k++) {
        int c = k_important[k];
        for (int j = 0; k_mcts_quads[k][j].start >= 0; j++) {
            for (int i = 0; i < 4; i++)
                g_mq[c][g_mq_n[c]][i] = k_mcts_quads[k][j].start + i * k_mcts_quads[k][j].step;
            g_mq_n[c]++;
        }
        for (int j = 0; k_mcts_trips[k][j].start >= 0; j++) {
            for (int i = 0; i < 3; i++)
                g_mt[c][g_mt_n[c]][i] = k_mcts_trips[k][j].start + i * k_mcts_trips[k][j].step;
            g_mt_n[c]++;
        }
    }
    /* alphabeta.go:299-328 lists the quads by first cell with x outermost, and
     * :249-298 lists the triplets by first cell with y outermost; from each first
     * cell the directions run (x+1), (y+1), (x-1,y+1), (x+1,y+1).            */
    static const int dx[4] = {1, 0, -1, 1}, dy[4] = {0, 1, 1, 1};
    int nq = 0, nt = 0;
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++)
            for (int d = 0; d < 4; d++)
                if (in_board(x + 3 * dx[d], y + 3 * dy[d])) {
                    for (int i = 0; i < 4; i++) g_aq[nq][i] = 5 * (x + i * dx[d]) + y + i * dy[d];
                    nq++;
                }
    for (int y = 0; y < 5; y++)
        for (int x = 0; x < 5; x++)
            for (int d = 0; d < 4; d++)
                if (in_board(x + 2 * dx[d], y + 2 * dy[d])) {
                    for (int i = 0; i < 3; i++) g_at[nt][i] = 5 * (x + i * dx[d]) + y + i * dy[d];
                    nt++;
                }
    /* calculateIndexedMatrices, alphabeta.go:37-50: table order preserved */
    for (int t = 0; t < 48; t++)
        for (int i = 0; i < 3; i++) { int c = g_at[t][i]; g_it[c][g_it_n[c]++] = t; }
    for (int q = 0; q < 28; q++)
        for (int i = 0; i < 4; i++) { int c = g_aq[q][i]; g_iq[c][g_iq_n[c]++] = q; }
    /* opening3.go:379-451: I, rot90 CCW, rot180, rot90 CW, flip rows, flip
     * columns, anti-transpose, transpose -- image of (x,y) under each.        */
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) {
            int c = 5 * x + y;
            g_mat[0][c] = 5 * x + y;
            g_mat[1][c] = 5 * y + (4 - x);
            g_mat[2][c] = 5 * (4 - x) + (4 - y);
            g_mat[3][c] = 5 * (4 - y) + x;
            g_mat[4][c] = 5 * (4 - x) + y;
            g_mat[5][c] = 5 * x + (4 - y);
            g_mat[6][c] = 5 * (4 - y) + (4 - x);
            g_mat[7][c] = 5 * y + x;
        }
}
static void tables(void) { pthread_once(&g_once, build_tables); }

int sqo_tables_mcts_important(int out[9]) { memcpy(out, k_important, sizeof k_important); return 9; }
int sqo_tables_mcts_quads(int cell, int out[][4]) {
    tables(); memcpy(out, g_mq[cell], sizeof(int) * 4 * g_mq_n[cell]); return g_mq_n[cell];
}
int sqo_tables_mcts_triplets(int cell, int out[][3]) {
    tables(); memcpy(out, g_mt[cell], sizeof(int) * 3 * g_mt_n[cell]); return g_mt_n[cell];
}
int sqo_tables_ab_quads(int out[28][4]) { tables(); memcpy(out, g_aq, sizeof g_aq); return 28; }
int sqo_tables_ab_triplets(int out[48][3]) { tables(); memcpy(out, g_at, sizeof g_at); return 48; }
int sqo_tables_ordered_cells(int out[25]) { memcpy(out, k_ordered, sizeof k_ordered); return 25; }
int sqo_tables_scores(int which, int out[25]) {
    /* alphabeta.go:343-349: 3 on the outer ring except its mid-edges (0), 4 on the
     * inner ring's corners, 1 on its mid-edges, 0 in the centre.  opening2.go:307-313
     * / opening3.go:366-372 carry the same table with [2][4] = 2.             */
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) {
            int ring = (x == 0 || x == 4 || y == 0 || y == 4);
            int mid = (x == 2 || y == 2);
            int v = ring ? (mid ? 0 : 3) : (x == 2 && y == 2) ? 0 : (mid ? 1 : 4);
            out[5 * x + y] = v;
        }
    if (which == 1) out[14] = 2;
    return 25;
}
int sqo_tables_no2(int out[4][3]) { memcpy(out, k_no2, sizeof k_no2); return 4; }
int sqo_tables_no_middle2(int out[4][4]) { memcpy(out, k_no_middle2, sizeof k_no_middle2); return 4; }
int sqo_tables_matrices(int out[8][25]) { tables(); memcpy(out, g_mat, sizeof g_mat); return 8; }
int sqo_tables_first_moves(int out[6]) {
    static const int fm[6] = {0, 1, 2, 6, 7, 12}; /* opening3.go:456-463 */
    memcpy(out, fm, sizeof fm); return 6;
}

static void unpack(sqo_board b, int bd[25]) {
    for (int c = 0; c < 25; c++)
        bd[c] = ((b.x >> c) & 1u) ? MAXIMIZER : ((b.o >> c) & 1u) ? MINIMIZER : UNSET;
}

/* ------------------------------------------------------------------------ */
/* MCTS-side board primitives                                                */
/* ------------------------------------------------------------------------ */

/* GetMoves, mcts.go:310-346.  Returns terminal flag; writes the move list.  */
static int get_moves(const int bd[25], int moves[25], int *n_moves) {
    *n_moves = 0;
    for (int k = 0; k < 9; k++) {
        int m = k_important[k];
        if (bd[m] != UNSET) {
            for (int j = 0; j < g_mq_n[m]; j++) {
                const int *q = g_mq[m][j];
                int sum = bd[q[0]] + bd[q[1]] + bd[q[2]] + bd[q[3]];
                if (sum == 4 || sum == -4) return 1;
            }
            for (int j = 0; j < g_mt_n[m]; j++) {
                const int *t = g_mt[m][j];
                int sum = bd[t[0]] + bd[t[1]] + bd[t[2]];
                if (sum == 3 || sum == -3) return 1;
            }
        }
    }
    int end_of_game = 1;
    for (int i = 0; i < 25; i++)
        if (bd[i] == UNSET) { end_of_game = 0; moves[(*n_moves)++] = i; }
    return end_of_game;
}

/* FindWinner, mcts.go:92-121: all quads in important-cell order, then triplets */
static int mcts_find_winner(const int bd[25]) {
    for (int k = 0; k < 9; k++) {
        int i = k_important[k];
        if (bd[i] != UNSET)
            for (int j = 0; j < g_mq_n[i]; j++) {
                const int *q = g_mq[i][j];
                int sum = bd[q[0]] + bd[q[1]] + bd[q[2]] + bd[q[3]];
                if (sum == 4) return MAXIMIZER;
                if (sum == -4) return MINIMIZER;
            }
    }
    for (int k = 0; k < 9; k++) {
        int i = k_important[k];
        if (bd[i] != UNSET)
            for (int j = 0; j < g_mt_n[i]; j++) {
                const int *t = g_mt[i][j];
                int sum = bd[t[0]] + bd[t[1]] + bd[t[2]];
                if (sum == 3) return MINIMIZER;
                if (sum == -3) return MAXIMIZER;
            }
    }
    return UNSET;
}

/* GetResult, mcts.go:348-388 (without the cache, which only memoises) */
static double mcts_get_result(const int bd[25], int playerjm) {
    for (int k = 0; k < 9; k++) {
        int i = k_important[k];
        if (bd[i] != UNSET)
            for (int j = 0; j < g_mq_n[i]; j++) {
                const int *q = g_mq[i][j];
                int sum = bd[q[0]] + bd[q[1]] + bd[q[2]] + bd[q[3]];
                if (sum == 4 || sum == -4) return sum == 4 * playerjm ? 1.0 : 0.0;
            }
    }
    for (int k = 0; k < 9; k++) {
        int i = k_important[k];
        if (bd[i] != UNSET)
            for (int j = 0; j < g_mt_n[i]; j++) {
                const int *t = g_mt[i][j];
                int sum = bd[t[0]] + bd[t[1]] + bd[t[2]];
                if (sum == 3 || sum == -3) return sum == 3 * playerjm ? 0.0 : 1.0;
            }
    }
    return 0.0;
}

/* FindWinner, alphabeta.go:355-378: flat lists, mark of the line's first cell */
static int ab_find_winner(const int bd[25]) {
    for (int q = 0; q < 28; q++) {
        int sum = bd[g_aq[q][0]] + bd[g_aq[q][1]] + bd[g_aq[q][2]] + bd[g_aq[q][3]];
        if (sum == 4 || sum == -4) return bd[g_aq[q][0]];
    }
    for (int t = 0; t < 48; t++) {
        int sum = bd[g_at[t][0]] + bd[g_at[t][1]] + bd[g_at[t][2]];
        if (sum == 3 || sum == -3) return -bd[g_at[t][0]];
    }
    return 0;
}

int sqo_mcts_terminal(sqo_board b) {
    int bd[25], mv[25], n; tables(); unpack(b, bd); return get_moves(bd, mv, &n);
}
int sqo_mcts_find_winner(sqo_board b) { int bd[25]; tables(); unpack(b, bd); return mcts_find_winner(bd); }
double sqo_mcts_get_result(sqo_board b, int pjm) { int bd[25]; tables(); unpack(b, bd); return mcts_get_result(bd, pjm); }
int sqo_ab_find_winner(sqo_board b) { int bd[25]; tables(); unpack(b, bd); return ab_find_winner(bd); }

static unsigned flags_of(const int bd[25]) {
    unsigned f = 0; int empty = 0;
    for (int q = 0; q < 28; q++) {
        int sum = bd[g_aq[q][0]] + bd[g_aq[q][1]] + bd[g_aq[q][2]] + bd[g_aq[q][3]];
        if (sum == 4) f |= SQO_X4;
        if (sum == -4) f |= SQO_O4;
    }
    for (int t = 0; t < 48; t++) {
        int sum = bd[g_at[t][0]] + bd[g_at[t][1]] + bd[g_at[t][2]];
        if (sum == 3) f |= SQO_X3;
        if (sum == -3) f |= SQO_O3;
    }
    for (int c = 0; c < 25; c++) empty += bd[c] == UNSET;
    if (!empty) f |= SQO_FULL;
    return f;
}
unsigned sqo_flags(sqo_board b) { int bd[25]; tables(); unpack(b, bd); return flags_of(bd); }

void sqo_classify(const sqo_board *b, int64_t n, uint8_t *flags, int8_t *wm, int8_t *wa) {
    tables();
    for (int64_t i = 0; i < n; i++) {
        int bd[25]; unpack(b[i], bd);
        if (flags) flags[i] = (uint8_t)flags_of(bd);
        if (wm) wm[i] = (int8_t)mcts_find_winner(bd);
        if (wa) wa[i] = (int8_t)ab_find_winner(bd);
    }
}

/* Enumeration of every board with k marks, X-count = ceil(k/2): rank r in
 * [0, C(25,k)*C(k,nx)) -> (which k cells, which nx of them are X), combinatorial number system. */
static uint64_t binom_tab[26][26];
static void binom_init(void) {
    if (binom_tab[0][0]) return;
    for (int n = 0; n < 26; n++) {
        binom_tab[n][0] = 1;
        for (int r = 1; r <= n; r++) binom_tab[n][r] = binom_tab[n - 1][r - 1] + (r <= n - 1 ? binom_tab[n - 1][r] : 0);
    }
}
static void unrank_comb(uint64_t r, int n, int k, int *out) {   /* k-subset of 0..n-1, ascending */
    int c = 0;
    for (int i = 0; i < k; i++) {
        for (;; c++) {
            uint64_t with = binom_tab[n - c - 1][k - i - 1];   /* subsets that take c next */
            if (r < with) { out[i] = c++; break; }
            r -= with;
        }
    }
}
int64_t sqo_count_alternating(int k) {
    binom_init();
    if (k < 0 || k > 25) return 0;
    return (int64_t)(binom_tab[25][k] * binom_tab[k][(k + 1) / 2]);
}
/* next k-combination of 0..n-1 in lexicographic order (the order unrank_comb counts in); 0 when c was the last */
static int next_comb(int *c, int n, int k) {
    int i = k - 1;
    while (i >= 0 && c[i] == n - k + i) i--;
    if (i < 0) return 0;
    c[i]++;
    for (int j = i + 1; j < k; j++) c[j] = c[j - 1] + 1;
    return 1;
}
void sqo_enumerate_alternating(int k, int64_t start, int64_t count, sqo_board *out) {
    binom_init();
    if (count <= 0) return;
    int nx = (k + 1) / 2;
    uint64_t per = binom_tab[k][nx];
    int cells[25], xs[25];
    /* unrank the first board of the range, then step: the X-subset runs fastest */
    unrank_comb((uint64_t)start / per, 25, k, cells);
    unrank_comb((uint64_t)start % per, k, nx, xs);
    for (int64_t i = 0; i < count; i++) {
        uint32_t all = 0, x = 0;
        for (int j = 0; j < k; j++) all |= 1u << cells[j];
        for (int j = 0; j < nx; j++) x |= 1u << cells[xs[j]];
        out[i].x = x; out[i].o = all & ~x;
        if (!next_comb(xs, k, nx)) {
            for (int j = 0; j < nx; j++) xs[j] = j;
            if (!next_comb(cells, 25, k)) break;         /* that was the last board with k marks */
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Random playouts                                                           */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed); }
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t *r) { /* xoshiro256** */
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
/* unbiased draw in [0,n), n <= 25: what rand.Intn promises (Go rejects the biased tail) */
static int rng_intn(rng_t *r, int n) {
    uint32_t un = (uint32_t)n, lim = (uint32_t)(-un) % un; /* 2^32 mod n */
    for (;;) {
        uint32_t v = (uint32_t)(rng_next(r) >> 32);
        uint64_t m = (uint64_t)v * un;
        if ((uint32_t)m >= lim) return (int)(m >> 32);
    }
}

/* One playout: the rollout loop mcts.go:173-181, then the result for both
 * players (mcts.go:190-192 calls GetResult per path node's playerJustMoved). */
static void one_playout(const int leaf[25], int player_just_moved, rng_t *r, uint64_t out[4]) {
    int bd[25], moves[25], n;
    memcpy(bd, leaf, sizeof bd);                      /* Clone, mcts.go:279-284 */
    int pjm = player_just_moved;
    int term = get_moves(bd, moves, &n);
    while (!term) {
        int m = moves[rng_intn(r, n)];
        pjm = -pjm; bd[m] = pjm;                       /* DoMove, mcts.go:292-295 */
        out[3]++;
        term = get_moves(bd, moves, &n);
    }
    if (mcts_get_result(bd, MAXIMIZER) == 1.0) out[0]++;
    else if (mcts_get_result(bd, MINIMIZER) == 1.0) out[1]++;
    else out[2]++;
}

void sqo_playouts(sqo_board b, int to_move, uint64_t n_playouts, uint64_t seed, uint64_t out[4]) {
    int bd[25]; rng_t r;
    tables(); unpack(b, bd); rng_seed(&r, seed);
    out[0] = out[1] = out[2] = out[3] = 0;
    for (uint64_t i = 0; i < n_playouts; i++) one_playout(bd, -to_move, &r, out);
}

typedef struct {
    const sqo_board *b; const int8_t *to_move; int64_t n_leaves; uint64_t per_leaf, seed;
    int tid, n_threads; uint64_t (*out)[4];
} batch_arg;
static void *batch_worker(void *p) {
    batch_arg *a = (batch_arg *)p;
    for (int64_t i = a->tid; i < a->n_leaves; i += a->n_threads)
        sqo_playouts(a->b[i], a->to_move[i], a->per_leaf, a->seed + 0x632be59bd9b4e019ull * (uint64_t)(i + 1), a->out[i]);
    return NULL;
}
void sqo_playouts_batch(const sqo_board *b, const int8_t *to_move, int64_t n_leaves,
                        uint64_t per_leaf, uint64_t seed, int n_threads, uint64_t (*out)[4]) {
    tables();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; batch_arg args[256];
    for (int t = 0; t < n_threads; t++) {
        args[t] = (batch_arg){b, to_move, n_leaves, per_leaf, seed, t, n_threads, out};
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}

/* ---- exact playout distribution: forward DP, one hash map per ply -------- */
typedef struct { uint64_t *key; double *val; uint64_t cap, n; } pmap;
static int pmap_init(pmap *m, uint64_t cap) {
    m->cap = cap; m->n = 0;
    m->key = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->val = (double *)malloc(cap * sizeof(double));
    if (!m->key || !m->val) return -1;
    memset(m->key, 0xff, cap * sizeof(uint64_t));
    return 0;
}
static void pmap_free(pmap *m) { free(m->key); free(m->val); m->key = NULL; m->val = NULL; }
static int pmap_add(pmap *m, uint64_t k, double v);
static int pmap_grow(pmap *m) {
    pmap big;
    if (pmap_init(&big, m->cap * 2)) return -1;
    for (uint64_t i = 0; i < m->cap; i++)
        if (m->key[i] != ~0ull) pmap_add(&big, m->key[i], m->val[i]);
    pmap_free(m); *m = big; return 0;
}
static int pmap_add(pmap *m, uint64_t k, double v) {
    if (2 * (m->n + 1) > m->cap && pmap_grow(m)) return -1;
    uint64_t h = (k * 0x9e3779b97f4a7c15ull) >> 20, mask = m->cap - 1;
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
        if (m->key[i] == k) { m->val[i] += v; return 0; }
        if (m->key[i] == ~0ull) { m->key[i] = k; m->val[i] = v; m->n++; return 0; }
    }
}

int64_t sqo_playout_exact(sqo_board b, int to_move, double p[3], double *mean_plies) {
    tables();
    p[0] = p[1] = p[2] = 0.0;
    double plies = 0.0; int64_t visited = 0;
    pmap cur, nxt;
    if (pmap_init(&cur, 1u << 10)) return -1;
    pmap_add(&cur, ((uint64_t)b.o << 25) | b.x, 1.0);
    int pjm = -to_move;
    for (int depth = 0; cur.n > 0; depth++) {
        if (pmap_init(&nxt, cur.cap)) { pmap_free(&cur); return -1; }
        for (uint64_t i = 0; i < cur.cap; i++) {
            if (cur.key[i] == ~0ull) continue;
            visited++;
            sqo_board s = {(uint32_t)(cur.key[i] & 0x1ffffffu), (uint32_t)(cur.key[i] >> 25)};
            double pr = cur.val[i];
            int bd[25], moves[25], n;
            unpack(s, bd);
            if (get_moves(bd, moves, &n)) { /* terminal: nobody moves again */
                if (mcts_get_result(bd, MAXIMIZER) == 1.0) p[0] += pr;
                else if (mcts_get_result(bd, MINIMIZER) == 1.0) p[1] += pr;
                else p[2] += pr;
                plies += pr * depth;
                continue;
            }
            for (int j = 0; j < n; j++) {
                sqo_board c = s;
                if (-pjm == MAXIMIZER) c.x |= 1u << moves[j]; else c.o |= 1u << moves[j];
                if (pmap_add(&nxt, ((uint64_t)c.o << 25) | c.x, pr / n)) { pmap_free(&cur); pmap_free(&nxt); return -1; }
            }
        }
        pmap_free(&cur); cur = nxt; pjm = -pjm;
    }
    pmap_free(&cur);
    if (mean_plies) *mean_plies = plies;
    return visited;
}

/* ------------------------------------------------------------------------ */
/* Alpha-beta dialects                                                       */
/* ------------------------------------------------------------------------ */
typedef struct {
    int bd[25];
    int max_depth;
    const int *scores;
    int64_t leaves;
    int sentinel; /* AB-3: value of a node with no empty cell is -+sentinel */
    int avoid;    /* AB-1 with SetAvoid(): boardValue = deltaValue2, alphabeta.go:382-454,474 */
    int cumulative; /* src/abbook: the recursion passes boardValue+delta down, abbook.go:228-229,255-256 */
} ab_t;

/* evaluator calls (deltaValue / the inlined evaluation of opening2-3) of the calling thread's
 * most recent search: the unit SURVEY.md 8(d) counts alpha-beta throughput in */
static __thread int64_t g_evals;
int64_t sqo_last_evals(void) { return g_evals; }

/* lines-through-(x,y) part shared by every dialect's evaluator:
 * alphabeta.go:123-147 == squavathr.go:317-341 == opening3.go:138-164.
 * Returns 1 and *value = the terminal score if (x,y) completes a line.      */
static int lines_value(const ab_t *a, int ply, int c, int *value) {
    const int *bd = a->bd; int v = 0;
    g_evals++;
    for (int k = 0; k < g_iq_n[c]; k++) {
        const int *q = g_aq[g_iq[c][k]];
        int sum = bd[q[0]] + bd[q[1]] + bd[q[2]] + bd[q[3]];
        if (sum == 4 || sum == -4) { *value = bd[q[0]] * (WIN - ply); return 1; }
        if (sum == 3 || sum == -3) v += sum * 10;
    }
    for (int k = 0; k < g_it_n[c]; k++) {
        const int *t = g_at[g_it[c][k]];
        int sum = bd[t[0]] + bd[t[1]] + bd[t[2]];
        if (sum == 3 || sum == -3) { *value = sum / 3 * (LOSS + ply); return 1; }
    }
    *value = v; return 0;
}

/* the no2 / noMiddle2 terms: squavathr.go:343-371 == alphabeta.go:412-439 (deltaValue2) */
static int avoid_terms(const ab_t *a, int c) {
    const int *bd = a->bd; int v = 0;
    for (int t = 0; t < 4; t++)
        for (int i = 0; i < 3; i++)
            if (k_no2[t][i] == c) {
                int sum = bd[k_no2[t][0]] + bd[k_no2[t][1]] + bd[k_no2[t][2]];
                if (sum == 2 || sum == -2) v += bd[c] * -100;
                break;
            }
    for (int q = 0; q < 4; q++) {
        const int *m = k_no_middle2[q]; int player = bd[c];
        if ((c == m[1] && player == bd[m[2]]) || (c == m[2] && player == bd[m[1]])) {
            int sum = bd[m[0]] + bd[m[1]] + bd[m[2]] + bd[m[3]];
            if (sum == 2 || sum == -2) v += player * -100;
        }
    }
    return v;
}

/* ---- AB-1: src/alphabeta ---- */
/* deltaValue, alphabeta.go:121-163; with a->avoid: deltaValue2, alphabeta.go:382-454 */
static int delta1(const ab_t *a, int ply, int c, int current, int *value) {
    int v;
    if (lines_value(a, ply, c, &v)) { *value = v; return 1; }
    if (a->avoid) v += avoid_terms(a, c);
    v += a->bd[c] * a->scores[c];
    int stop = 0;
    if (ply >= a->max_depth) { stop = 1; v += current; }
    *value = v; return stop;
}
/* alphaBeta, alphabeta.go:165-223: the child's mark is placed BEFORE the parent's
 * move (c) is evaluated; a stop returns from the first empty cell.           */
static int ab1(ab_t *a, int ply, int player, int alpha, int beta, int c, int board_value) {
    int value = player == MAXIMIZER ? 2 * LOSS : 2 * WIN;
    for (int i = 0; i < 25; i++) {
        if (a->bd[i] != UNSET) continue;
        a->bd[i] = player;
        int delta, stop = delta1(a, ply, c, board_value, &delta);
        if (stop) { a->bd[i] = UNSET; a->leaves++; return delta; }
        int n = ab1(a, ply + 1, -player, alpha, beta, i, a->cumulative ? board_value + delta : delta);
        a->bd[i] = UNSET;
        if (player == MAXIMIZER) {
            if (n > value) value = n;
            if (value > alpha) alpha = value;
        } else {
            if (n < value) value = n;
            if (value < beta) beta = value;
        }
        if (beta <= alpha) return value;
    }
    return value;
}

/* movekeeper.go:26-36 folded over a value row */
static void keep_moves(const int32_t values[25], uint32_t *best_mask, int32_t *best_value) {
    int max = 2 * LOSS; uint32_t mask = 0;
    for (int c = 0; c < 25; c++) {
        if (values[c] == SQO_NO_MOVE) continue;
        if (values[c] >= max) {
            if (values[c] > max) { max = values[c]; mask = 0; }
            mask |= 1u << c;
        }
    }
    if (best_mask) *best_mask = mask;
    if (best_value) *best_value = mask ? max : 0; /* movekeeper.go:40-44 */
}

static void ab1_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                     uint32_t *best_mask, int32_t *best_value, int64_t *leaves, int avoid);
void sqo_ab1_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                  uint32_t *best_mask, int32_t *best_value, int64_t *leaves) {
    ab1_root(b, max_depth, scores, values, best_mask, best_value, leaves, 0);
}
void sqo_ab5_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                  uint32_t *best_mask, int32_t *best_value, int64_t *leaves) {
    ab1_root(b, max_depth, scores, values, best_mask, best_value, leaves, 1);
}
void sqo_ab6_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                  uint32_t *best_mask, int32_t *best_value, int64_t *leaves) {
    ab1_root(b, max_depth, scores, values, best_mask, best_value, leaves, 2);
}
static void ab1_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                     uint32_t *best_mask, int32_t *best_value, int64_t *leaves, int avoid) {
    ab_t a; tables(); unpack(b, a.bd); g_evals = 0;
    a.max_depth = max_depth; a.scores = scores; a.leaves = 0; a.sentinel = 2 * WIN;
    a.avoid = avoid == 1; a.cumulative = avoid == 2;
    for (int c = 0; c < 25; c++) {            /* ChooseMove, alphabeta.go:98-110 */
        values[c] = SQO_NO_MOVE;
        if (a.bd[c] != UNSET) continue;
        a.bd[c] = MAXIMIZER;
        int value, stop = delta1(&a, 0, c, 0, &value);
        if (!stop) value = ab1(&a, 1, MINIMIZER, 2 * LOSS, 2 * WIN, c, value);
        a.bd[c] = UNSET;
        values[c] = value;
    }
    keep_moves(values, best_mask, best_value);
    if (leaves) *leaves = a.leaves;
}

/* ---- AB-2: squavathr ---- */
/* deltaValue, squavathr.go:315-387 */
static int delta2(const ab_t *a, int ply, int c, int current, int *value) {
    const int *bd = a->bd; int v;
    if (lines_value(a, ply, c, &v)) { *value = v; return 1; }
    v += avoid_terms(a, c);                /* :343-371 */
    v += bd[c] * a->scores[c];
    int stop = 0;
    if (ply == a->max_depth) { stop = 1; v += current; }
    *value = v; return stop;
}
/* alphaBeta, squavathr.go:389-446: evaluate at entry, cumulative value */
static int ab2(ab_t *a, int ply, int player, int alpha, int beta, int c, int board_value) {
    int delta;
    if (delta2(a, ply, c, board_value, &delta)) { a->leaves++; return delta; }
    board_value += delta;
    int value = player == MAXIMIZER ? 2 * LOSS : 2 * WIN;
    for (int i = 0; i < 25; i++) {
        if (a->bd[i] != UNSET) continue;
        a->bd[i] = player;
        int n = ab2(a, ply + 1, -player, alpha, beta, i, board_value);
        a->bd[i] = UNSET;
        if (player == MAXIMIZER) {
            if (n > value) value = n;
            if (value > alpha) alpha = value;
        } else {
            if (n < value) value = n;
            if (value < beta) beta = value;
        }
        if (beta <= alpha) { a->leaves++; return value; }
    }
    return value;
}

void sqo_ab2_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                  uint32_t *best_mask, int32_t *best_value, int64_t *leaves) {
    ab_t a; tables(); unpack(b, a.bd); g_evals = 0;
    a.max_depth = max_depth - 1;              /* squavathr.go:239 */
    a.scores = scores; a.leaves = 0; a.sentinel = 2 * WIN; a.avoid = 0; a.cumulative = 0;
    for (int c = 0; c < 25; c++) {            /* chooseMove, squavathr.go:241-253 */
        values[c] = SQO_NO_MOVE;
        if (a.bd[c] != UNSET) continue;
        int val, stop = delta2(&a, 0, c, 0, &val);   /* on the UNMARKED board */
        if (stop) { a.leaves++; values[c] = val; continue; }
        a.bd[c] = MAXIMIZER;                      /* newState, :230 */
        values[c] = ab2(&a, 1, MINIMIZER, 2 * LOSS, 2 * WIN, c, val);  /* worker, :451-460 */
        a.bd[c] = UNSET;
    }
    keep_moves(values, best_mask, best_value);
    if (leaves) *leaves = a.leaves;
}

/* ---- AB-3: opening2 / opening3 ---- */
/* alphaBeta, opening3.go:136-225 (opening2.go:91-180 differs only in the value
 * a node with no legal move returns: -+WIN there, -+2*WIN here).              */
static int ab3(ab_t *a, int ply, int next_player, int alpha, int beta, int c, int board_value) {
    int delta;
    if (lines_value(a, ply, c, &delta)) {
        a->leaves++;
        /* opening3.go:159 writes -sum/3*(WIN-ply); same number as lines_value's */
        return delta;
    }
    delta += a->bd[c] * a->scores[c];
    board_value += delta;
    if (ply == a->max_depth) { a->leaves++; return board_value; }
    int value = next_player == MAXIMIZER ? -a->sentinel : a->sentinel;
    for (int k = 0; k < 25; k++) {
        int i = k_ordered[k];
        if (a->bd[i] != UNSET) continue;
        a->bd[i] = next_player;
        int n = ab3(a, ply + 1, -next_player, alpha, beta, i, board_value);
        a->bd[i] = UNSET;
        if (next_player == MAXIMIZER) {
            if (n > value) value = n;
            if (value > alpha) alpha = value;
        } else {
            if (n < value) value = n;
            if (value < beta) beta = value;
        }
        if (beta <= alpha) break;
    }
    a->leaves++;
    return value;
}

int32_t sqo_ab_subtree(int dialect, sqo_board b, int last_cell, int ply, int side,
                       int32_t carried, int max_depth, const int scores[25], int64_t *leaves) {
    ab_t a; tables(); unpack(b, a.bd); g_evals = 0;
    a.max_depth = max_depth; a.scores = scores; a.leaves = 0;
    a.sentinel = dialect == 4 ? WIN : 2 * WIN;
    a.avoid = dialect == 5; a.cumulative = dialect == 6;
    int v = 0;
    switch (dialect) {
    case 1: case 5: case 6: v = ab1(&a, ply, side, 2 * LOSS, 2 * WIN, last_cell, carried); break;
    case 2: v = ab2(&a, ply, side, 2 * LOSS, 2 * WIN, last_cell, carried); break;
    case 3: case 4: v = ab3(&a, ply, side, LOSS, WIN, last_cell, carried); break;
    }
    if (leaves) *leaves = a.leaves;
    return v;
}

/* ------------------------------------------------------------------------ */
/* NegaScout, src/negascout/negascout.go (playoff5's "N" player)             */
/* ------------------------------------------------------------------------ */
static const int k_ns_initial_order[25] = {6, 8, 18, 16, 1, 3, 9, 19, 23, 21, 15, 5, 0, 4, 24, 20,
                                           12, 7, 13, 17, 11, 2, 10, 14, 22};   /* negascout.go:392-399 */
static const int k_ns_deadly[4][4] = {{5, 11, 17, 23}, {21, 17, 13, 9}, {1, 7, 13, 19}, {15, 11, 7, 3}}; /* :147-152 */
typedef struct {
    int bd[25];
    int max_depth;
    const int *scores;
    int64_t leaves;
    int order[3][25], n_order;   /* orderedMoves[0] for MINIMIZER, [2] for MAXIMIZER (:403) */
} ns_t;

/* reorderMoves, negascout.go:411-499: good / dull / bad empty cells */
static void ns_reorder(ns_t *s) {
    int good[25], bad[25], dull[25], ng = 0, nb = 0, nd = 0;
    for (int k = 0; k < 25; k++) {
        int c = k_ns_initial_order[k];
        if (s->bd[c] != UNSET) continue;
        int interesting = 0;
        for (int q = 0; q < g_iq_n[c] && !interesting; q++) {
            const int *l = g_aq[g_iq[c][q]];
            int sum = s->bd[l[0]] + s->bd[l[1]] + s->bd[l[2]] + s->bd[l[3]];
            if (sum == 3 || sum == 2) { good[ng++] = c; interesting = 1; }
            else if (sum == -3 || sum == -2) { bad[nb++] = c; interesting = 1; }
        }
        for (int t = 0; t < g_it_n[c] && !interesting; t++) {
            const int *l = g_at[g_it[c][t]];
            int sum = s->bd[l[0]] + s->bd[l[1]] + s->bd[l[2]];
            if (sum == -2) { good[ng++] = c; interesting = 1; }
            else if (sum == 2) { bad[nb++] = c; interesting = 1; }
        }
        if (!interesting) dull[nd++] = c;
    }
    int n = 0;
    for (int i = 0; i < ng; i++) s->order[2][n++] = good[i];      /* MAXIMIZER: good, dull, bad */
    for (int i = 0; i < nd; i++) s->order[2][n++] = dull[i];
    for (int i = 0; i < nb; i++) s->order[2][n++] = bad[i];
    n = 0;
    for (int i = 0; i < nb; i++) s->order[0][n++] = bad[i];       /* MINIMIZER: bad, dull, good */
    for (int i = 0; i < nd; i++) s->order[0][n++] = dull[i];
    for (int i = 0; i < ng; i++) s->order[0][n++] = good[i];
    s->n_order = n;
}

/* staticValue, negascout.go:168-230: whole-board evaluation; quads and triplets are visited
 * once per "checkable" cell they contain, in that order. */
static int ns_static(ns_t *s, int ply, int *value) {
    int v = 0;
    s->leaves++;
    g_evals++;
    for (int k = 0; k < 9; k++) {
        int c = k_important[k];                                    /* checkableCells, :159-163 */
        for (int q = 0; q < g_iq_n[c]; q++) {
            const int *l = g_aq[g_iq[c][q]];
            int sum = s->bd[l[0]] + s->bd[l[1]] + s->bd[l[2]] + s->bd[l[3]];
            if (sum == 4 || sum == -4) { *value = s->bd[l[0]] * (WIN - ply); return 1; }
            if (sum == 3 || sum == -3) v += sum * 10;
        }
    }
    for (int k = 0; k < 9; k++) {
        int c = k_important[k];
        for (int t = 0; t < g_it_n[c]; t++) {
            const int *l = g_at[g_it[c][t]];
            int sum = s->bd[l[0]] + s->bd[l[1]] + s->bd[l[2]];
            if (sum == 3 || sum == -3) { *value = -sum / 3 * (WIN - ply); return 1; }
        }
    }
    for (int d = 0; d < 4; d++) {
        const int *l = k_ns_deadly[d];
        int outer = s->bd[l[0]] + s->bd[l[3]], inner = s->bd[l[1]] + s->bd[l[2]];
        if ((inner == 2 || inner == -2) && outer == 0) v -= inner / 2 * 5;
    }
    if (v == 0)
        for (int c = 0; c < 25; c++) v += s->bd[c] * s->scores[c];
    *value = v;
    return ply > s->max_depth;
}

/* negaScout, negascout.go:232-266 */
static int ns_search(ns_t *s, int ply, int player, int alpha, int beta) {
    int bv;
    if (ns_static(s, ply, &bv)) return player * bv;
    int score = 3 * LOSS, n = beta;
    const int *order = s->order[player + 1];
    for (int k = 0; k < s->n_order; k++) {
        int c = order[k];
        if (s->bd[c] != UNSET) continue;
        s->bd[c] = player;
        int cur = -ns_search(s, ply + 1, -player, -n, -alpha);
        if (cur > score) {
            if (n == beta || ply == s->max_depth - 2) score = cur;
            else score = -ns_search(s, ply + 1, -player, -beta, -cur);
        }
        s->bd[c] = UNSET;
        if (score > alpha) alpha = score;
        if (alpha >= beta) break;
        n = alpha + 1;
    }
    return score;
}

/* ChooseMove, negascout.go:81-122, with movekeeper.New(3*LOSS, ...) (movekeeper.go:26-52).
 * values[c] = what SetMove was given for cell c (the running alpha), SQO_NO_MOVE for cells not
 * visited; visit_order[k] = k-th visited cell (orderedMoves[2] order; -1 padded); *best_mask = the
 * cells movekeeper holds at the end, *first_best = the one it returns when deterministic (-1 if none). */
void sqo_negascout_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                        int32_t visit_order[25], uint32_t *best_mask, int32_t *best_value,
                        int32_t *first_best, int64_t *leaves) {
    ns_t s; tables(); unpack(b, s.bd); g_evals = 0;
    s.max_depth = max_depth; s.scores = scores; s.leaves = 0;
    ns_reorder(&s);
    for (int c = 0; c < 25; c++) { values[c] = SQO_NO_MOVE; visit_order[c] = -1; }
    int beta = 2 * WIN, alpha = 2 * LOSS, score = 3 * LOSS, n = beta;
    int max = 3 * LOSS, first = -1, nv = 0; uint32_t mask = 0;
    for (int k = 0; k < s.n_order; k++) {
        int c = s.order[2][k];
        if (s.bd[c] != UNSET) continue;
        s.bd[c] = MAXIMIZER;
        int cur = -ns_search(&s, 1, MINIMIZER, -n, -alpha);
        if (cur > score) {
            if (n == beta) score = cur;
            else score = -ns_search(&s, 1, MINIMIZER, -beta, -cur);
        }
        s.bd[c] = UNSET;
        if (score > alpha) alpha = score;
        values[c] = alpha; visit_order[nv++] = c;
        if (alpha >= max) {                                       /* movekeeper.SetMove */
            if (alpha > max) { max = alpha; mask = 0; first = c; }
            mask |= 1u << c;
        }
        if (alpha >= beta) break;
        n = alpha + 1;
    }
    if (best_mask) *best_mask = mask;
    if (best_value) *best_value = mask ? max : 0;
    if (first_best) *first_best = first;
    if (leaves) *leaves = s.leaves;
}

/* opening3.go:66-128: a (first, reply) pair is skipped when some symmetry maps it
 * onto a pair already searched under ANY first move.                         */
int sqo_opening3_subtrees(int first[150], int reply[150]) {
    tables();
    int computed[6][25]; memset(computed, 0, sizeof computed);
    int fm[6]; sqo_tables_first_moves(fm);
    for (int k = 0; k < 6; k++) computed[k][fm[k]] = MAXIMIZER;
    int n = 0;
    for (int k = 0; k < 6; k++) {
        int f = fm[k];
        for (int r = 0; r < 25; r++) {
            if (r == f) continue;
            int matched = 0;
            for (int m = 0; m < 8 && !matched; m++) {
                int a = g_mat[m][f], c = g_mat[m][r];
                for (int j = 0; j < 6; j++)
                    if (computed[j][a] == MAXIMIZER && computed[j][c] == MINIMIZER) { matched = 1; break; }
            }
            if (!matched) { computed[k][r] = MINIMIZER; first[n] = f; reply[n] = r; n++; }
        }
    }
    return n;
}
