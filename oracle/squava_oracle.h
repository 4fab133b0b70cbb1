/*
 * squava_oracle.h -- CPU restatement of the squava reference's hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped path (squava_b200/ + libsquava_b200.so) never calls into it.
 *
 * PARITY: PINNED against output of the reference's own code, obtained three ways (the
 * reference has no tests or golden vectors of its own, and this image has no Go toolchain,
 * no JavaScript engine and no Python 2):
 *   1. squava2.py, the Python prototype in the reference tree that src/mcts and squavam.go
 *      were transliterated from, loaded from /root/reference and driven directly
 *      (oracle/tools/make_ref_golden.py): GetMoves / GetResult on 4349 boards and the
 *      outcome counts of its own rollout loop -> tests/golden/ref_squava2_vectors.json
 *   2. the Go files themselves -- src/alphabeta + src/movekeeper (ChooseMove with deltaValue
 *      and deltaValue2), squavathr.go (deltaValue, alphaBeta), opening2.go / opening3.go
 *      (alphaBeta, Matrices, orderedCells), src/mcts (FindWinner, GetMoves, GetResult) --
 *      translated MECHANICALLY to Python in memory (oracle/tools/go2py.py: a statement-by-
 *      statement translator for the small Go subset they use) and executed
 *      (oracle/tools/make_ref_golden_go.py); what is not translatable (goroutines/channels
 *      in chooseMove, flag/time in main) is re-driven line for line
 *      -> tests/golden/ref_go_vectors.json.  Root values, chosen moves, maxima AND the
 *      order-dependent leaf counters of this file equal the Go code's.
 *   3. squava.html, the browser transliteration of the alpha-beta program, its <script>
 *      translated the same way (oracle/tools/make_ref_golden_js.py)
 *      -> tests/golden/ref_squava_html_vectors.json
 * plus every line / scan-order table against tests/golden/ref_tables.json (parsed from the
 * Go sources).  CAVEAT: 2 and 3 execute a mechanical translation, not a Go / JavaScript
 * implementation; Go's math/rand is not reproduced (random choices are compared
 * statistically).  An independent Python restatement (oracle/pyoracle.py) and SURVEY.md
 * Appendix B's vectors are further opinions.
 *
 * Deliberately array-and-table-walk code (an int board[25] and lists of cell
 * indices walked in the reference's own order), NOT bit tricks: it stands in for
 * the reference's algorithm, both as the checker and as the timed CPU baseline.
 *
 * Cell index c = 5*x + y, x = first coordinate everywhere in the reference
 * (src/mcts/mcts.go:55-58).  Marks: +1 MAXIMIZER 'X', -1 MINIMIZER 'O', 0 UNSET
 * (src/mcts/mcts.go:26-30).  Boards cross this API as two 25-bit masks.
 */
#ifndef SQUAVA_ORACLE_H
#define SQUAVA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t x, o; } sqo_board; /* bit c set <=> board[c] == +1 / -1 */

#define SQO_NO_MOVE INT32_MIN /* values[] entry for an occupied cell */

/* classification flag bits (same meaning as include/squava_b200.h) */
#define SQO_X4 1u
#define SQO_O4 2u
#define SQO_X3 4u
#define SQO_O3 8u
#define SQO_FULL 16u

/* ---- tables (for the table tests) -------------------------------------- */
/* Each returns the number of entries written.  Lines are cell lists in the
 * reference's own cell order.                                              */
int sqo_tables_mcts_important(int out[9]);                 /* mcts.go:404      */
int sqo_tables_mcts_quads(int cell, int out[][4]);         /* mcts.go:409-425  */
int sqo_tables_mcts_triplets(int cell, int out[][3]);      /* mcts.go:430-446  */
int sqo_tables_ab_quads(int out[28][4]);                   /* alphabeta.go:299-328 */
int sqo_tables_ab_triplets(int out[48][3]);                /* alphabeta.go:249-298 */
int sqo_tables_ordered_cells(int out[25]);                 /* opening3.go:328-361 */
int sqo_tables_scores(int which, int out[25]);             /* 0: alphabeta.go:343-349, 1: opening3.go:366-372 */
int sqo_tables_no2(int out[4][3]);                         /* squavathr.go:305-311 */
int sqo_tables_no_middle2(int out[4][4]);                  /* squavathr.go:297-302 */
int sqo_tables_matrices(int out[8][25]);                   /* opening3.go:379-451 */
int sqo_tables_first_moves(int out[6]);                    /* opening3.go:456-463 */

/* ---- terminal test / classification ------------------------------------ */
/* GetMoves' second return value: any owned quad or triplet, or no empty cell
 * (mcts.go:310-346).                                                        */
int sqo_mcts_terminal(sqo_board b);
/* (*MCTS).FindWinner, mcts.go:92-121.  +1 / -1 / 0.                         */
int sqo_mcts_find_winner(sqo_board b);
/* (*GameState).GetResult(playerjm), mcts.go:348-388.  1.0 or 0.0.           */
double sqo_mcts_get_result(sqo_board b, int playerjm);
/* (*AlphaBeta).FindWinner, alphabeta.go:355-378.                            */
int sqo_ab_find_winner(sqo_board b);
/* flags by plain counting over the 28 quads / 48 triplets.                  */
unsigned sqo_flags(sqo_board b);
/* batch form used by the parity tests: any output pointer may be NULL.      */
void sqo_classify(const sqo_board *b, int64_t n, uint8_t *flags,
                  int8_t *winner_mcts, int8_t *winner_ab);

/* every board with k marks and X-count = ceil(k/2), by rank (test enumeration) */
int64_t sqo_count_alternating(int k);
void sqo_enumerate_alternating(int k, int64_t start, int64_t count, sqo_board *out);

/* ---- random playouts (mcts.go:173-181 + GetResult) ---------------------- */
/* n_playouts uniformly random playouts from (b, to_move); out = {max_wins,
 * min_wins, draws, total_plies}.  RNG: xoshiro256** seeded by splitmix64(seed),
 * unbiased bounded draws (what Go's rand.Intn guarantees).                  */
void sqo_playouts(sqo_board b, int to_move, uint64_t n_playouts, uint64_t seed,
                  uint64_t out[4]);
/* many leaves, n_threads pthreads (leaf-interleaved); out[i] as above.  The
 * CPU baseline leg of bench.py.                                             */
void sqo_playouts_batch(const sqo_board *b, const int8_t *to_move, int64_t n_leaves,
                        uint64_t playouts_per_leaf, uint64_t seed, int n_threads,
                        uint64_t (*out)[4]);
/* EXACT outcome distribution of a uniformly random playout, by forward DP over
 * reachable states: p[0]=P(MAX wins) p[1]=P(MIN wins) p[2]=P(draw),
 * *mean_plies = E[plies].  Returns number of DP states visited, <0 on OOM.  */
int64_t sqo_playout_exact(sqo_board b, int to_move, double p[3], double *mean_plies);

/* ---- alpha-beta dialects ------------------------- */
/* AB-1 src/alphabeta ChooseMove (alphabeta.go:92-117) with movekeeper
 * (movekeeper.go:26-52): MAXIMIZER to move.  values[c] = value of root move c or
 * SQO_NO_MOVE; *best_mask = cells tied at the max; *best_value = max (0 if no
 * move, as movekeeper returns); *leaves = leafNodeCount.                    */
void sqo_ab1_root(sqo_board b, int max_depth, const int scores[25],
                  int32_t values[25], uint32_t *best_mask, int32_t *best_value,
                  int64_t *leaves);
/* AB-1 after SetAvoid() (alphabeta.go:473-475): the evaluator is deltaValue2 (:382-454),
 * i.e. deltaValue plus the no2 / noMiddle2 -100 terms; playoff5's "G" player.   */
void sqo_ab5_root(sqo_board b, int max_depth, const int scores[25],
                  int32_t values[25], uint32_t *best_mask, int32_t *best_value,
                  int64_t *leaves);
/* src/abbook's search (abbook.go:93-131,171-275; playoff5's "B" player once its book is done):
 * AB-1's delayed evaluation, but the recursion hands boardValue+delta down (:228-229,255-256). */
void sqo_ab6_root(sqo_board b, int max_depth, const int scores[25],
                  int32_t values[25], uint32_t *best_mask, int32_t *best_value,
                  int64_t *leaves);
/* AB-2 squavathr chooseMove (squavathr.go:236-269) + worker (:448-466).
 * max_depth is chooseMove's argument (it decrements it itself, :239).       */
void sqo_ab2_root(sqo_board b, int max_depth, const int scores[25],
                  int32_t values[25], uint32_t *best_mask, int32_t *best_value,
                  int64_t *leaves);
/* src/negascout ChooseMove (negascout.go:81-122: reorderMoves :411-499, negaScout :232-266, the
 * whole-board staticValue :168-230) with movekeeper.  The search is NOT a plain minimax (null
 * windows accepted two plies above the horizon, a running alpha handed to SetMove), so every
 * output depends on the visiting order: values[c] = the alpha recorded for root move c,
 * visit_order = the root cells in the order searched (-1 padded), *best_mask = what movekeeper
 * holds, *first_best = its deterministic choice (-1: no move), *leaves = staticValue calls.   */
void sqo_negascout_root(sqo_board b, int max_depth, const int scores[25], int32_t values[25],
                        int32_t visit_order[25], uint32_t *best_mask, int32_t *best_value,
                        int32_t *first_best, int64_t *leaves);
/* One explicit subtree: the board already holds the move `last_cell`; `side` is
 * to move at `ply`; `carried` is the boardValue argument.
 *  dialect 1: (*AlphaBeta).alphaBeta(ply, side, -20000, +20000, x, y, carried)  alphabeta.go:165
 *  dialect 2: squavathr alphaBeta(max_depth, bd, ply, side, -20000, +20000, x, y, carried) :389
 *             (max_depth here is the worker's, i.e. already decremented)
 *  dialect 3: opening3 alphaBeta(bd, ply, side, -10000, +10000, x, y, carried)  opening3.go:136
 *  dialect 4: opening2 alphaBeta(...)  opening2.go:91 (full-board value -+10000)
 *  dialect 5: dialect 1 with the deltaValue2 evaluator (SetAvoid)
 *  dialect 6: (*AlphaBetaBook).alphaBeta(...)  abbook.go:215 (dialect 1, boardValue+delta handed down) */
int32_t sqo_ab_subtree(int dialect, sqo_board b, int last_cell, int ply, int side,
                       int32_t carried, int max_depth, const int scores[25],
                       int64_t *leaves);

/* evaluator calls made by the calling thread's most recent sqo_ab* search */
int64_t sqo_last_evals(void);

/* opening3's symmetry de-duplication (opening3.go:66-128): writes the
 * (first, reply) cell pairs it would search, in its order; returns count.   */
int sqo_opening3_subtrees(int first[150], int reply[150]);

#ifdef __cplusplus
}
#endif
#endif
