/*
 * squava_b200.h -- C ABI of libsquava_b200.so: the B200 (sm_100a) implementation
 * of squava's data-parallel hot path.
 *
 * The reference (bediger4000/squava) is pure Go with no FFI of its own; each
 * entry point below replaces the Go code a maintainer would cut out and bind
 * with cgo (go/squavab200/squavab200.go shows the binding; INTEGRATION.md the
 * call sites).  Reference citations are file:line in the reference tree.
 *
 * Conventions
 *   - Every function returns SQ_OK (0) or a negative sq_status; nothing throws,
 *     nothing aborts.  sq_strerror() names a status; sq_last_error() gives the
 *     CUDA detail of the calling thread's last failure.
 *   - The caller owns every buffer; the library keeps no pointer after a call
 *     returns (the cgo pointer rule).  Calls may arrive on any OS thread.  Calls on
 *     different devices never wait for each other; on one device up to four calls
 *     are in flight at a time (each on its own stream and staging buffers), further
 *     ones queue.  Only sq_init / sq_shutdown exclude everything else.
 *   - There is NO CPU fallback: without a usable sm_100 device sq_init() fails
 *     and every compute call returns SQ_ERR_NOT_INIT.
 *   - Board: two 25-bit masks, bit c = 5*x + y with x the first coordinate
 *     everywhere in the reference (src/mcts/mcts.go:55-58, squavam.go:362);
 *     x-mask = MAXIMIZER (+1, 'X'), o-mask = MINIMIZER (-1, 'O').
 */
#ifndef SQUAVA_B200_H
#define SQUAVA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_ABI_VERSION 3

typedef enum sq_status {
    SQ_OK = 0,
    SQ_ERR_INVALID = -1,      /* bad argument (null pointer, overlapping masks, bad dialect ...) */
    SQ_ERR_NOT_INIT = -2,     /* sq_init() has not succeeded */
    SQ_ERR_CUDA = -3,         /* CUDA runtime error; see sq_last_error() */
    SQ_ERR_NO_DEVICE = -4,    /* no CUDA device / not compute capability 10.x */
    SQ_ERR_UNSUPPORTED = -5,  /* combination the reference has no counterpart for */
    SQ_ERR_NOMEM = -6
} sq_status;

typedef struct sq_board { uint32_t x, o; } sq_board;

#define SQ_MAXIMIZER 1
#define SQ_MINIMIZER (-1)
#define SQ_NO_MOVE INT32_MIN /* values[][c] for an occupied cell */

/* classification flags */
#define SQ_X4 1u   /* X owns a 4-line */
#define SQ_O4 2u
#define SQ_X3 4u   /* X owns a 3-line */
#define SQ_O3 8u
#define SQ_FULL 16u /* no empty cell */

/* alpha-beta dialects */
#define SQ_AB_ALPHABETA 1 /* src/alphabeta/alphabeta.go:92-223   -- the bit-exact contract */
#define SQ_AB_SQUAVATHR 2 /* squavathr.go:236-269,315-446 */
#define SQ_AB_OPENING3 3  /* opening3.go:136-225 */
#define SQ_AB_OPENING2 4  /* opening2.go:91-180 (full-board node value -+10000) */
#define SQ_AB_ALPHABETA_AVOID 5 /* src/alphabeta after SetAvoid(): deltaValue2, alphabeta.go:382-454,473-475 */
#define SQ_AB_ABBOOK 6    /* src/abbook/abbook.go:93-131,171-275: dialect 1 with boardValue+delta handed down */

/* ---- lifecycle -------------------------------------------------------- */
/* device_ids == NULL / n_devices <= 0: use device 0.  seed keys every random
 * stream (the reference seeds math/rand from the clock, squavam.go:48).  With
 * n_devices > 1 the host-buffer calls below shard their units over the devices
 * (contiguous ranges; explicit subtrees by estimated size; no collective); results do not
 * depend on n_devices.                                                     */
int sq_init(const int *device_ids, int n_devices, uint64_t seed);
void sq_shutdown(void);
const char *sq_strerror(int status);
const char *sq_last_error(void);
/* The same message copied into the caller's buffer.  sq_last_error() reads a thread-local: a cgo
 * caller whose goroutine may migrate between two C calls either pins its OS thread around the pair
 * (runtime.LockOSThread) or calls this from the same C stub as the failing call.               */
int sq_last_error_copy(char *buf, int len);
int sq_abi_version(void);
int sq_device_count(void);          /* devices this library was initialised on */
int sq_sm_count(int slot);          /* multiprocessors of device slot (0 <= slot < sq_device_count()) */
int sq_device_id(int slot);         /* CUDA device ordinal of device slot, or a negative status             */

/* ---- terminal-position test ------------------------------------------- */
/* Replaces (*MCTS).FindWinner src/mcts/mcts.go:92-121 (winner_mcts_order) and
 * (*AlphaBeta).FindWinner src/alphabeta/alphabeta.go:355-378 / findWinner
 * squavathr.go:271-294 (winner_ab_order) over a batch; flags also carries what
 * GetMoves' terminal test (mcts.go:310-346) needs: terminal <=> flags != 0.
 * Winners are +1 / -1 / 0; the two differ only on boards where both colours own
 * a line (different scan orders).  Any output pointer may be NULL.          */
int sq_classify(const sq_board *boards, int64_t n, uint8_t *flags,
                int8_t *winner_mcts_order, int8_t *winner_ab_order);

/* ---- random playouts --------------------------------------------------- */
/* Replaces the rollout block of UCT(): squavam.go:128-136 / src/mcts/mcts.go:
 * 173-181 / squavam2.go:152-160 (Clone, GetMoves, rand.Intn, DoMove until
 * terminal) and the GetResult calls of the backprop loop (mcts.go:190-192,
 * 348-388), for playouts_per_leaf independent playouts from each leaf.
 * to_move[i] is the side about to move (+1/-1), i.e. -playerJustMoved.
 * out[i] = {MAXIMIZER wins, MINIMIZER wins, draws, total plies played}.
 * A node whose playerJustMoved is +1 adds out[i][0] to its wins, one whose
 * playerJustMoved is -1 adds out[i][1]; draws add 0 to both (mcts.go:386-387).
 * Random stream: Philox4x32-10 keyed by (seed, stream_id, leaf_index_base + i,
 * sub-stream): the same call returns the same counts on any number of devices. */
int sq_playouts(const sq_board *leaves, const int8_t *to_move, int64_t n_leaves,
                uint64_t playouts_per_leaf, uint64_t stream_id, uint64_t (*out)[4]);

/* One playout per leaf (what UCT's loop does per iteration, mcts.go:173-192) with the result packed
 * into ONE byte per leaf instead of four counters: bits 0-1 = 1 MAXIMIZER won, 2 MINIMIZER won,
 * 3 draw (GetResult 0.0 for both); bits 2-7 = plies played (0 <=> the leaf was terminal, which is
 * what a freshly expanded node needs to know).  Same random stream as sq_playouts(.., 1, ..):
 * outcome[i] encodes exactly the counts that call would return.  For hosts that feed 64K-leaf
 * batches at 10^7 leaves/s (configs[2]): 1 B instead of 32 B back per leaf.                    */
int sq_playouts_packed(const sq_board *leaves, const int8_t *to_move, int64_t n_leaves,
                       uint64_t stream_id, uint8_t *outcome);

/* The two calls above on ONE device of those sq_init() named (slot = index into that list) instead of
 * sharded over all of them: what a root-parallel host uses -- one tree, one host-thread group and one
 * device per searcher (squavam2.go:93-112 runs N searchers for one decision), root statistics merged
 * once per decision.  The stream depends on (seed, stream_id, leaf index) only, so a leaf gives the same
 * counts on every device.                                                                         */
int sq_playouts_on(int slot, const sq_board *leaves, const int8_t *to_move, int64_t n_leaves,
                   uint64_t playouts_per_leaf, uint64_t stream_id, uint64_t (*out)[4]);
int sq_playouts_packed_on(int slot, const sq_board *leaves, const int8_t *to_move, int64_t n_leaves,
                          uint64_t stream_id, uint8_t *outcome);

/* ---- fixed-depth alpha-beta -------------------------------------------- */
/* Replaces ChooseMove's per-root-move loop: src/alphabeta/alphabeta.go:92-117
 * (dialect 1, or 5 = the same engine after SetAvoid(); max_depth = p.maxDepth as SetDepth
 * leaves it, :79-89), src/abbook/abbook.go:104-123 (dialect 6, SetDepth :80-90) or
 * squavathr chooseMove + its worker pool, squavathr.go:236-269,448-466 (dialect
 * 2, max_depth = chooseMove's argument).  MAXIMIZER is to move in every
 * position.  scores = the 25-entry bias table (alphabeta.go:330,343-349).
 * values[i][c] = value of root move c (SQ_NO_MOVE where occupied);
 * best_value[i] / best_mask[i] = movekeeper's max and the cells tied at it
 * (src/movekeeper/movekeeper.go:26-52; 0 / 0 if the board is full).
 * nodes[i] (may be NULL) = evaluator calls the GPU search spent on position i
 * (order-dependent, like the reference's leaf counter; not a parity quantity). */
int sq_alphabeta_root(const sq_board *positions, int64_t n, int dialect, int max_depth,
                      const int8_t scores[25], int32_t (*values)[25],
                      int32_t *best_value, uint32_t *best_mask, uint64_t *nodes);

/* ---- NegaScout ----------------------------------------------------------- */
/* Replaces (*NegaScout).ChooseMove's search, src/negascout/negascout.go:81-122 (reorderMoves
 * :411-499, negaScout :232-266, staticValue :168-230), for a batch of positions with
 * MAXIMIZER to move; max_depth = p.maxDepth as SetDepth leaves it (:69-79).  This search is
 * not a plain minimax (null-window results are accepted two plies above the horizon, the
 * root hands its running alpha to movekeeper), so the GPU replays the reference's recursion
 * step for step -- one warp per position, and a search that runs long spread over many warps
 * as calls of negaScout whose results are consumed in the reference's order -- and every
 * output is bit-exact including nodes[]:
 * values[i][c] = the alpha recorded by moves.SetMove for root move c (SQ_NO_MOVE if not
 * visited); visit_order[i][k] = k-th root cell searched, -1 padded (may be NULL);
 * best_value / best_mask = movekeeper's max and the cells it holds; first_best[i] = the cell
 * it returns when deterministic (-1: no legal move); nodes[i] = staticValue calls
 * (p.leafNodeCount).  Output pointers other than values may be NULL.            */
int sq_negascout_root(const sq_board *positions, int64_t n, int max_depth,
                      const int8_t scores[25], int32_t (*values)[25], int8_t (*visit_order)[25],
                      int32_t *best_value, uint32_t *best_mask, int8_t *first_best,
                      uint64_t *nodes);

/* One unit of squavathr's toDo channel (squavathr.go:17-24,448-466) or one of
 * opening2/opening3's top-level calls (opening2.go:62-71, opening3.go:104-109):
 * the board already holds the move at last_cell, side_to_move moves at `ply`,
 * `carried` is the boardValue argument.  max_depth is the callee's (squavathr's
 * worker sees chooseMove's value minus one).  value[i] = what the reference's
 * alphaBeta returns for that call.  |carried| must not exceed 15000
 * (SQ_ERR_UNSUPPORTED; the reference's callers pass 0 or one evaluation).   */
typedef struct sq_subtree {
    sq_board b;
    int8_t last_cell;     /* 0..24 */
    int8_t ply;
    int8_t side_to_move;  /* +1 / -1 */
    int8_t reserved;
    int32_t carried;
} sq_subtree;
int sq_alphabeta_subtrees(const sq_subtree *subtrees, int64_t n, int dialect, int max_depth,
                          const int8_t scores[25], int32_t *value, uint64_t *nodes);

/* ---- one UCT decision, tree in HBM ---------------------------------------
 * Replaces UCT(rootstate, itermax, UCTK) of src/mcts/mcts.go:140-200 as squavam.go:60-78 calls it, for throughput
 * runs (BASELINE configs[2]: 10^9 one-playout iterations through ONE tree).  The tree lives in device memory of
 * device slot `slot`; an iteration batch is three launches on one stream -- select (`walkers` threads, each taking
 * its share of the batch's walks one after the other: visits are added on the way down, so later walks see earlier
 * ones), the playout kernel on the leaves (random stream ids stream_base + batch number), update -- and only the
 * root's children come back.  Select is UCB1 = w/v + uctk*sqrt(f*ln(N)/v), f = 2 if ucb_factor2 else 1, first
 * maximum in cell order; expansion takes a random untried move (mcts.go:166-171).  player_just_moved: +1 X, -1 O
 * (GameState.playerJustMoved).  sync_levels (-1: default 4): on that many levels below the root the walks of a batch
 * move level by level, and the K walks that stand on one node share out its children as K sequential selections
 * with virtual loss would (0: every walk on its own from the root -- thousands that arrive together then take the
 * same child).  max_nodes 0: the arena is sized from the iteration count and the free memory.
 * The caller picks the move: argmax of UCB1 over the children (mcts.go:196-198).                               */
typedef struct {
    int32_t n_children;
    int32_t child_move[25];          /* cell of every root child that was visited, ascending           */
    double child_visits[25], child_wins[25];
    double root_visits;
    int64_t iterations, batches;
    uint64_t nodes, arena_slots;     /* tree slots handed out / reserved                               */
    int32_t arena_full;              /* 1: the arena ran out and the search went on without growing    */
    double device_ms;                /* first select to last update, CUDA events                       */
} sq_mcts_result;
int sq_mcts_decide(int slot, const sq_board *root, int player_just_moved, int64_t iterations, double uctk,
                   int ucb_factor2, int64_t batch, int playouts_per_leaf, int walkers, int sync_levels,
                   uint64_t max_nodes, uint64_t stream_base, uint64_t seed, sq_mcts_result *out);
const char *sq_mcts_last_error(void);   /* message of the last failed sq_mcts_decide on this thread    */

/* ---- device-resident forms ---------------------------------------------
 * Same semantics, but every data pointer is DEVICE memory on device slot
 * `slot`, the work is enqueued on `stream` (a cudaStream_t; NULL = the legacy
 * default stream) and the call returns without synchronising.  These are what
 * a host that already keeps its leaves in HBM calls (bench.py's `value` leg,
 * the leaf-parallel MCTS driver's double-buffered batches).                 */
int sq_classify_dev(int slot, const sq_board *boards, int64_t n, uint8_t *flags,
                    int8_t *winner_mcts_order, int8_t *winner_ab_order, void *stream);
/* out must hold n_leaves*4 uint64 and is OVERWRITTEN (zeroed, then accumulated). */
int sq_playouts_dev(int slot, const sq_board *leaves, const int8_t *to_move, int64_t n_leaves,
                    uint64_t leaf_index_base, uint64_t playouts_per_leaf, uint64_t stream_id,
                    uint64_t *out, void *stream);

/* ---- measurement helpers ------------------------------------------------ */
/* INT32 issue-rate probe (the roofline denominator):
 * runs dependent-chain kernels on every SM of device slot and reports
 * thread-operations per second.  kind 0: LOP3 only (ALU pipe), 1: IMAD only
 * (FMA pipe), 2: LOP3+IMAD interleaved.  *ops_per_s is the best of `reps`.    */
int sq_int32_peak(int slot, int kind, int reps, double *ops_per_s);
/* Kernel launches issued by this library since sq_init() (all devices).      */
uint64_t sq_launch_count(void);
/* Duration in ms of the most recent playout kernel on device slot, measured
 * with CUDA events on the stream it ran on (synchronises that stream).       */
int sq_last_playout_kernel_ms(int slot, float *ms);
/* 1 if this library was built with -DSQ_CHECKED=1 (`make -C squava_b200/csrc checked`): the kernels then carry bounds
 * checks of their own (compute-sanitizer is closed on the development pool, profiles/r02_sanitizer.md).
 * sq_debug_check_failures: out = {failed checks since the last call, code of the first, its two arguments}, all
 * zero in an unchecked build.                                                                              */
int sq_abi_checked(void);
int sq_debug_check_failures(uint32_t out[4]);
/* Philox4x32-10 block on the device (known-answer tests of the RNG).         */
int sq_debug_philox(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* SQUAVA_B200_H */
