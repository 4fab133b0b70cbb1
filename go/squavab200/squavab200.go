// Package squavab200 is the cgo binding of libsquava_b200.so (include/squava_b200.h): what a
// maintainer of bediger4000/squava adds to call the B200 hot path from the unmodified Go
// programs.  NOT COMPILED IN THIS REPO'S CI: the build image has no Go toolchain (see
// DESIGN.md); the same ABI is exercised by the C++ host programs and the ctypes tests.
//
// Build (on a box with Go >= 1.16 and the library installed):
//   CGO_CFLAGS="-I/path/to/repo/include" CGO_LDFLAGS="-L/path/to/repo/squava_b200 -lsquava_b200" go build ./...
package squavab200

/*
#include <stdlib.h>
#include "squava_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Board is the [25]int / [5][5]int board of the reference as two masks: bit 5*x+y.
type Board struct{ X, O uint32 }

// FromCells converts the reference's [25]int board (src/mcts/mcts.go:11).
func FromCells(cells *[25]int) Board {
	var b Board
	for i, v := range cells {
		switch v {
		case 1:
			b.X |= 1 << uint(i)
		case -1:
			b.O |= 1 << uint(i)
		}
	}
	return b
}

// FromGrid converts the reference's [5][5]int board (src/alphabeta/alphabeta.go:9).
func FromGrid(bd *[5][5]int) Board {
	var flat [25]int
	for i, row := range bd {
		for j, v := range row {
			flat[5*i+j] = v
		}
	}
	return FromCells(&flat)
}

// call runs ONE library call and, if it failed, reads the error detail.  sq_last_error() is a
// thread-local of the C library, so the goroutine is pinned to its OS thread across the pair:
// without that the scheduler may move it between the two cgo calls and the detail read would be
// empty or another call's.
func call(what string, f func() C.int) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != C.SQ_OK {
		// the reference has no error returns: failure is panic or os.Exit
		panic(fmt.Sprintf("%s: %s (%s)", what, C.GoString(C.sq_strerror(rc)), C.GoString(C.sq_last_error())))
	}
}

// Init selects the GPUs (nil = device 0) and the seed of every random stream.
func Init(devices []int, seed uint64) {
	if len(devices) == 0 {
		call("sq_init", func() C.int {
		return C.sq_init(nil, 0, C.uint64_t(seed))
	})
		return
	}
	ids := make([]C.int, len(devices))
	for i, d := range devices {
		ids[i] = C.int(d)
	}
	call("sq_init", func() C.int {
		return C.sq_init(&ids[0], C.int(len(ids)), C.uint64_t(seed))
	})
}

// Shutdown releases the devices.
func Shutdown() { C.sq_shutdown() }

// Classify: flags, FindWinner in mcts.go order, FindWinner in alphabeta.go order.
func Classify(boards []Board) (flags []uint8, winnerMCTS, winnerAB []int8) {
	n := len(boards)
	flags, winnerMCTS, winnerAB = make([]uint8, n), make([]int8, n), make([]int8, n)
	if n == 0 {
		return
	}
	call("sq_classify", func() C.int {
		return C.sq_classify((*C.sq_board)(unsafe.Pointer(&boards[0])), C.int64_t(n),
		(*C.uint8_t)(unsafe.Pointer(&flags[0])), (*C.int8_t)(unsafe.Pointer(&winnerMCTS[0])),
		(*C.int8_t)(unsafe.Pointer(&winnerAB[0])))
	})
	return
}

// Counts of playouts from one leaf: {MAXIMIZER wins, MINIMIZER wins, draws, plies}.
type Counts [4]uint64

// Playouts runs perLeaf random playouts from every leaf (replaces squavam.go:128-147).
func Playouts(leaves []Board, toMove []int8, perLeaf, streamID uint64) []Counts {
	n := len(leaves)
	out := make([]Counts, n)
	if n == 0 {
		return out
	}
	call("sq_playouts", func() C.int {
		return C.sq_playouts((*C.sq_board)(unsafe.Pointer(&leaves[0])), (*C.int8_t)(unsafe.Pointer(&toMove[0])),
		C.int64_t(n), C.uint64_t(perLeaf), C.uint64_t(streamID),
		(*[4]C.uint64_t)(unsafe.Pointer(&out[0])))
	})
	return out
}

// RootValues is one row of AlphaBetaRoot.
type RootValues struct {
	Values    [25]int32 // NoMove where occupied
	BestValue int32
	BestMask  uint32 // cells tied at BestValue: movekeeper's set
	Nodes     uint64
}

// NoMove marks an occupied cell in RootValues.Values.
const NoMove = int32(-1 << 31)

// Dialects of the reference's alpha-beta programs.
const (
	AlphaBeta = 1 // src/alphabeta
	Squavathr = 2 // squavathr.go
	Opening3  = 3 // opening3.go
	Opening2  = 4 // opening2.go
	Avoid     = 5 // src/alphabeta after SetAvoid(): deltaValue2
	ABBook    = 6 // src/abbook: boardValue+delta handed down
)

// AlphaBetaRoot replaces ChooseMove's loop (alphabeta.go:98-110) / chooseMove + the worker
// pool (squavathr.go:236-269,448-466) for a batch of positions, MAXIMIZER to move.
func AlphaBetaRoot(positions []Board, dialect, maxDepth int, scores *[25]int8) []RootValues {
	n := len(positions)
	out := make([]RootValues, n)
	if n == 0 {
		return out
	}
	values := make([][25]int32, n)
	bv := make([]int32, n)
	bm := make([]uint32, n)
	nodes := make([]uint64, n)
	call("sq_alphabeta_root", func() C.int {
		return C.sq_alphabeta_root((*C.sq_board)(unsafe.Pointer(&positions[0])), C.int64_t(n), C.int(dialect),
		C.int(maxDepth), (*C.int8_t)(unsafe.Pointer(&scores[0])), (*[25]C.int32_t)(unsafe.Pointer(&values[0])),
		(*C.int32_t)(unsafe.Pointer(&bv[0])), (*C.uint32_t)(unsafe.Pointer(&bm[0])),
		(*C.uint64_t)(unsafe.Pointer(&nodes[0])))
	})
	for i := range out {
		out[i] = RootValues{values[i], bv[i], bm[i], nodes[i]}
	}
	return out
}

// NegaScoutResult is what (*NegaScout).ChooseMove needs from its search (negascout.go:81-122).
type NegaScoutResult struct {
	Values     [25]int32 // alpha handed to moves.SetMove per root cell, NoMove if not visited
	VisitOrder [25]int8  // root cells in the order searched, -1 padded
	BestValue  int32
	BestMask   uint32 // the cells movekeeper holds
	FirstBest  int8   // movekeeper's deterministic choice, -1 if the board is full
	Nodes      uint64 // staticValue calls = p.leafNodeCount
}

// NegaScoutRoot replaces reorderMoves + the root loop + negaScout (negascout.go:86-116,232-266).
func NegaScoutRoot(positions []Board, maxDepth int, scores *[25]int8) []NegaScoutResult {
	n := len(positions)
	out := make([]NegaScoutResult, n)
	if n == 0 {
		return out
	}
	values := make([][25]int32, n)
	order := make([][25]int8, n)
	bv := make([]int32, n)
	bm := make([]uint32, n)
	fb := make([]int8, n)
	nodes := make([]uint64, n)
	call("sq_negascout_root", func() C.int {
		return C.sq_negascout_root((*C.sq_board)(unsafe.Pointer(&positions[0])), C.int64_t(n), C.int(maxDepth),
		(*C.int8_t)(unsafe.Pointer(&scores[0])), (*[25]C.int32_t)(unsafe.Pointer(&values[0])),
		(*[25]C.int8_t)(unsafe.Pointer(&order[0])), (*C.int32_t)(unsafe.Pointer(&bv[0])),
		(*C.uint32_t)(unsafe.Pointer(&bm[0])), (*C.int8_t)(unsafe.Pointer(&fb[0])),
		(*C.uint64_t)(unsafe.Pointer(&nodes[0])))
	})
	for i := range out {
		out[i] = NegaScoutResult{values[i], order[i], bv[i], bm[i], fb[i], nodes[i]}
	}
	return out
}

// MctsResult is what UCT() leaves at the root (mcts.go:196-198): the visited children of the root and their counts.
type MctsResult struct {
	ChildMove                []int // cells, ascending
	ChildVisits, ChildWins   []float64
	RootVisits               float64
	Iterations, Batches      int64
	Nodes                    uint64
	ArenaFull                bool
	DeviceSeconds            float64
}

// MctsDecide replaces UCT(rootstate, itermax, UCTK) (mcts.go:140-200) for searches that keep no tree between moves:
// the whole tree lives in GPU memory, select / playouts / update are device kernels (DESIGN 3.5).
// playerJustMoved: +1 X, -1 O (GameState.playerJustMoved).  The caller picks argmax UCB1 as bestMove() does.
func MctsDecide(slot int, root Board, playerJustMoved int, iterations int64, uctk float64, ucbFactor2 bool, batch int64,
	playoutsPerLeaf int, streamBase, seed uint64) MctsResult {
	var r C.sq_mcts_result
	f2 := C.int(0)
	if ucbFactor2 {
		f2 = 1
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := C.sq_mcts_decide(C.int(slot), (*C.sq_board)(unsafe.Pointer(&root)), C.int(playerJustMoved), C.int64_t(iterations),
		C.double(uctk), f2, C.int64_t(batch), C.int(playoutsPerLeaf), 0, -1, 0, C.uint64_t(streamBase), C.uint64_t(seed), &r); rc != C.SQ_OK {
		panic(fmt.Sprintf("sq_mcts_decide: %s (%s)", C.GoString(C.sq_strerror(rc)), C.GoString(C.sq_mcts_last_error())))
	}
	out := MctsResult{RootVisits: float64(r.root_visits), Iterations: int64(r.iterations), Batches: int64(r.batches),
		Nodes: uint64(r.nodes), ArenaFull: r.arena_full != 0, DeviceSeconds: float64(r.device_ms) * 1e-3}
	for i := 0; i < int(r.n_children); i++ {
		out.ChildMove = append(out.ChildMove, int(r.child_move[i]))
		out.ChildVisits = append(out.ChildVisits, float64(r.child_visits[i]))
		out.ChildWins = append(out.ChildWins, float64(r.child_wins[i]))
	}
	return out
}

// Subtree is one unit of squavathr's toDo channel / one top-level call of opening2/3.
type Subtree struct {
	B          Board
	LastCell   int8
	Ply        int8
	SideToMove int8
	_          int8
	Carried    int32
}

// AlphaBetaSubtrees evaluates explicit subtrees; value[i] is what the reference's alphaBeta returns.
func AlphaBetaSubtrees(st []Subtree, dialect, maxDepth int, scores *[25]int8) (value []int32, nodes []uint64) {
	n := len(st)
	value, nodes = make([]int32, n), make([]uint64, n)
	if n == 0 {
		return
	}
	call("sq_alphabeta_subtrees", func() C.int {
		return C.sq_alphabeta_subtrees((*C.sq_subtree)(unsafe.Pointer(&st[0])), C.int64_t(n), C.int(dialect),
		C.int(maxDepth), (*C.int8_t)(unsafe.Pointer(&scores[0])), (*C.int32_t)(unsafe.Pointer(&value[0])),
		(*C.uint64_t)(unsafe.Pointer(&nodes[0])))
	})
	return
}
