// players.go -- Player adapters over the B200 library: drop-ins for the engines playoff5.go builds in
// createPlayers (playoff5.go:185-221).  Each type has the reference's seven-method Player surface
// (playoff5.go:21-29) and the constructors / setters playoff5 type-asserts for, so a maintainer changes only the
// import path in createPlayers:
//
//	case "A": p = squavab200.NewAlphaBeta(deterministic, maxDepth)   // was alphabeta.New(...)
//	case "G": a := squavab200.NewAlphaBeta(...); a.SetAvoid(); p = a
//	case "N": p = squavab200.NewNegaScout(deterministic, maxDepth)   // was negascout.New(...)
//	case "M": p = squavab200.NewMCTS(deterministic, maxDepth)        // was mcts.New(...)
//
// What stays in Go is what the reference keeps outside its hot loops: the board the engine owns, SetDepth's table,
// movekeeper's tie handling (src/movekeeper/movekeeper.go:26-52) and the MCTS tree.  NOT COMPILED HERE (no Go
// toolchain in the build image, DESIGN.md); the C++ host library squava_b200/host is the tested twin of this file.
package squavab200

import (
	"fmt"
	"math"
	"math/rand"
)

const (
	maximizer = 1
	minimizer = -1
	unset     = 0
	loss      = -10000
)

// Scores is the package-level bias table every alpha-beta instance shares (alphabeta.go:330).
var Scores [25]int8

var defaultScores = [25]int8{3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3} // alphabeta.go:343-349

// keepMoves is movekeeper.SetMove / ChooseMove over a finished row of root values.
func keepMoves(values *[25]int32, deterministic bool) (x, y, value int) {
	max, n := 2*loss, 0
	var cells [25]int
	for c, v := range values {
		if v == NoMove || int(v) < max {
			continue
		}
		if int(v) > max {
			max, n = int(v), 0
		}
		cells[n] = c
		n++
	}
	if n == 0 {
		return -1, -1, 0 // cat got the game
	}
	r := 0
	if !deterministic {
		r = rand.Intn(n)
	}
	return cells[r] / 5, cells[r] % 5, max
}

func printGrid(b Board) { // PrintBoard, alphabeta.go:226-247
	fmt.Printf("   0 1 2 3 4\n")
	for i := 0; i < 5; i++ {
		fmt.Printf("%d  ", i)
		for j := 0; j < 5; j++ {
			c := uint(5*i + j)
			switch {
			case b.X>>c&1 == 1:
				fmt.Printf("X ")
			case b.O>>c&1 == 1:
				fmt.Printf("O ")
			default:
				fmt.Printf("_ ")
			}
		}
		fmt.Printf("\n")
	}
}

func (b *Board) set(x, y, player int) {
	bit := uint32(1) << uint(5*x+y)
	b.X &^= bit
	b.O &^= bit
	if player == maximizer {
		b.X |= bit
	} else if player == minimizer {
		b.O |= bit
	}
}

// ---- alpha-beta (src/alphabeta, and src/abbook's search) ---------------------------------------------

// AlphaBetaPlayer replaces alphabeta.AlphaBeta: ChooseMove's 25 root searches are ONE AlphaBetaRoot call.
type AlphaBetaPlayer struct {
	bd            Board
	name          string
	maxDepth      int
	deterministic bool
	dialect       int
}

func NewAlphaBeta(deterministic bool, maxdepth int) *AlphaBetaPlayer {
	return &AlphaBetaPlayer{name: "AlphaBeta", maxDepth: maxdepth, deterministic: deterministic, dialect: AlphaBeta}
}
func (p *AlphaBetaPlayer) Name() string                 { return p.name }
func (p *AlphaBetaPlayer) MakeMove(x, y int, player int) { p.bd.set(x, y, player) }
func (p *AlphaBetaPlayer) SetAvoid()                    { p.name = "A/B+Avoid"; p.dialect = Avoid } // alphabeta.go:473-475
func (p *AlphaBetaPlayer) PrintBoard()                  { printGrid(p.bd) }
func (p *AlphaBetaPlayer) SetDepth(moveCounter int) { // alphabeta.go:79-89
	if moveCounter < 4 {
		p.maxDepth = 8
	}
	if moveCounter > 3 {
		p.maxDepth = 10
	}
	if moveCounter > 10 {
		p.maxDepth = 12
	}
}
func (p *AlphaBetaPlayer) SetScores(randomize bool) { // alphabeta.go:334-351
	if randomize {
		vals := [11]int8{-5, -4, -3 - 2, -1, 0, 1, 2, 3, 4, 5, 0}
		for i := range Scores {
			Scores[i] = vals[rand.Intn(11)]
		}
		return
	}
	Scores = defaultScores
}
func (p *AlphaBetaPlayer) ChooseMove() (x, y, value, leafCount int) { // alphabeta.go:92-117
	r := AlphaBetaRoot([]Board{p.bd}, p.dialect, p.maxDepth, &Scores)[0]
	x, y, value = keepMoves(&r.Values, p.deterministic)
	if x >= 0 {
		p.MakeMove(x, y, maximizer)
	}
	return x, y, value, int(r.Nodes)
}
func (p *AlphaBetaPlayer) FindWinner() int { // alphabeta.go:355-378
	_, _, ab := Classify([]Board{p.bd})
	return int(ab[0])
}

// ---- NegaScout (src/negascout) ---------------------------------------------------------------------------

type NegaScoutPlayer struct {
	bd            Board
	maxDepth      int
	deterministic bool
}

func NewNegaScout(deterministic bool, maxdepth int) *NegaScoutPlayer {
	return &NegaScoutPlayer{maxDepth: maxdepth, deterministic: deterministic}
}
func (p *NegaScoutPlayer) Name() string                 { return "NegaScout" }
func (p *NegaScoutPlayer) MakeMove(x, y int, player int) { p.bd.set(x, y, player) }
func (p *NegaScoutPlayer) PrintBoard()                  { printGrid(p.bd) }
func (p *NegaScoutPlayer) SetScores(randomize bool)     { (&AlphaBetaPlayer{}).SetScores(randomize) }
func (p *NegaScoutPlayer) SetDepth(moveCounter int) { // negascout.go:69-79
	if moveCounter < 4 {
		p.maxDepth = 6
	}
	if moveCounter > 3 {
		p.maxDepth = 8
	}
	if moveCounter > 10 {
		p.maxDepth = 10
	}
}
func (p *NegaScoutPlayer) ChooseMove() (x, y, value, leafCount int) { // negascout.go:81-122
	r := NegaScoutRoot([]Board{p.bd}, p.maxDepth, &Scores)[0]
	cell := int(r.FirstBest)
	if !p.deterministic && r.BestMask != 0 { // movekeeper picks uniformly among the cells it holds
		var held []int
		for _, c := range r.VisitOrder { // in the order the search visited them
			if c >= 0 && r.BestMask>>uint(c)&1 == 1 {
				held = append(held, int(c))
			}
		}
		cell = held[rand.Intn(len(held))]
	}
	if cell < 0 {
		return -1, -1, 0, int(r.Nodes)
	}
	p.MakeMove(cell/5, cell%5, maximizer)
	return cell / 5, cell % 5, int(r.BestValue), int(r.Nodes)
}
func (p *NegaScoutPlayer) FindWinner() int {
	_, _, ab := Classify([]Board{p.bd})
	return int(ab[0])
}

// ---- MCTS (src/mcts) ---------------------------------------------------------------------------------------

type node struct { // mcts.go:15-23
	move            int
	parent          *node
	children        []*node
	wins, visits    float64
	untried         []int
	playerJustMoved int
	fresh           bool // created in the batch in flight: learns from its own rollout whether it is terminal
}

// MCTSPlayer replaces mcts.MCTS.  UCT keeps select / expand / backprop (mcts.go:123-199) and hands the
// rollout block (:173-181) plus GetResult (:190-192) to the GPU, Batch leaves at a time.
type MCTSPlayer struct {
	bd              Board
	playerJustMoved int
	iterations      int
	uctk            float64
	Batch, PerLeaf  int
	root            *node
	stream          uint64
}

func NewMCTS(deterministic bool, maxdepth int) *MCTSPlayer {
	return &MCTSPlayer{playerJustMoved: minimizer, iterations: 500000, uctk: 1.0, Batch: 256, PerLeaf: 1} // mcts.go:39-46
}
func (p *MCTSPlayer) Name() string         { return "MCTS" }
func (p *MCTSPlayer) SetUCTK(k float64)    { p.uctk = k }
func (p *MCTSPlayer) SetIterations(n int)  { p.iterations = n }
func (p *MCTSPlayer) SetDepth(int)         {}
func (p *MCTSPlayer) SetScores(bool)       {}
func (p *MCTSPlayer) PrintBoard()          { printGrid(p.bd) }
func (p *MCTSPlayer) FindWinner() int { // mcts.go:92-121
	_, m, _ := Classify([]Board{p.bd})
	return int(m[0])
}
func (p *MCTSPlayer) MakeMove(x, y int, player int) { // mcts.go:78-90, updateMoves :297-308
	p.bd.set(x, y, player)
	p.playerJustMoved = player
	if p.root != nil {
		var keep *node
		for _, c := range p.root.children {
			if c.move == 5*x+y {
				keep = c
			}
		}
		if keep != nil {
			keep.parent = nil
		}
		p.root = keep
	}
}

func empties(b Board) []int {
	var m []int
	for c := 0; c < 25; c++ {
		if (b.X|b.O)>>uint(c)&1 == 0 {
			m = append(m, c)
		}
	}
	return m
}

func (n *node) ucb1(k float64) float64 { // mcts.go:248-250 (with the factor 2)
	const eps = math.SmallestNonzeroFloat64
	return n.wins/(n.visits+eps) + k*math.Sqrt(2*math.Log(n.parent.visits)/(n.visits+eps))
}
func (n *node) bestMove(k float64) *node { // mcts.go:218-238
	var best *node
	score := math.SmallestNonzeroFloat64
	for _, c := range n.children {
		if u := c.ucb1(k); u > score {
			best, score = c, u
		}
	}
	return best
}

func (p *MCTSPlayer) ChooseMove() (x, y, value, leafCount int) { // mcts.go:62-76,123-199
	if p.root == nil {
		p.root = &node{move: -1, playerJustMoved: p.playerJustMoved, untried: empties(p.bd)}
	}
	type pick struct {
		n  *node
		b  Board
		tm int8
	}
	done := 0
	for done < p.iterations {
		per := p.PerLeaf
		if per > p.iterations-done {
			per = p.iterations - done
		}
		batch := p.Batch
		if batch*per > p.iterations-done {
			batch = (p.iterations - done) / per
		}
		picks := make([]pick, 0, batch)
		for i := 0; i < batch; i++ {
			n, b, pjm := p.root, p.bd, p.playerJustMoved
			for !n.fresh && len(n.untried) == 0 && len(n.children) > 0 { // select
				n = n.bestMove(p.uctk)
				pjm = -pjm
				b.set(n.move/5, n.move%5, pjm)
			}
			if len(n.untried) > 0 { // expand
				k := rand.Intn(len(n.untried))
				m := n.untried[k]
				n.untried = append(n.untried[:k], n.untried[k+1:]...)
				pjm = -pjm
				b.set(m/5, m%5, pjm)
				c := &node{move: m, parent: n, playerJustMoved: pjm, fresh: true}
				n.children = append(n.children, c)
				n = c
			}
			for q := n; q != nil; q = q.parent { // visits at selection time (virtual loss), wins on return
				q.visits += float64(per)
			}
			picks = append(picks, pick{n, b, int8(-pjm)})
		}
		leaves := make([]Board, len(picks))
		toMove := make([]int8, len(picks))
		for i, pk := range picks {
			leaves[i], toMove[i] = pk.b, pk.tm
		}
		p.stream++
		out := Playouts(leaves, toMove, uint64(per), p.stream)
		for i, pk := range picks {
			if pk.n.fresh {
				pk.n.fresh = false
				if out[i][3] != 0 { // zero plies <=> the leaf was terminal
					pk.n.untried = empties(pk.b)
				}
			}
			for q := pk.n; q != nil; q = q.parent { // Update, mcts.go:268-271: GetResult(q.playerJustMoved)
				if q.playerJustMoved == maximizer {
					q.wins += float64(out[i][0])
				} else {
					q.wins += float64(out[i][1])
				}
			}
		}
		done += len(picks) * per
	}
	best := p.root.bestMove(p.uctk)
	x, y, value = best.move/5, best.move%5, int(1000.*best.ucb1(p.uctk)) // mcts.go:196-198
	p.MakeMove(x, y, maximizer)
	return x, y, value, done
}
