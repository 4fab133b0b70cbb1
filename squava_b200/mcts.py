"""Python mirror of the reference's UCT driver (src/mcts/mcts.go:123-199, squavam.go:98-157)
with the rollout block replaced by batched GPU playouts, plus root-parallel merging across
ranks.  The production host program is the C++ one (squava_b200/host); this mirror exists
for torchrun-style multi-GPU runs, where the merge is a torch.distributed all-reduce.

select / expand / backprop and UCB1 stay on the host exactly as in the reference; `rollout`
is injectable so the host logic can be tested without a device.
"""
import math

import numpy as np

MAXIMIZER, MINIMIZER = 1, -1
EPS = 5e-324   # math.SmallestNonzeroFloat64


class Node:                      # mcts.go:15-23
    __slots__ = ("move", "parent", "children", "wins", "visits", "untried", "pjm")

    def __init__(self, move, parent, pjm):
        self.move, self.parent, self.pjm = move, parent, pjm
        self.children, self.untried = [], None        # None = waiting for its first batch
        self.wins = self.visits = 0.0

    def ucb1(self, uctk, factor2):                    # mcts.go:248-250 / squavam.go:197-199
        f = 2.0 if factor2 else 1.0
        return self.wins / (self.visits + EPS) + uctk * math.sqrt(f * math.log(self.parent.visits) / (self.visits + EPS))

    def best_move(self, uctk, factor2):               # mcts.go:218-238
        best, score = None, EPS
        for c in self.children:
            u = c.ucb1(uctk, factor2)
            if u > score:
                best, score = c, u
        return best


def gpu_rollout(sq):
    """rollout(boards[(x,o)], to_move, playouts_per_leaf, stream_id) -> u64[n,4] via sq_playouts."""
    def f(boards, to_move, ppl, stream_id):
        return sq.playouts(np.array(boards, dtype=np.uint32), np.array(to_move, dtype=np.int8), ppl, stream_id)
    return f


def uct(x, o, player_just_moved, itermax, uctk, rollout, rng, batch=256, ppl=1, factor2=True,
        root=None, stream_base=0):
    """-> root node after itermax rollouts.  Leaf-parallel: `batch` leaves per rollout call."""
    empties = lambda bx, bo: [c for c in range(25) if not ((bx | bo) >> c) & 1]
    if root is None:
        root = Node(-1, None, player_just_moved)
        root.untried = empties(x, o)
    done, call = 0, 0
    while done < itermax:
        p = min(ppl, itermax - done)
        b = max(1, min(batch, (itermax - done) // p))
        picked = []
        for _ in range(b):
            node, bx, bo, pjm = root, x, o, player_just_moved
            while node.untried is not None and not node.untried and node.children:
                node = node.best_move(uctk, factor2)
                pjm = -pjm
                if pjm == MAXIMIZER:
                    bx |= 1 << node.move
                else:
                    bo |= 1 << node.move
            fresh = False
            if node.untried:
                m = node.untried.pop(int(rng.integers(len(node.untried))))
                pjm = -pjm
                if pjm == MAXIMIZER:
                    bx |= 1 << m
                else:
                    bo |= 1 << m
                child = Node(m, node, pjm)
                node.children.append(child)
                node, fresh = child, True
            n = node
            while n is not None:
                n.visits += p
                n = n.parent
            picked.append((node, bx, bo, -pjm, fresh))
        out = rollout([(bx, bo) for _, bx, bo, _, _ in picked], [tm for _, _, _, tm, _ in picked], p,
                      stream_base + call)
        call += 1
        for (node, bx, bo, _, fresh), r in zip(picked, out):
            if fresh:
                node.untried = empties(bx, bo) if int(r[3]) != 0 else []     # zero plies <=> terminal leaf
            n = node
            while n is not None:
                n.wins += float(r[0] if n.pjm == MAXIMIZER else r[1])
                n = n.parent
        done += b * p
    return root


def root_statistics(root):
    visits, wins = np.zeros(25), np.zeros(25)
    for c in root.children:
        visits[c.move], wins[c.move] = c.visits, c.wins
    return visits, wins


def choose_from_statistics(visits, wins, uctk, factor2=True):
    """bestMove over merged root statistics: argmax UCB1, first maximum wins (mcts.go:218-238)."""
    total = visits.sum()
    best, score = -1, EPS
    for m in range(25):
        if visits[m] <= 0:
            continue
        u = wins[m] / (visits[m] + EPS) + uctk * math.sqrt((2.0 if factor2 else 1.0) * math.log(total) / (visits[m] + EPS))
        if u > score:
            best, score = m, u
    return best, score
