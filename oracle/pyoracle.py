"""Independent pure-Python restatement of the squava hot path.

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by tests/ (and by
oracle/tools/make_golden.py).  It exists so that the C oracle
(oracle/squava_oracle.c) is checked by a second, separately written
restatement: this one takes its line tables from tests/golden/ref_tables.json
(parsed out of the reference's Go sources), the C one builds them by rule.

PARITY: see oracle/squava_oracle.h -- this restatement is checked against the same vectors
produced by the reference's own code (tests/test_reference_vectors.py).

Boards are lists of 25 ints in {-1, 0, +1}; cell c = 5*x + y.
Reference lines cited per function (paths relative to the reference root).
"""
import functools
import json
import os

WIN, LOSS = 10000, -10000
MAX, MIN, UNSET = 1, -1, 0
NO_MOVE = -(2 ** 31)

_G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_tables.json")
with open(_G) as _f:
    T = json.load(_f)

IMPORTANT = T["mcts"]["important_cells"]
MQ = T["mcts"]["quads"]        # [25][..][4]
MT = T["mcts"]["triplets"]     # [25][..][3]
AQ = T["ab"]["quads"]          # [28][4]
AT = T["ab"]["triplets"]       # [48][3]
IQ = [[q for q in AQ if c in q] for c in range(25)]   # alphabeta.go:37-50
IT = [[t for t in AT if c in t] for c in range(25)]
SCORES_AB = T["ab"]["scores"]
SCORES_OPENING = T["op3"]["scores"]
ORDERED = T["op3"]["ordered_cells"]
NO2 = T["thr"]["no2"]
NOMID2 = T["thr"]["no_middle2"]
MATRICES = T["op3"]["matrices"]
FIRST_MOVES = T["op3"]["first_moves"]


def from_masks(x, o):
    return [1 if (x >> c) & 1 else -1 if (o >> c) & 1 else 0 for c in range(25)]


def to_masks(bd):
    x = sum(1 << c for c in range(25) if bd[c] == 1)
    o = sum(1 << c for c in range(25) if bd[c] == -1)
    return x, o


# --- src/mcts/mcts.go -------------------------------------------------------
def get_moves(bd):
    """mcts.go:310-346 -> (moves, terminal)"""
    for m in IMPORTANT:
        if bd[m] != UNSET:
            for q in MQ[m]:
                if abs(sum(bd[i] for i in q)) == 4:
                    return [], True
            for t in MT[m]:
                if abs(sum(bd[i] for i in t)) == 3:
                    return [], True
    moves = [i for i in range(25) if bd[i] == UNSET]
    return moves, not moves


def mcts_find_winner(bd):
    """mcts.go:92-121"""
    for i in IMPORTANT:
        if bd[i] != UNSET:
            for q in MQ[i]:
                s = sum(bd[j] for j in q)
                if s == 4:
                    return MAX
                if s == -4:
                    return MIN
    for i in IMPORTANT:
        if bd[i] != UNSET:
            for t in MT[i]:
                s = sum(bd[j] for j in t)
                if s == 3:
                    return MIN
                if s == -3:
                    return MAX
    return UNSET


def get_result(bd, playerjm):
    """mcts.go:348-388"""
    for i in IMPORTANT:
        if bd[i] != UNSET:
            for q in MQ[i]:
                s = sum(bd[j] for j in q)
                if abs(s) == 4:
                    return 1.0 if s == 4 * playerjm else 0.0
    for i in IMPORTANT:
        if bd[i] != UNSET:
            for t in MT[i]:
                s = sum(bd[j] for j in t)
                if abs(s) == 3:
                    return 0.0 if s == 3 * playerjm else 1.0
    return 0.0


def ab_find_winner(bd):
    """alphabeta.go:355-378"""
    for q in AQ:
        if abs(sum(bd[j] for j in q)) == 4:
            return bd[q[0]]
    for t in AT:
        if abs(sum(bd[j] for j in t)) == 3:
            return -bd[t[0]]
    return 0


def flags(bd):
    f = 0
    for q in AQ:
        s = sum(bd[j] for j in q)
        f |= 1 if s == 4 else 2 if s == -4 else 0
    for t in AT:
        s = sum(bd[j] for j in t)
        f |= 4 if s == 3 else 8 if s == -3 else 0
    if all(v != 0 for v in bd):
        f |= 16
    return f


def playout_exact(bd, to_move):
    """Exact (P(MAX wins), P(MIN wins), P(draw), E[plies]) of mcts.go:173-181's loop."""
    @functools.lru_cache(maxsize=None)
    def go(x, o, mover):
        b = from_masks(x, o)
        moves, term = get_moves(b)
        if term:
            if get_result(b, MAX) == 1.0:
                return (1.0, 0.0, 0.0, 0.0)
            if get_result(b, MIN) == 1.0:
                return (0.0, 1.0, 0.0, 0.0)
            return (0.0, 0.0, 1.0, 0.0)
        acc = [0.0, 0.0, 0.0, 0.0]
        for m in moves:
            r = go(x | (1 << m), o, -mover) if mover == MAX else go(x, o | (1 << m), -mover)
            for k in range(3):
                acc[k] += r[k] / len(moves)
            acc[3] += (r[3] + 1.0) / len(moves)
        return tuple(acc)
    return go(*to_masks(bd), to_move)


# --- evaluators ---------------------------------------------------------------
def _lines(bd, ply, c):
    """alphabeta.go:123-147.  -> (terminal, value)"""
    v = 0
    for q in IQ[c]:
        s = sum(bd[j] for j in q)
        if abs(s) == 4:
            return True, bd[q[0]] * (WIN - ply)
        if abs(s) == 3:
            v += s * 10
    for t in IT[c]:
        s = sum(bd[j] for j in t)
        if abs(s) == 3:
            return True, (s // 3) * (LOSS + ply)
    return False, v


def avoid_terms(bd, c):
    """alphabeta.go:412-439 (deltaValue2) == squavathr.go:343-371"""
    v = 0
    for t in NO2:
        if c in t and abs(sum(bd[j] for j in t)) == 2:
            v += bd[c] * -100
    for q in NOMID2:
        p = bd[c]
        if (c == q[1] and p == bd[q[2]]) or (c == q[2] and p == bd[q[1]]):
            if abs(sum(bd[j] for j in q)) == 2:
                v += p * -100
    return v


def delta1(bd, ply, c, current, max_depth, scores, avoid=False):
    """alphabeta.go:121-163; avoid=True: deltaValue2, alphabeta.go:382-454"""
    term, v = _lines(bd, ply, c)
    if term:
        return True, v
    if avoid:
        v += avoid_terms(bd, c)
    v += bd[c] * scores[c]
    if ply >= max_depth:
        return True, v + current
    return False, v


def ab1(bd, ply, player, alpha, beta, c, bv, max_depth, scores, avoid=False, cumulative=False):
    """alphabeta.go:165-223; cumulative=True: abbook.go:215-275 (boardValue+delta handed down)"""
    value = 2 * LOSS if player == MAX else 2 * WIN
    for i in range(25):
        if bd[i] != UNSET:
            continue
        bd[i] = player
        stop, delta = delta1(bd, ply, c, bv, max_depth, scores, avoid)
        if stop:
            bd[i] = UNSET
            return delta
        n = ab1(bd, ply + 1, -player, alpha, beta, i, bv + delta if cumulative else delta, max_depth, scores,
                avoid, cumulative)
        bd[i] = UNSET
        if player == MAX:
            value = max(value, n)
            alpha = max(alpha, value)
        else:
            value = min(value, n)
            beta = min(beta, value)
        if beta <= alpha:
            return value
    return value


def ab1_root(bd, max_depth, scores, avoid=False, cumulative=False):
    """alphabeta.go:92-117 (cumulative=True: abbook.go:104-123) -> values[25]"""
    bd = list(bd)
    out = [NO_MOVE] * 25
    for c in range(25):
        if bd[c] != UNSET:
            continue
        bd[c] = MAX
        stop, v = delta1(bd, 0, c, 0, max_depth, scores, avoid)
        if not stop:
            v = ab1(bd, 1, MIN, 2 * LOSS, 2 * WIN, c, v, max_depth, scores, avoid, cumulative)
        bd[c] = UNSET
        out[c] = v
    return out


def delta2(bd, ply, c, current, max_depth, scores):
    """squavathr.go:315-387"""
    term, v = _lines(bd, ply, c)
    if term:
        return True, v
    for t in NO2:
        if c in t:
            if abs(sum(bd[j] for j in t)) == 2:
                v += bd[c] * -100
    for q in NOMID2:
        p = bd[c]
        if (c == q[1] and p == bd[q[2]]) or (c == q[2] and p == bd[q[1]]):
            if abs(sum(bd[j] for j in q)) == 2:
                v += p * -100
    v += bd[c] * scores[c]
    if ply == max_depth:
        return True, v + current
    return False, v


def ab2(bd, ply, player, alpha, beta, c, bv, max_depth, scores):
    """squavathr.go:389-446"""
    stop, delta = delta2(bd, ply, c, bv, max_depth, scores)
    if stop:
        return delta
    bv += delta
    value = 2 * LOSS if player == MAX else 2 * WIN
    for i in range(25):
        if bd[i] != UNSET:
            continue
        bd[i] = player
        n = ab2(bd, ply + 1, -player, alpha, beta, i, bv, max_depth, scores)
        bd[i] = UNSET
        if player == MAX:
            value = max(value, n)
            alpha = max(alpha, value)
        else:
            value = min(value, n)
            beta = min(beta, value)
        if beta <= alpha:
            return value
    return value


def ab2_root(bd, max_depth, scores):
    """squavathr.go:236-269 (+ worker :448-466)"""
    bd = list(bd)
    md = max_depth - 1
    out = [NO_MOVE] * 25
    for c in range(25):
        if bd[c] != UNSET:
            continue
        stop, v = delta2(bd, 0, c, 0, md, scores)
        if not stop:
            bd[c] = MAX
            v = ab2(bd, 1, MIN, 2 * LOSS, 2 * WIN, c, v, md, scores)
            bd[c] = UNSET
        out[c] = v
    return out


def ab3(bd, ply, nxt, alpha, beta, c, bv, max_depth, scores, sentinel=2 * WIN):
    """opening3.go:136-225 (sentinel=WIN gives opening2.go:91-180)"""
    term, delta = _lines(bd, ply, c)
    if term:
        return delta
    delta += bd[c] * scores[c]
    bv += delta
    if ply == max_depth:
        return bv
    value = -sentinel if nxt == MAX else sentinel
    for i in ORDERED:
        if bd[i] != UNSET:
            continue
        bd[i] = nxt
        n = ab3(bd, ply + 1, -nxt, alpha, beta, i, bv, max_depth, scores, sentinel)
        bd[i] = UNSET
        if nxt == MAX:
            value = max(value, n)
            alpha = max(alpha, value)
        else:
            value = min(value, n)
            beta = min(beta, value)
        if beta <= alpha:
            break
    return value


def minimax_ab1(bd, ply, player, c, bv, max_depth, scores):
    """Plain minimax (no pruning) of AB-1's tree: SURVEY.md Appendix C's val(N)."""
    value = None
    for i in range(25):
        if bd[i] != UNSET:
            continue
        bd[i] = player
        stop, delta = delta1(bd, ply, c, bv, max_depth, scores)
        if stop:
            bd[i] = UNSET
            return delta
        n = minimax_ab1(bd, ply + 1, -player, i, delta, max_depth, scores)
        bd[i] = UNSET
        value = n if value is None else (max(value, n) if player == MAX else min(value, n))
    if value is None:
        return 2 * LOSS if player == MAX else 2 * WIN
    return value


def best_mask(values):
    """movekeeper.go:26-52 as a set: cells tied at the maximum"""
    live = [v for v in values if v != NO_MOVE]
    if not live:
        return 0, 0
    m = max(live)
    return sum(1 << c for c in range(25) if values[c] == m), m


def opening3_subtrees():
    """opening3.go:66-128 -> [(first, reply)]"""
    computed = [[0] * 25 for _ in range(6)]
    for k, f in enumerate(FIRST_MOVES):
        computed[k][f] = MAX
    out = []
    for k, f in enumerate(FIRST_MOVES):
        for r in range(25):
            if r == f:
                continue
            matched = False
            for m in MATRICES:
                a, c = m[f], m[r]
                if any(b[a] == MAX and b[c] == MIN for b in computed):
                    matched = True
                    break
            if not matched:
                computed[k][r] = MIN
                out.append((f, r))
    return out


# ---- NegaScout, src/negascout/negascout.go -------------------------------------------------
NS_INITIAL_ORDER = [6, 8, 18, 16, 1, 3, 9, 19, 23, 21, 15, 5, 0, 4, 24, 20, 12, 7, 13, 17, 11, 2, 10, 14, 22]  # :392-399
NS_DEADLY = [[5, 11, 17, 23], [21, 17, 13, 9], [1, 7, 13, 19], [15, 11, 7, 3]]                                  # :147-152


def ns_reorder(bd):
    """reorderMoves, negascout.go:411-499 -> (order for MINIMIZER, order for MAXIMIZER)"""
    good, bad, dull = [], [], []
    for c in NS_INITIAL_ORDER:
        if bd[c] != UNSET:
            continue
        kind = None
        for q in IQ[c]:
            s = sum(bd[k] for k in q)
            if s in (3, 2):
                kind = good
            elif s in (-3, -2):
                kind = bad
            if kind is not None:
                break
        if kind is None:
            for t in IT[c]:
                s = sum(bd[k] for k in t)
                if s == -2:
                    kind = good
                elif s == 2:
                    kind = bad
                if kind is not None:
                    break
        (dull if kind is None else kind).append(c)
    return bad + dull + good, good + dull + bad


def ns_static(bd, ply, max_depth, scores):
    """staticValue, negascout.go:168-230 -> (stop, value)"""
    v = 0
    for c in IMPORTANT:
        for q in IQ[c]:
            s = sum(bd[k] for k in q)
            if s in (4, -4):
                return True, bd[q[0]] * (WIN - ply)
            if s in (3, -3):
                v += s * 10
    for c in IMPORTANT:
        for t in IT[c]:
            s = sum(bd[k] for k in t)
            if s in (3, -3):
                return True, -(s // 3) * (WIN - ply)
    for d in NS_DEADLY:
        outer, inner = bd[d[0]] + bd[d[3]], bd[d[1]] + bd[d[2]]
        if inner in (2, -2) and outer == 0:
            v -= inner // 2 * 5
    if v == 0:
        v = sum(bd[c] * scores[c] for c in range(25))
    return ply > max_depth, v


def ns_search(bd, ply, player, alpha, beta, max_depth, scores, orders, count):
    """negaScout, negascout.go:232-266"""
    count[0] += 1
    stop, bv = ns_static(bd, ply, max_depth, scores)
    if stop:
        return player * bv
    score, n = 3 * LOSS, beta
    for c in orders[(player + 1) // 2]:
        if bd[c] != UNSET:
            continue
        bd[c] = player
        cur = -ns_search(bd, ply + 1, -player, -n, -alpha, max_depth, scores, orders, count)
        if cur > score:
            if n == beta or ply == max_depth - 2:
                score = cur
            else:
                score = -ns_search(bd, ply + 1, -player, -beta, -cur, max_depth, scores, orders, count)
        bd[c] = UNSET
        alpha = max(alpha, score)
        if alpha >= beta:
            break
        n = alpha + 1
    return score


def negascout_root(bd, max_depth, scores):
    """ChooseMove, negascout.go:81-122 -> (values[25], visit order, leaves)"""
    bd = list(bd)
    orders = ns_reorder(bd)
    values, visited, count = [NO_MOVE] * 25, [], [0]
    beta, alpha, score = 2 * WIN, 2 * LOSS, 3 * LOSS
    n = beta
    for c in orders[1]:
        if bd[c] != UNSET:
            continue
        bd[c] = MAX
        cur = -ns_search(bd, 1, MIN, -n, -alpha, max_depth, scores, orders, count)
        if cur > score:
            score = cur if n == beta else -ns_search(bd, 1, MIN, -beta, -cur, max_depth, scores, orders, count)
        bd[c] = UNSET
        alpha = max(alpha, score)
        values[c] = alpha
        visited.append(c)
        if alpha >= beta:
            break
        n = alpha + 1
    return values, visited, count[0]
