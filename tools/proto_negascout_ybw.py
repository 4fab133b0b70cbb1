"""Prototype (CPU, uses the oracle -- a design study, not product code): exact speculative splitting of
src/negascout's search, young-brothers-wait style.  The first child of a node is searched alone, its siblings as one
parallel wave with the window they would get if nothing changed; a result is used only if the sequential algorithm
would have made exactly that call, so values, visit order and leaf counts stay bit-exact (checked below).  Reported per
split depth S: wasted nodes and the critical path (sum over waves of the largest unit) relative to the sequential
node count.  Finding: exact, but the waves contain searches the sequential order never makes (after a
cut-off, or with a window that later moved) and these are often LARGER than anything it does make: critical path
anywhere from 0.02x to 1.6x sequential (typically 0.2-0.9x) with up to 8x wasted nodes on mid-game positions --
too erratic to build a kernel on.

usage: proto_negascout_ybw.py [seed]"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import pyoracle as P
import oracle_lib
from squava_b200 import positions
import numpy as np
import ctypes as C

WIN, LOSS = 10000, -10000
o = oracle_lib.load()

# fast unit search through the C oracle: need a C entry for negaScout from arbitrary node. Use python for now.
def unit(bd, ply, player, alpha, beta, D, sc, orders):
    cnt = [0]
    v = P.ns_search(bd, ply, player, alpha, beta, D, sc, orders, cnt)
    return v, cnt[0]

class Stats:
    def __init__(self): self.accepted = 0; self.wasted = 0; self.waves = []

def node(bd, ply, player, alpha, beta, D, sc, orders, S, st, root=False, record=None):
    """host-level node; returns (value, critical_path_nodes).  Children at level S are units."""
    cp = 0
    if not root:
        st.accepted += 1
        stop, bv = P.ns_static(bd, ply, D, sc)
        if stop:
            return player * bv, 1
        cp = 1
    score, n = 3 * LOSS, beta
    kids = [c for c in orders[(player + 1) // 2] if bd[c] == 0]
    def search_child(c, a, b):
        bd[c] = player
        if ply + 1 >= S or ply + 1 > D - 1:
            v, k = unit(bd, ply + 1, -player, a, b, D, sc, orders)
            r = (v, k, k)          # value, nodes, critical path
        else:
            sub = Stats()
            v, cpath = node(bd, ply + 1, -player, a, b, D, sc, orders, S, sub)
            r = (v, sub.accepted, cpath, sub.wasted)
        bd[c] = 0
        return r
    i = 0
    while i < len(kids):
        if i == 0:
            r = search_child(kids[0], -n, -alpha)
            st.accepted += r[1]; cp += r[2]
            if len(r) > 3: st.wasted += r[3]
            cur = -r[0]
            if cur > score:
                score = cur        # n == beta
            alpha = max(alpha, score)
            if record is not None: record.append((kids[0], alpha))
            if alpha >= beta: return score, cp
            n = alpha + 1
            i = 1
            continue
        # wave: all remaining children with (-n, -alpha)
        wave = [search_child(c, -n, -alpha) for c in kids[i:]]
        cp += max(w[2] for w in wave)
        restart = False
        for j, w in enumerate(wave):
            c = kids[i + j]
            st.accepted += w[1]
            if len(w) > 3: st.wasted += w[3]
            cur = -w[0]
            if cur > score:
                if n == beta or (not root and ply == D - 2):
                    score = cur
                else:
                    r = search_child(c, -beta, -cur)
                    st.accepted += r[1]; cp += r[2]
                    if len(r) > 3: st.wasted += r[3]
                    score = -r[0]
            old_alpha = alpha
            alpha = max(alpha, score)
            if record is not None: record.append((c, alpha))
            if alpha >= beta:
                st.wasted += sum(x[1] for x in wave[j + 1:])
                return score, cp
            n = alpha + 1
            if alpha != old_alpha:
                st.wasted += sum(x[1] for x in wave[j + 1:])
                i = i + j + 1
                restart = True
                break
        if not restart:
            break
    return score, cp

g = np.random.Generator(np.random.PCG64(int(sys.argv[1]) if len(sys.argv) > 1 else 1))
sc = o.tables()["scores_ab"]
for marks, D in [(12, 6), (10, 6), (14, 8), (8, 4), (10, 5)]:
    x, oo = positions.random_position(g, marks)
    bd = [1 if x >> c & 1 else -1 if oo >> c & 1 else 0 for c in range(25)]
    vals, order, bm, bv, fb, leaves = o.negascout_root(x, oo, D, sc)
    orders = P.ns_reorder(bd)
    for S in (1, 2, 3):
        st = Stats(); rec = []
        t = time.time()
        v, cp = node(list(bd), 0, 1, 2 * LOSS, 2 * WIN, D, sc, orders, S, st, root=True, record=rec)
        ok = rec == [(c, vals[c]) for c in order] and st.accepted == leaves
        print(marks, D, "S", S, "exact", ok, "nodes", leaves, "wasted %.2f" % (st.wasted / leaves), "critical path %.3f" % (cp / leaves), "speedup %.1f" % (leaves / cp), "%.1fs" % (time.time() - t))
