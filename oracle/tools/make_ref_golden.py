#!/usr/bin/env python3
"""Golden vectors produced BY THE REFERENCE'S OWN CODE: squava2.py, the Python-2 UCT prototype in
the reference tree (same board, same line tables and the same GetMoves / GetResult / rollout loop
that src/mcts and squavam.go were transliterated from -- README.md:89-91).

The file is Python 2; it is loaded from /root/reference at run time, its print statements are
rewritten in memory (nothing is copied into this repo), and its SquavaState class is driven
directly.  Output: tests/golden/ref_squava2_vectors.json
  boards    random reachable states (every prefix of random games, terminal ones included) and random
            ternary boards, each with what the reference says: the move list GetMoves() returns
            (empty <=> terminal) and GetResult(+1) / GetResult(-1) (null where the reference itself
            asserts: a full board without a line)
  rollouts  for fixed positions, the outcome counts and total plies of the reference's own rollout
            loop (squava2.py:238-240) under a seeded `random`

TEST INFRASTRUCTURE; run once in the build container (the reference does not travel).
"""
import io
import json
import os
import random
import re
import sys
import contextlib

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "ref_squava2_vectors.json")


def load_reference():
    src = open(os.path.join(REF, "squava2.py")).read()
    src = re.sub(r"^(\s*(?:else:\s*)?)print (.*)$", r"\1print(\2)", src, flags=re.M)   # py2 print statements
    src = src.replace("raw_input", "input")
    ns = {"__name__": "squava2_reference"}
    exec(compile(src, "squava2.py (reference, print statements rewritten in memory)", "exec"), ns)
    return ns


def masks(board):
    x = sum(1 << i for i in range(25) if board[i] == 1)
    o = sum(1 << i for i in range(25) if board[i] == -1)
    return x, o


def describe(state):
    moves = state.GetMoves()
    rec = {"x": masks(state.board)[0], "o": masks(state.board)[1], "moves": moves}
    if moves == []:
        res = []
        for p in (1, -1):
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    res.append(state.GetResult(p))
            except AssertionError:
                res.append(None)           # "This should not happen": full board, no line
        rec["result_x"], rec["result_o"] = res
    return rec


def main():
    ns = load_reference()
    S = ns["SquavaState"]
    rng = random.Random(20261020)
    boards = []
    # (i) every state of random games
    for g in range(220):
        st = S()
        boards.append(describe(st))
        while st.GetMoves() != []:
            st.DoMove(rng.choice(st.GetMoves()))
            boards.append(describe(st))
    # (ii) random ternary boards, mostly unreachable
    for g in range(1500):
        st = S()
        fill = rng.random()
        st.board = [rng.choice((1, -1)) if rng.random() < fill else 0 for _ in range(25)]
        boards.append(describe(st))
    # (iii) the four drawn full boards and near-full boards
    for xm in (0x1364d93, 0x1552ab5, 0x15aa955, 0x19364d9):
        st = S()
        st.board = [1 if (xm >> i) & 1 else -1 for i in range(25)]
        boards.append(describe(st))
        for hole in range(0, 25, 6):
            st2 = st.Clone()
            st2.board[hole] = 0
            boards.append(describe(st2))

    rollouts = []
    positions = [(0x1080038, 0x0c00181, 1), (0x0504208, 0x00280a1, 1), (0x0c0090a, 0x02400d4, 1),
                 (0x1408051, 0x0a21500, 1), (0x05c2408, 0x1800274, 1), (0x0412825, 0x098c0c0, 1),
                 (0x0000421 ^ 0x0000020, 0x0001040, 1), (0x0100844, 0x0011081, -1), (0, 0, 1)]
    for i, (xm, om, to_move) in enumerate(positions):
        n = 100000 if (xm | om) else 40000
        random.seed(777 + i)                      # the reference's rollout uses the module-level random
        root = S()
        root.board = [1 if (xm >> c) & 1 else -1 if (om >> c) & 1 else 0 for c in range(25)]
        root.playerJustMoved = -to_move
        if root.GetMoves() == []:
            continue
        xw = ow = plies = 0
        for _ in range(n):
            state = root.Clone()
            while state.GetMoves() != []:                                   # squava2.py:238-240
                state.DoMove(random.choice(state.GetMoves()))
                plies += 1
            with contextlib.redirect_stdout(io.StringIO()):
                try:
                    r = state.GetResult(1)
                except AssertionError:
                    r = None
            if r == 1.0:
                xw += 1
            elif r == 0.0:
                ow += 1
        rollouts.append({"x": xm, "o": om, "to_move": to_move, "n": n, "x_wins": xw, "o_wins": ow,
                         "draws": n - xw - ow, "plies": plies})
        print("rollouts", i, rollouts[-1], flush=True)
    with open(OUT, "w") as f:
        json.dump({"source": "squava2.py of the reference, executed by oracle/tools/make_ref_golden.py",
                   "boards": boards, "rollouts": rollouts}, f, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;", len(boards), "boards")


if __name__ == "__main__":
    main()
