"""The reference-facing host programs (squava_b200/host, C++ over the C ABI) on a GPU:
output format, flags and values against the oracle / golden vectors."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "squava_b200", "host")
G = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))


def run(prog, *args, stdin=""):
    exe = os.path.join(HOST, prog)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST, "-s"])
    p = subprocess.run([exe, *args], input=stdin, capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout, p.stderr


def test_squavam_one_decision_config1():
    """configs[0]: squavam -C -i 10000 -u 1.0, computer first; EOF on stdin exits 0 (squavam.go:351-353)"""
    rc, out, err = run("squavam", "-C", "-i", "10000", "-u", "1.0", "-seed", "7")
    assert rc == 0, err
    m = re.search(r"My move: (\d) (\d) (\S+)", out)
    assert m and 0 <= int(m.group(1)) <= 4 and 0 <= int(m.group(2)) <= 4
    assert out.startswith("   0 1 2 3 4\n0  _ _ _ _ _ \n")
    assert "Enter move (N M): " in out
    boards = out.split("   0 1 2 3 4\n")
    assert boards[2].count("O") == 1            # with -C the computer's marks print as O (squavam.go:52-54)


def test_squavam_plays_a_game_against_scripted_moves():
    moves = "\n".join("%d %d" % (x, y) for x in range(5) for y in range(5)) + "\n"
    rc, out, err = run("squavam", "-i", "3000", "-batch", "64", "-seed", "3", stdin=moves)
    assert rc == 0, err
    assert re.search(r"Player [XO] (wins|loses)!", out) or "Enter move" in out


def test_opening2_matches_oracle():
    rc, out, err = run("opening2", "-d", "8")
    assert rc == 0, err
    rows = re.findall(r"<(\d),(\d)>\t(-?\d+) \((-?\d+)\) \[(\d+)\]", out)
    want = [c for c in G["opening2"] if c["depth"] == 8]
    assert [(5 * int(a) + int(b), int(v)) for a, b, v, _, _ in rows] == [(c["first"], c["value"]) for c in want]


def test_opening3_matches_oracle():
    rc, out, err = run("opening3", "-d", "7")
    assert rc == 0, err
    rows = re.findall(r"<(\d),(\d)>:<(\d),(\d)>\t(-?\d+) \((-?\d+)\) \[(\d+)\]", out)
    want = [c for c in G["opening3"] if c["depth"] == 7]
    assert len(rows) == 85
    assert [(5 * int(a) + int(b), 5 * int(p) + int(q), int(v)) for a, b, p, q, v, _, _ in rows] == \
        [(c["first"], c["reply"], c["value"]) for c in want]
    assert re.findall(r"Unique moves computed: (\d+)", out) == ["14", "24", "14", "14", "14", "5"]


def test_squavathr_game_every_move_checked(oracle, sq):
    """squavathr -D driven move by move over its stdin/stdout protocol by a careful human (the CPU oracle's own
    depth-4 choice from the human's side, so the game lasts into the endgame).  EVERY computer move is checked: the printed
    value must be the maximum of chooseMove's root values and the move the first cell of movekeeper's tie set
    (squavathr.go:236-269, movekeeper.go:26-52) for the position it was chosen in, at the depth setDepth gives
    (:183-198: 10 / 12 / -d).  From move 11 on (depth = -d 15, more than the empty cells left: the search runs
    into full boards) the row comes from the CPU oracle; the early, deep searches are far beyond the oracle's
    reach in a test and are compared with sq_alphabeta_root called directly -- the search itself is oracle-checked
    at those depths in test_gpu_alphabeta.py and test_gpu_configs.py."""
    import numpy as np
    from squava_b200 import positions
    exe = os.path.join(HOST, "squavathr")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST, "-s"])
    p = subprocess.Popen([exe, "-D", "-d", "15"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    sc = [3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3]     # squavathr.go:620-628

    def read_until_prompt():
        buf = ""
        while not buf.endswith("Your move: "):
            ch = p.stdout.read(1)
            if ch == "":
                break
            buf += ch
        return buf
    x = o = 0                      # x: the computer (MAXIMIZER), o: the human
    counter, late, moves_checked = 0, 0, 0
    order = [12, 0, 24, 6, 18, 4, 20, 2, 22, 10, 14, 8, 16, 1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23]
    sq.init()
    try:
        text = read_until_prompt()
        while text.endswith("Your move: "):
            empties = [c for c in order if not (x | o) >> c & 1]
            safe = [c for c in empties if not positions.has_line(o | 1 << c)]
            _, hbm, _, _ = oracle.ab2_root(o, x, 4, sc)                  # the human's view: its own marks are the maximizer's
            best = [c for c in (safe or empties) if hbm >> c & 1]
            h = (best or safe or empties)[0]
            o |= 1 << h; counter += 1
            p.stdin.write("%d %d\n" % (h // 5, h % 5)); p.stdin.flush()
            text = read_until_prompt()
            m = re.search(r"My move: (\d) (\d) \((-?\d+)\) \[(\d+)\]", text)
            if not m:
                break                                   # the human's move ended the game
            depth = 10 if counter < 4 else 12 if counter <= 10 else 15
            if counter > 10:
                vals, bm, bv, _ = oracle.ab2_root(x, o, depth, sc)
                late += 1
            else:
                v, bvv, bmm, _ = sq.alphabeta_root(np.array([(x, o)], dtype=sq.BOARD), 2, depth, sc)
                vals, bm, bv = v[0].tolist(), int(bmm[0]), int(bvv[0])
            cell = 5 * int(m.group(1)) + int(m.group(2))
            assert int(m.group(3)) == bv and cell == min(c for c in range(25) if bm >> c & 1), (counter, m.group(0), vals)
            assert int(m.group(4)) > 0
            x |= 1 << cell; counter += 1; moves_checked += 1
        rest = text + p.stdout.read()
        assert p.wait(timeout=60) == 0, p.stderr.read()
        assert moves_checked >= 5, moves_checked
        assert late >= 1 or counter <= 10, (moves_checked, late, counter)       # a game that reaches move 11 has oracle-checked moves
        assert re.search(r"(X wins|O wins|Cat wins)", rest)
    finally:
        sq.shutdown()
        if p.poll() is None:
            p.kill()


def test_playoff5_game_ends_consistently(oracle):
    rc, out, err = run("playoff5", "-1", "M", "-2", "M", "-i1", "20000", "-i2", "20000", "-seed", "5")
    assert rc == 0, err
    assert re.search(r"(player 1 X \(MCTS\) wins|player 2 O \(MCTS\) wins|Cat wins)", out)
    assert "Winner disagreement" not in out


def test_playoff5_batch_mode_line_format():
    """playoff5 -n 2: one line per game, "<i> <first> <second> <depth> <randomize> <moves> <winner> x,y ..." (playoff5.go:135,166-181)"""
    rc, out, err = run("playoff5", "-n", "2", "-1", "A", "-2", "A", "-seed", "9")
    assert rc == 0, err
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 2
    for i, l in enumerate(lines):
        m = re.match(r"^%d AlphaBeta AlphaBeta 10 false (\d+) (-?[01])((?: \d[+-]?,\d[+-]?)+)$" % i, l)
        assert m, l
        assert len(m.group(3).split()) == int(m.group(1))


def test_squavathr_book_transcripts_match_go():
    """-B: the "My move" lines of bookStart / bookDefend (squavathr.go:662-810) against the
    reference's own functions (golden made by make_ref_golden_abbook.py).  stdin ends where the
    book asks for the next human move, so no search starts (EOF exits 0, squavathr.go:566-568)."""
    V = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_go_abbook_vectors.json")))["squavathr_book"]
    n = 0
    for c in V:
        used = len(c["replies"]) - c["unread"]
        want = c["said"]
        if c["role"] == "start":
            # the book returned: the search (depth 12 from six marks) would be next; stop one read earlier
            replies = c["replies"][:used] if c["count"] < 0 else c["replies"][:used - 1]
            env = dict(os.environ, SQUAVA_BOOK_PICK=str(c["pick"]))
            args = ["-B", "-C", "-D"]
        else:
            replies = [c["first"]] + c["replies"][:used]
            if (1 + c["count"]) % 2 == 1:      # the computer would search next: stop at the book's only read
                replies, want = [c["first"]], c["said"][:1]
            env = dict(os.environ)
            args = ["-B", "-D"]
        stdin = "".join("%d %d\n" % (r // 5, r % 5) for r in replies)
        exe = os.path.join(HOST, "squavathr")
        p = subprocess.run([exe, *args], input=stdin, capture_output=True, text=True, timeout=120, env=env)
        assert p.returncode == 0, p.stderr
        assert p.stdout.startswith("Using opening book\n")
        said = [[int(a), int(b)] for a, b in re.findall(r"My move: (-?\d) (-?\d)\n", p.stdout)]
        assert said == want, (c, p.stdout)
        n += 1
    assert n > 40


def test_playoff5_negascout_vs_book_game_replays_on_the_oracle(oracle):
    """playoff5 -1 N -2 B -D: every searched move, its value and its leaf count must be what the CPU
    oracle gets for the same position at the depth SetDepth gives (negascout.go:69-79, abbook.go:80-90)"""
    rc, out, err = run("playoff5", "-1", "N", "-2", "B", "-D")
    assert rc == 0, err
    sc = oracle.tables()["scores_ab"]
    lines = re.findall(r"^([XO]) \((NegaScout|A/B\+Book)\) <(\d),(\d)> \((-?\d+)\) \[(\d+)\]", out, flags=re.M)
    assert len(lines) >= 6 and lines[0][1] == "NegaScout" and lines[1][1] == "A/B+Book"
    mine = {"X": [0, 0], "O": [0, 0]}            # [own marks, opponent marks] as each engine sees the board
    searched = 0
    for counter, (who, name, a, b, value, leaves) in enumerate(lines):
        cell, value, leaves = 5 * int(a) + int(b), int(value), int(leaves)
        own, opp = mine[who]
        if name == "NegaScout":
            depth = 6 if counter < 4 else 8 if counter <= 10 else 10
            vals, order, bm, bv, fb, lv = oracle.negascout_root(own, opp, depth, sc)
            assert (cell, value, leaves) == (fb, bv, lv), (counter, lines[counter])
            searched += 1
        elif leaves:                                # the book is over: src/abbook's search
            depth = 6 if counter < 4 else 8 if counter <= 10 else 10
            vals, bm, bv, lv = oracle.ab6_root(own, opp, depth, sc)
            assert value == bv and cell == min(k for k in range(25) if bm >> k & 1), (counter, lines[counter])
            searched += 1
        other = "O" if who == "X" else "X"
        mine[who][0] |= 1 << cell
        mine[other][1] |= 1 << cell
    assert searched >= 5
    assert re.search(r"(player 1 X \(NegaScout\) wins|player 2 O \(A/B\+Book\) wins|Cat wins)", out)


@pytest.mark.parametrize("pair", [("A", "G"), ("N", "B"), ("B", "N")])
def test_playoff5_lockstep_equals_sequential(pair):
    """-n with -r: createPlayers receives `randomize` as `deterministic` (playoff5.go:133), so the games are
    deterministic; advancing them in lockstep (one batched call per ply) must print the same lines"""
    a = run("playoff5", "-n", "3", "-1", pair[0], "-2", pair[1], "-r", "-seed", "3")
    b = run("playoff5", "-n", "3", "-1", pair[0], "-2", pair[1], "-r", "-seed", "3", "-lockstep")
    assert a[0] == 0 and b[0] == 0, (a[2], b[2])
    assert a[1] == b[1] and len(a[1].splitlines()) == 3
    c = run("playoff5", "-n", "3", "-1", pair[0], "-2", pair[1], "-r", "-seed", "3", "-lockstep", "-records")
    recs = [l.split("\t") for l in c[1].splitlines()]
    for line, rec in zip(a[1].splitlines(), recs):
        f = line.split(" ")
        count, winner = int(f[5]), int(f[6])
        assert rec[0] == {1: "X", -1: "O", 0: "C"}[winner] and int(rec[1]) == count
        assert rec[2:] == [m.replace("+", "").replace("-", "") for m in f[7:]]


def test_recreate_replays_a_playoff5_line(tmp_path):
    """recreate.go:10-96: non-move fields are reported and skipped but still counted, markers dropped"""
    p = tmp_path / "game.txt"
    p.write_text("3 A/B X 10 false 5 1 1,1 0+,0 2,2- 3,3 4,4\n")
    rc, out, err = run("recreate", str(p))
    assert rc == 0
    assert [l for l in out.splitlines() if "move" in l] == ["X move 1,1", "O move 0,0", "X move 2,2", "O move 3,3", "X move 4,4"]
    assert err.count("problem") == 7 and 'Move 2, "A/B", problem' in err
    assert out.rstrip().endswith("O _ _ _ _ \n_ X _ _ _ \n_ _ X _ _ \n_ _ _ O _ \n_ _ _ _ X")


def test_opening3_book_table_and_squavathr_book(tmp_path, oracle):
    """opening3 -book writes a value for all 600 (first, reply) pairs by carrying the 85 computed
    pairs through the program's own 8 symmetries; squavathr -book plays its first move from it"""
    book = tmp_path / "book.txt"
    rc, out, err = run("opening3", "-d", "4", "-book", str(book))
    assert rc == 0, err
    rows = [tuple(int(v) for v in l.split()) for l in book.read_text().splitlines()]
    assert len(rows) == 600 and all(r[5] == 4 for r in rows)
    table = {(5 * a + b, 5 * c + d): v for a, b, c, d, v, _ in rows}
    # the computed pairs carry the values the program printed (and the oracle's)
    sc = oracle.tables()["scores_opening"]
    printed = re.findall(r"<(\d),(\d)>:<(\d),(\d)>\t(-?\d+) ", out)
    assert len(printed) == 85
    for a, b, c, d, v in printed:
        f, r = 5 * int(a) + int(b), 5 * int(c) + int(d)
        assert table[(f, r)] == int(v) == oracle.ab_subtree(3, 1 << f, 1 << r, r, 2, 1, sc[f], 4, sc)[0]
    # closed under the symmetries: transpose and 180-degree rotation
    for (f, r), v in table.items():
        tf, tr = 5 * (f % 5) + f // 5, 5 * (r % 5) + r // 5
        assert table[(tf, tr)] == v and table[(24 - f, 24 - r)] == v
    # computer first: the first move whose worst reply is best
    best = max(min(v for (f, r), v in table.items() if f == ff) for ff in range(25))
    firsts = [ff for ff in range(25) if min(v for (f, r), v in table.items() if f == ff) == best]
    rc, out, err = run("squavathr", "-C", "-D", "-book", str(book))
    m = re.search(r"My move: (\d) (\d) \((-?\d+)\) \[book\]", out)
    assert rc == 0 and m and 5 * int(m.group(1)) + int(m.group(2)) == firsts[0] and int(m.group(3)) == best
    # human first at 1,1: the reply that is worst for the first mover
    rc, out, err = run("squavathr", "-D", "-book", str(book), stdin="1 1\n")
    m = re.search(r"My move: (\d) (\d) \((-?\d+)\) \[book\]", out)
    worst = min(v for (f, r), v in table.items() if f == 6)
    assert rc == 0 and m and int(m.group(3)) == -worst and table[(6, 5 * int(m.group(1)) + int(m.group(2)))] == worst


def test_playoff5_lockstep_with_an_mcts_player():
    """MCTS players keep their own trees and are stepped game by game inside the lockstep loop"""
    rc, out, err = run("playoff5", "-n", "2", "-1", "M", "-2", "A", "-lockstep", "-seed", "2")
    assert rc == 0, err
    lines = out.splitlines()
    assert len(lines) == 2
    for g, line in enumerate(lines):
        f = line.split(" ")
        assert f[:3] == [str(g), "MCTS", "AlphaBeta"] and int(f[5]) == len(f) - 7 and f[6] in ("1", "-1", "0")


def test_squavam_fast_arena_tree_decision():
    """squavam -fast: one decision through the arena tree on all host threads, real GPU rollouts"""
    rc, out, err = run("squavam", "-C", "-i", "300000", "-u", "1.0", "-batch", "8192", "-fast", "-stats", "-seed", "4")
    assert rc == 0, err
    m = re.search(r"My move: (\d) (\d) (\S+)", out)
    s = re.search(r"fast tree: (\d+) iterations, (\d+) batches, (\d+) node slots, (\d+) threads", out)
    assert m and s and int(s.group(1)) == 300000 and int(s.group(2)) == 37 and int(s.group(3)) > 300000
    assert 0 <= int(m.group(1)) <= 4 and 0 <= int(m.group(2)) <= 4


def test_squavam_root_parallel_decision():
    """squavam -fast -root-parallel: independent arena trees (one per device named by -gpus; here one), statistics
    merged by move, every iteration counted once"""
    rc, out, err = run("squavam", "-C", "-i", "4000000", "-u", "1.0", "-batch", "1024", "-ppl", "64", "-fast", "-root-parallel",
                       "-stats", "-seed", "4")
    assert rc == 0, err
    m = re.search(r"My move: (\d) (\d) (\S+)", out)
    s = re.search(r"fast tree: (\d+) iterations, (\d+) batches", out)
    assert m and s and int(s.group(1)) == 4000000
    assert len(re.findall(r"^searcher \d+: \d+ iterations", out, flags=re.M)) >= 1


def test_squavam_gpu_tree_decision():
    """squavam -gpu-tree: one decision with the whole tree in GPU memory (sq_mcts_decide); the first batches grow
    256, 512, ... 32768, then 65536 per batch: 8 + 45 batches for 3e6 iterations"""
    rc, out, err = run("squavam", "-C", "-i", "3000000", "-u", "1.0", "-gpu-tree", "-stats", "-seed", "4")
    assert rc == 0, err
    assert re.search(r"My move: [0-4] [0-4] ", out)
    s = re.search(r"gpu tree: (\d+) iterations, (\d+) batches, (\d+) node slots of (\d+);", out)
    assert s and int(s.group(1)) == 3000000 and int(s.group(2)) == 53 and int(s.group(3)) > 1000000
