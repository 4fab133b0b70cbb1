"""Host-side logic that needs no device: input generators, the playout emulator's RNG,
the item split, and that the emulator reproduces exact playout statistics."""
import numpy as np

import playout_emulator as E


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert E.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert E.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert E.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                           [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_item_split_covers_every_playout():
    for n in (1, 2, 1023, 1024, 1025, 4096, 10 ** 6, 10 ** 6 + 7):
        s, q, extra = E.item_split(n)
        assert s * q + extra == n and q + (1 if extra else 0) <= E.ITEM_QUOTA and q >= 1


def test_midgame_batch_is_reachable_and_nonterminal(oracle):
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(200, seed=11)
    for b, t in zip(boards, tm):
        x, o = int(b["x"]), int(b["o"])
        nx, no = bin(x).count("1"), bin(o).count("1")
        assert x & o == 0 and 8 <= nx + no <= 14 and nx - no in (0, 1)
        assert t == (1 if nx == no else -1)
        assert oracle.flags(x, o) == 0
    b2, _ = positions.midgame_batch(200, seed=11)
    assert (b2 == boards).all()


def test_emulator_statistics_match_exact_distribution(oracle):
    x, o = 0x0412825, 0x098c0c0
    p, plies = oracle.playout_exact(x, o, 1)
    n = 20000
    out = E.emulate_leaf(x, o, 1, n, 3, 99, 5)
    assert sum(out[:3]) == n
    for k in range(3):
        sigma = (p[k] * (1 - p[k]) / n) ** 0.5
        assert abs(out[k] / n - p[k]) <= 4 * sigma + 1e-12
    assert abs(out[3] / n - plies) < 0.05


def test_recreate_program_follows_the_reference(tmp_path):
    """recreate.go:10-96 (no GPU involved): one move per line when the file has several lines, else blank-separated;
    a field that is not a move is reported and skipped BUT still counted, so it flips the X/O alternation
    (markers[moveCounter%2], moveCounter starting at 1: the first move is drawn as O); "+"/"-" value markers
    after a digit are dropped; at most 25 fields are looked at"""
    import os
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "squava_b200", "host")
    exe = os.path.join(host, "recreate")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", host, "-s", "recreate"])
    p = tmp_path / "lines.txt"
    p.write_text("1,1\n0+,0\n2,2-\nbad\n3,3\n")
    r = subprocess.run([exe, str(p)], input="\n" * 10, capture_output=True, text=True, timeout=30)
    assert r.returncode == 0
    assert [l for l in r.stdout.splitlines() if "move" in l] == ["O move 1,1", "X move 0,0", "O move 2,2", "O move 3,3"]
    assert 'Move 4, "bad", problem' in r.stderr and "Move 6" in r.stderr      # the empty field after the last newline
    q = tmp_path / "record.txt"
    q.write_text("X\t9\t" + "\t".join("%d,%d" % (i // 5, i % 5) for i in range(9)) + "\n")
    r = subprocess.run([exe, str(q)], input="\n" * 30, capture_output=True, text=True, timeout=30)
    # a tab-separated playoff4 record is ONE blank-separated field to recreate.go: it cannot replay it as is
    assert r.returncode == 0 and "move" not in r.stdout and "problem" in r.stderr
    s = tmp_path / "many.txt"
    s.write_text(" ".join("%d,%d" % (i // 5, i % 5) for i in range(25)) + " 0,0 1,1")
    r = subprocess.run([exe, str(s)], input="\n" * 40, capture_output=True, text=True, timeout=30)
    assert len([l for l in r.stdout.splitlines() if "move" in l]) == 25
