"""Synthetic position batches for the benchmark configurations.

Input generation only: nothing here is on the measured path.  Positions are
reachable, non-terminal boards obtained by playing k uniformly random legal
moves from the empty board with X moving first, redrawing the whole game when
any prefix is terminal.
"""
import numpy as np

from . import BOARD

SEED_C2 = 0x5155415641

# all 3-lines of the 5x5 board as 25-bit masks (a 4-line contains a 3-line)
_TRIPLETS = []
for _x in range(5):
    for _y in range(5):
        for _dx, _dy in ((1, 0), (0, 1), (1, 1), (-1, 1)):
            _cells = [(_x + i * _dx, _y + i * _dy) for i in range(3)]
            if all(0 <= a < 5 and 0 <= b < 5 for a, b in _cells):
                _TRIPLETS.append(sum(1 << (5 * a + b) for a, b in _cells))


def has_line(mask):
    """True if the mask owns three (hence possibly four) in a row."""
    return any((mask & t) == t for t in _TRIPLETS)


def random_position(rng, k):
    """k random legal moves from the empty board, X first; never terminal on the way."""
    while True:
        cells = rng.permutation(25)[:k]
        x = o = 0
        ok = True
        for i, c in enumerate(cells):
            if i % 2 == 0:
                x |= 1 << int(c)
                ok = not has_line(x)
            else:
                o |= 1 << int(c)
                ok = not has_line(o)
            if not ok:
                break
        if ok:
            return x, o


def midgame_batch(n, marks=(8, 9, 10, 11, 12, 13, 14), seed=SEED_C2):
    """-> (boards BOARD[n], to_move i8[n]).  Config 2: marks 8..14; config 4: (8, 10, 12, 14)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    boards = np.zeros(n, BOARD)
    to_move = np.zeros(n, np.int8)
    marks = list(marks)
    for i in range(n):
        k = marks[int(rng.integers(len(marks)))]
        x, o = random_position(rng, k)
        boards[i] = (x, o)
        to_move[i] = 1 if k % 2 == 0 else -1
    return boards, to_move


def empty_board_batch(n):
    return np.zeros(n, BOARD), np.ones(n, np.int8)
