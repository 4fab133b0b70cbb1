"""Pure-Python emulation of the playout kernel's per-item random stream and ply
loop (squava_b200/csrc/kernels.cuh), for BIT-EXACT tests of small cases.

TEST INFRASTRUCTURE.  The terminal test here walks explicit line lists (it does
not reuse the kernel's shift tricks), so agreement checks the kernel's logic and
not just its arithmetic.
"""
M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
ITEM_QUOTA = 1024


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return [c0, c1, c2, c3]


def _lines(length):
    out = []
    for x in range(5):
        for y in range(5):
            for dx, dy in ((1, 0), (0, 1), (1, 1), (-1, 1)):
                cells = [(x + i * dx, y + i * dy) for i in range(length)]
                if all(0 <= a < 5 and 0 <= b < 5 for a, b in cells):
                    out.append(sum(1 << (5 * a + b) for a, b in cells))
    return out


TRIPLETS, QUADS = _lines(3), _lines(4)


def owns(mask, lines):
    return any((mask & l) == l for l in lines)


def item_split(per_leaf):
    s = (per_leaf + ITEM_QUOTA - 1) // ITEM_QUOTA
    return s, per_leaf // s, per_leaf % s


def emulate_item(x, o, to_move, quota, leaf_index, j, seed, stream_id):
    """-> [max_wins, min_wins, draws, plies] of item j of a NON-terminal leaf."""
    key = ((seed & MASK) ^ (stream_id >> 32), (seed >> 32) & MASK)
    arr = [c for c in range(25) if not ((x | o) >> c) & 1]
    n0 = len(arr)
    out = [0, 0, 0, 0]
    a, b = (x, o) if to_move > 0 else (o, x)      # a: mover's cells
    mover_is_first = True
    n = n0
    blk = 0
    rem = quota
    while rem:
        w = philox4x32_10([blk, j, leaf_index & MASK, stream_id & MASK], key)
        blk += 1
        for word in w:
            r = word
            for _ in range(3):
                if rem == 0:
                    continue
                prod = r * n
                idx, r = prod >> 32, prod & MASK
                cell = arr[idx]
                arr[idx], arr[n - 1] = arr[n - 1], cell
                m = a | (1 << cell)
                out[3] += 1
                three, four = owns(m, TRIPLETS), owns(m, QUADS)
                if three or n - 1 == 0:
                    if four or three:
                        mover_wins = four
                        first_wins = mover_wins == mover_is_first
                        x_wins = first_wins == (to_move > 0)
                        out[0 if x_wins else 1] += 1
                    else:
                        out[2] += 1
                    rem -= 1
                    a, b = (x, o) if to_move > 0 else (o, x)
                    n = n0
                    mover_is_first = True
                else:
                    a, b = b, m
                    n -= 1
                    mover_is_first = not mover_is_first
    return out


def emulate_leaf(x, o, to_move, per_leaf, leaf_index, seed, stream_id):
    s, q, extra = item_split(per_leaf)
    tot = [0, 0, 0, 0]
    for j in range(s):
        r = emulate_item(x, o, to_move, q + (1 if j < extra else 0), leaf_index, j, seed, stream_id)
        tot = [u + v for u, v in zip(tot, r)]
    return tot
