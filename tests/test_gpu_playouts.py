"""Random playouts on the GPU: bit-exact against the Python emulation of the kernel's
random stream, and within 3 sigma of the EXACT outcome distribution (oracle DP) at 1e7
playouts (BASELINE north_star's stated tolerance)."""
import numpy as np
import pytest

import playout_emulator as E

pytestmark = pytest.mark.gpu
SEED = 0x5155415641


@pytest.fixture(scope="module")
def dev(sq):
    sq.init(seed=SEED)
    yield sq
    sq.shutdown()


def boards_of(pairs, sq):
    return np.array(pairs, dtype=np.uint32).view(sq.BOARD).reshape(-1)


def test_philox_on_device(dev):
    for ctr, key in (([0, 0, 0, 0], [0, 0]), ([0xffffffff] * 4, [0xffffffff] * 2),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])):
        assert [int(v) for v in dev.debug_philox(ctr, key)] == E.philox4x32_10(ctr, key)


def test_bit_exact_vs_emulator(dev):
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(24, seed=3)
    boards = np.concatenate([boards, np.zeros(1, dev.BOARD)])       # + the empty board
    tm = np.concatenate([tm, np.array([-1], np.int8)])
    for per_leaf, stream in ((1, 0), (7, 1), (300, 2), (1500, (5 << 32) | 9)):
        out = dev.playouts(boards, tm, per_leaf, stream)
        assert (out[:, :3].sum(axis=1) == per_leaf).all()
        for i in range(boards.shape[0]):
            want = E.emulate_leaf(int(boards[i]["x"]), int(boards[i]["o"]), int(tm[i]), per_leaf, i, SEED, stream)
            assert [int(v) for v in out[i]] == want, (i, per_leaf)


def test_three_sigma_vs_exact_distribution_at_1e7(dev, oracle):
    from squava_b200 import positions
    # configs[1] draws its leaves from 8..14 marks: two positions of every size, the 17-empty ones included
    parts = [positions.midgame_batch(2, marks=(k,), seed=21 + k) for k in (8, 9, 10, 11, 12, 13, 14)]
    boards = np.concatenate([p[0] for p in parts]); tm = np.concatenate([p[1] for p in parts])
    assert sorted(bin(int(b["x"]) | int(b["o"])).count("1") for b in boards) == [8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14]
    n = 10_000_000
    out = dev.playouts(boards, tm, n, 77)
    for i in range(boards.shape[0]):
        p, plies = oracle.playout_exact(int(boards[i]["x"]), int(boards[i]["o"]), int(tm[i]))
        assert int(out[i, :3].sum()) == n
        for k in range(3):
            tol = 3.0 * (p[k] * (1 - p[k]) / n) ** 0.5      # 3 sigma, binomial
            assert abs(int(out[i, k]) / n - p[k]) <= tol + 1e-9, (i, k, int(out[i, k]) / n, p[k])
        assert abs(int(out[i, 3]) / n - plies) < 3e-3


def test_empty_board_matches_cpu_oracle_sampling(dev, oracle):
    """from the empty board the DP is too large; compare with the oracle's own sampling
    (two-sample 3 sigma: 3*sqrt(2p(1-p)/n))."""
    n = 400_000
    out = dev.playouts(np.zeros(1, dev.BOARD), np.ones(1, np.int8), n, 5)[0]
    ref = oracle.playouts(0, 0, 1, n, 1234)
    assert int(out[2]) <= 3 and ref[2] <= 3       # P(draw) = 4/C(25,13) = 7.7e-7 per playout
    p = ref[0] / n
    assert abs(int(out[0]) / n - p) <= 3.0 * (2 * p * (1 - p) / n) ** 0.5
    assert abs(int(out[3]) / n - ref[3] / n) < 0.05
    assert 11.7 < int(out[3]) / n < 12.0                 # SURVEY: mean 11.87 plies


def test_terminal_and_full_leaves_play_zero_plies(dev, oracle):
    pairs = [(0b1111, 0b1110000000), (0b111, 0b11000000), (0b1100000, 0b111),
             (0x1364d93, 0x1ffffff & ~0x1364d93)]
    b = boards_of(pairs, dev)
    out = dev.playouts(b, np.array([-1, -1, 1, -1], np.int8), 1000, 3)
    for i, (x, o) in enumerate(pairs):
        w = oracle.mcts_find_winner(x, o)
        want = [1000 if w == 1 else 0, 1000 if w == -1 else 0, 1000 if w == 0 else 0, 0]
        assert [int(v) for v in out[i]] == want


def test_one_move_left(dev, oracle):
    # 24 marks, X to move into the last cell
    x, full = 0x1364d93, 0x1ffffff
    for c in range(25):
        if not (x >> c) & 1:
            continue
        xb, ob = x & ~(1 << c), full & ~x
        if oracle.flags(xb, ob):
            continue
        out = dev.playouts(boards_of([(xb, ob)], dev), np.ones(1, np.int8), 50, 0)[0]
        assert [int(v) for v in out] == [0, 0, 50, 50]      # completes the drawn board
        break
    else:
        pytest.fail("no usable cell")


def test_ragged_and_empty_inputs(dev):
    assert dev.playouts(np.zeros(0, dev.BOARD), np.zeros(0, np.int8), 10).shape == (0, 4)
    out = dev.playouts(np.zeros(3, dev.BOARD), np.ones(3, np.int8), 0)
    assert (out == 0).all()
    with pytest.raises(dev.SquavaError):
        dev.playouts(boards_of([(3, 1)], dev), np.ones(1, np.int8), 10)      # overlapping masks
    with pytest.raises(dev.SquavaError):
        dev.playouts(np.zeros(1, dev.BOARD), np.zeros(1, np.int8), 10)       # to_move must be +-1


def test_results_do_not_depend_on_batching(dev):
    """leaf_index keys the stream: a leaf gives the same counts alone or inside a batch
    (what makes the multi-GPU shard merge exact)."""
    import torch
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(64, seed=8)
    full = dev.playouts(boards, tm, 2000, 4)
    tb = torch.from_numpy(boards.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).cuda()
    tt = torch.from_numpy(tm.copy()).cuda()
    out = torch.zeros((64, 4), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream()
    # second half only, as a separate shard with its global base index
    dev.playouts_dev(tb[32:].data_ptr(), tt[32:].data_ptr(), 32, 2000, out[32:].data_ptr(),
                     leaf_index_base=32, stream_id=4, stream=s)
    dev.playouts_dev(tb[:32].data_ptr(), tt[:32].data_ptr(), 32, 2000, out[:32].data_ptr(),
                     leaf_index_base=0, stream_id=4, stream=s)
    torch.cuda.synchronize()
    assert (out.cpu().numpy().astype(np.uint64) == full).all()
    assert dev.last_playout_kernel_ms() > 0


def test_linearity_and_symmetry_at_scale(dev):
    """size-independent properties at a BASELINE-sized call: counts add up, colour swap
    mirrors the outcome, a reflected board has the same distribution within 3 sigma."""
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(512, seed=31)
    n = 200_000
    out = dev.playouts(boards, tm, n, 11)
    assert (out[:, :3].sum(axis=1) == n).all()
    assert (out[:, 3] >= 1 * n).all() and (out[:, 3] <= 17 * n).all()
    sw = boards.copy(); sw["x"], sw["o"] = boards["o"], boards["x"]
    out2 = dev.playouts(sw, -tm, n, 11)
    assert (out2[:, 0] == out[:, 1]).all() and (out2[:, 1] == out[:, 0]).all() and (out2[:, 3] == out[:, 3]).all()
    def transpose(m):
        r = np.zeros_like(m)
        for x in range(5):
            for y in range(5):
                r |= ((m >> np.uint32(5 * x + y)) & np.uint32(1)) << np.uint32(5 * y + x)
        return r
    tr = boards.copy(); tr["x"], tr["o"] = transpose(boards["x"]), transpose(boards["o"])
    out3 = dev.playouts(tr, tm, n, 12)
    p = out[:, 0] / n
    tol = 4.5 * np.sqrt(2 * p * (1 - p) / n) + 1e-9
    assert (np.abs(out3[:, 0] / n - p) <= tol).all()


def test_overlapping_launches_on_two_streams(dev):
    """two playout kernels in flight on the same device must not share scheduling state"""
    import torch
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(256, seed=41)
    want = dev.playouts(boards, tm, 20000, 6)
    tb = torch.from_numpy(boards.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).cuda()
    tt = torch.from_numpy(tm.copy()).cuda()
    out = torch.zeros((256, 4), dtype=torch.int64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(3):
        dev.playouts_dev(tb[:128].data_ptr(), tt[:128].data_ptr(), 128, 20000, out[:128].data_ptr(),
                         leaf_index_base=0, stream_id=6, stream=s1)
        dev.playouts_dev(tb[128:].data_ptr(), tt[128:].data_ptr(), 128, 20000, out[128:].data_ptr(),
                         leaf_index_base=128, stream_id=6, stream=s2)
        torch.cuda.synchronize()
        assert (out.cpu().numpy().astype(np.uint64) == want).all()
