"""BASELINE-sized parity cases and the multi-device / multi-thread behaviour of the C ABI.

configs[3] (AB-1, depth 10) on EIGHT-MARK positions -- the 10^6..7*10^7-leaf trees that drive the adaptive
splitting of the GPU search -- against committed oracle vectors (tests/golden/c3_eight_mark_depth10.json,
made by oracle/tools/make_golden_c3.py) and, for the two smallest, against the oracle run live."""
import json
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
C3 = json.load(open(os.path.join(HERE, "golden", "c3_eight_mark_depth10.json")))["cases"]


@pytest.fixture(scope="module")
def dev(sq):
    sq.init()
    yield sq
    sq.shutdown()


def n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def boards_of(pairs, sq):
    return np.array(pairs, dtype=np.uint32).view(sq.BOARD).reshape(-1)


def test_c3_eight_mark_positions_depth10_golden(dev, oracle):
    sc = oracle.tables()["scores_ab"]
    assert len(C3) >= 8 and all(bin(c["x"] | c["o"]).count("1") == 8 and c["depth"] == 10 for c in C3)
    b = boards_of([(c["x"], c["o"]) for c in C3], dev)
    values, bv, bm, nodes = dev.alphabeta_root(b, 1, 10, sc)
    for i, c in enumerate(C3):
        assert values[i].tolist() == c["values"], i
        assert (int(bm[i]), int(bv[i])) == (c["best_mask"], c["best_value"])
    # one at a time (a lone big tree takes the deep-splitting path) and inside a mixed batch (many small units first)
    i = max(range(len(C3)), key=lambda k: C3[k]["leaves"])
    v1, _, m1, _ = dev.alphabeta_root(b[i:i + 1], 1, 10, sc)
    assert v1[0].tolist() == C3[i]["values"] and int(m1[0]) == C3[i]["best_mask"]
    from squava_b200 import positions
    small, _ = positions.midgame_batch(200, marks=(12, 14), seed=17)
    mixed = np.concatenate([small[:100], b, small[100:]])
    vm, _, _, _ = dev.alphabeta_root(mixed, 1, 10, sc)
    for k, c in enumerate(C3):
        assert vm[100 + k].tolist() == c["values"], k


def test_c3_eight_mark_positions_depth10_live_oracle(dev, oracle):
    """the golden file is not trusted blindly: the two smallest cases are searched again by the CPU oracle here"""
    sc = oracle.tables()["scores_ab"]
    pick = sorted(range(len(C3)), key=lambda k: C3[k]["leaves"])[:2]
    with ThreadPoolExecutor(max_workers=2) as ex:
        live = list(ex.map(lambda k: oracle.ab1_root(C3[k]["x"], C3[k]["o"], 10, sc), pick))
    values, bv, bm, _ = dev.alphabeta_root(boards_of([(C3[k]["x"], C3[k]["o"]) for k in pick], dev), 1, 10, sc)
    for j, k in enumerate(pick):
        ov, obm, obv, leaves = live[j]
        assert ov == C3[k]["values"] and leaves == C3[k]["leaves"]
        assert values[j].tolist() == ov and (int(bm[j]), int(bv[j])) == (obm, obv)


def test_c4_opening3_depth10_golden(dev):
    """configs[4] at BASELINE size: opening3 -d 10 (the program's default), all 85 subtrees, against the oracle's
    committed vectors (tests/golden/c4_opening3_depth10.json, oracle/tools/make_golden_c4.py)"""
    g = json.load(open(os.path.join(HERE, "golden", "c4_opening3_depth10.json")))["cases"]
    assert len(g) == 85 and all(c["depth"] == 10 for c in g)
    so = [3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 2, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3]
    st = np.zeros(85, dev.SUBTREE)
    for i, c in enumerate(g):
        st[i] = (1 << c["first"], 1 << c["reply"], c["reply"], 2, 1, 0, so[c["first"]])
    v, nodes = dev.alphabeta_subtrees(st, 3, 10, so)
    assert v.tolist() == [c["value"] for c in g]
    assert (nodes > 0).all()


def test_near_full_boards_search_deeper_than_the_empties(dev, oracle):
    """max_depth > empty cells: the search runs into full boards (node value -+20000, which outranks a win) and
    root windows get cut at their edge -- squavathr -d 15 and playoff depth 12 do this in every endgame.
    Includes the four lineless full boards with k marks taken off."""
    sc = oracle.tables()["scores_ab"]
    rng = np.random.Generator(np.random.PCG64(5))
    pairs = []
    for full13 in (0x1364d93, 0x1552ab5, 0x15aa955, 0x19364d9):
        full12 = 0x1ffffff & ~full13                    # 13 / 12 marks, no line anywhere: every subset is lineless too
        c13 = [c for c in range(25) if full13 >> c & 1]; c12 = [c for c in range(25) if full12 >> c & 1]
        for nb in (0, 1, 2, 3):
            for _ in range(3):
                # X holds the 13 (X moved first): nb + 1 X and nb O taken off -> X to move, X places the last mark,
                # the full board is a MIN node worth +20000
                x, o = full13, full12
                for c in rng.choice(c13, size=nb + 1, replace=False): x &= ~(1 << int(c))
                for c in rng.choice(c12, size=nb, replace=False): o &= ~(1 << int(c))
                pairs.append((x, o))
                # O holds the 13 (the engine moves second and sees itself as X, playoff5.go:80,95): the full board is a
                # MAX node worth -20000 -- the root-window edge a cut-off unit must not overwrite
                x, o = full12, full13
                for c in rng.choice(c12, size=nb + 1, replace=False): x &= ~(1 << int(c))
                for c in rng.choice(c13, size=nb + 1, replace=False): o &= ~(1 << int(c))
                pairs.append((x, o))
    from squava_b200 import positions
    more, _ = positions.midgame_batch(24, marks=(18, 20, 22), seed=91)
    pairs += [(int(b["x"]), int(b["o"])) for b in more]
    assert len(pairs) > 100
    b = boards_of(pairs, dev)
    sentinels = 0
    for dialect, f in ((1, oracle.ab1_root), (2, oracle.ab2_root), (5, oracle.ab5_root), (6, oracle.ab6_root)):
        for depth in (8, 12, 15):
            values, bv, bm, _ = dev.alphabeta_root(b, dialect, depth, sc)
            for i, (x, o) in enumerate(pairs):
                ov, obm, obv, _ = f(x, o, depth, sc)
                assert values[i].tolist() == ov, (dialect, depth, hex(x), hex(o))
                assert (int(bm[i]), int(bv[i])) == (obm, obv)
                sentinels += sum(abs(v) == 20000 for v in ov)
    assert sentinels > 100            # both +20000 and -20000 root values are in play here


def test_packed_outcomes_equal_the_counted_form(dev):
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(5000, seed=12)
    boards = np.concatenate([boards, boards_of([(0b1111, 0b1110000000), (0x1364d93, 0x1ffffff & ~0x1364d93), (0, 0)], dev)])
    tm = np.concatenate([tm, np.array([-1, -1, 1], np.int8)])
    for stream in (0, 7, (3 << 32) | 1):
        full = dev.playouts(boards, tm, 1, stream)
        pk = dev.playouts_packed(boards, tm, stream)
        code = pk & 3
        assert (code >= 1).all()
        assert ((code == 1) == (full[:, 0] == 1)).all() and ((code == 2) == (full[:, 1] == 1)).all()
        assert ((code == 3) == (full[:, 2] == 1)).all() and ((pk >> 2) == full[:, 3]).all()
    big = np.tile(boards, 20)[:70000]; bt = np.tile(tm, 20)[:70000]          # above the small-call staging size
    full = dev.playouts(big, bt, 1, 5); pk = dev.playouts_packed(big, bt, 5)
    assert ((pk >> 2) == full[:, 3]).all() and ((pk & 3 == 1) == (full[:, 0] == 1)).all()


def test_playouts_on_one_slot_equal_the_sharded_call(dev):
    """sq_playouts_on / sq_playouts_packed_on (root-parallel hosts: one tree, one device) give the counts the
    sharded call gives: the stream depends on (seed, stream_id, leaf index) only"""
    from squava_b200 import positions
    import ctypes as C
    boards, tm = positions.midgame_batch(700, seed=71)
    want = dev.playouts(boards, tm, 3000, 5)
    assert (dev.playouts(boards, tm, 3000, 5, slot=0) == want).all()
    pk = np.empty(700, np.uint8)
    rc = dev.lib().sq_playouts_packed_on(0, boards.ctypes.data, tm.ctypes.data, 700, 9, pk.ctypes.data)
    assert rc == 0 and (pk == dev.playouts_packed(boards, tm, 9)).all()
    with pytest.raises(dev.SquavaError):
        dev.playouts(boards, tm, 10, 5, slot=99)


def test_concurrent_host_threads_on_one_device(dev, oracle):
    """calls arriving on several OS threads (cgo's situation) each take their own call context: results must be
    exactly those of the same calls made one after the other"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, tm = positions.midgame_batch(2048, seed=52)
    ab_b, _ = positions.midgame_batch(64, marks=(12, 14), seed=53)
    want_p = [dev.playouts(boards, tm, 3000, s) for s in range(6)]
    want_a = dev.alphabeta_root(ab_b, 1, 10, sc)[0]
    want_c = dev.classify(boards)
    got_p, got_a, got_c, errs = [None] * 6, [None] * 3, [None] * 3, []

    def worker(k):
        try:
            for rep in range(3):
                got_p[k] = dev.playouts(boards, tm, 3000, k)
                if k < 3:
                    got_a[k] = dev.alphabeta_root(ab_b, 1, 10, sc)[0]
                    got_c[k] = dev.classify(boards)
        except Exception as e:        # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    for k in range(6):
        assert (got_p[k] == want_p[k]).all(), k
    for k in range(3):
        assert (got_a[k] == want_a).all() and all((a == b).all() for a, b in zip(got_c[k], want_c))


@pytest.mark.skipif(n_devices() < 2, reason="needs two GPUs")
def test_two_devices_in_process_equal_one(sq, oracle):
    """sq_init(ids, 2): host-buffer calls shard over the devices (what squavam/squavathr/opening -gpus N use);
    every result must equal the one-device result bit for bit"""
    from squava_b200 import positions
    sc, so = oracle.tables()["scores_ab"], oracle.tables()["scores_opening"]
    boards, tm = positions.midgame_batch(3001, seed=61)
    ab_b, _ = positions.midgame_batch(33, marks=(10, 12, 14), seed=62)
    st = np.zeros(85, sq.SUBTREE)
    for i, (f, r) in enumerate(oracle.opening3_subtrees()):
        st[i] = (1 << f, 1 << r, r, 2, 1, 0, so[f])
    res = {}
    for devs in ([0], [0, 1]):
        sq.init(devices=devs)
        try:
            assert sq.device_count() == len(devs)
            res[len(devs)] = (sq.classify(boards), sq.playouts(boards, tm, 5000, 3), sq.playouts_packed(boards, tm, 4),
                              sq.alphabeta_root(ab_b, 1, 10, sc)[:3], sq.alphabeta_subtrees(st, 3, 8, so)[0],
                              sq.negascout_root(ab_b, 6, sc))
        finally:
            sq.shutdown()
    a, b = res[1], res[2]
    assert all((p == q).all() for p, q in zip(a[0], b[0]))
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all()
    assert all((p == q).all() for p, q in zip(a[3], b[3])) and (a[4] == b[4]).all()
    assert all((a[5][k] == b[5][k]).all() for k in a[5])


@pytest.mark.skipif(n_devices() < 2, reason="needs two GPUs")
def test_calls_on_two_devices_overlap(sq):
    """two host threads, one device each: the second call must not wait for the first (one mutex per call
    context, none across devices).  Both calls together must take clearly less than twice one call."""
    import time
    import torch
    from squava_b200 import positions
    boards, tm = positions.midgame_batch(4096, seed=63)
    sq.init(devices=[0, 1])
    try:
        bufs = []
        for d in (0, 1):
            dv = torch.device("cuda", d)
            tb = torch.from_numpy(boards.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dv)
            tt = torch.from_numpy(tm.copy()).to(dv)
            out = torch.zeros((4096, 4), dtype=torch.int64, device=dv)
            bufs.append((tb, tt, out, torch.cuda.Stream(device=dv)))

        def one(d, reps):
            tb, tt, out, s = bufs[d]
            for r in range(reps):
                sq.playouts_dev(tb.data_ptr(), tt.data_ptr(), 4096, 200_000, out.data_ptr(), stream_id=r, slot=d, stream=s)
            s.synchronize()
        one(0, 1); one(1, 1)
        t0 = time.perf_counter(); one(0, 4); t_one = time.perf_counter() - t0
        th = [threading.Thread(target=one, args=(d, 4)) for d in (0, 1)]
        t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; t_both = time.perf_counter() - t0
        assert t_both < 1.5 * t_one, (t_one, t_both)
        assert (bufs[0][2].cpu() == bufs[1][2].cpu()).all()
    finally:
        sq.shutdown()
