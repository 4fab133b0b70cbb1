"""src/negascout's ChooseMove on the GPU (one warp per position) vs the CPU oracle: the search is
order-dependent, so EVERYTHING must match: the alpha recorded per root move, the visiting order,
movekeeper's set and deterministic choice, and the staticValue call count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(sq):
    sq.init()
    yield sq
    sq.shutdown()


def check(dev, oracle, boards, depth, sc):
    r = dev.negascout_root(boards, depth, sc)
    for i in range(boards.shape[0]):
        vals, order, bm, bv, fb, leaves = oracle.negascout_root(int(boards[i]["x"]), int(boards[i]["o"]), depth, sc)
        assert r["values"][i].tolist() == vals, (depth, i)
        assert [int(c) for c in r["visit_order"][i] if c >= 0] == order
        assert (int(r["best_mask"][i]), int(r["best_value"][i]), int(r["first_best"][i])) == (bm, bv, fb)
        assert int(r["nodes"][i]) == leaves
    return r


def test_live_oracle_midgame(dev, oracle):
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    for marks, depth, count in [(16, 8, 24), (14, 6, 24), (12, 6, 16), (10, 6, 12), (10, 8, 6), (8, 4, 12), (6, 4, 8),
                                (20, 10, 24), (22, 10, 16), (24, 10, 8), (4, 3, 6), (2, 3, 4), (0, 2, 1), (12, 0, 8), (12, 1, 8),
                                (12, 2, 8), (12, 3, 8)]:
        boards, _ = positions.midgame_batch(count, marks=(marks,), seed=1000 + marks * 16 + depth)
        check(dev, oracle, boards, depth, sc)


def test_setdepth_table_depths_from_opening_positions(dev, oracle):
    """SetDepth gives 6 below four marks (negascout.go:69-79): the biggest trees the engine searches"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(4, marks=(2,), seed=5)
    r = check(dev, oracle, boards, 6, sc)
    assert int(r["nodes"].max()) > 100_000


def test_random_bias_tables(dev, oracle):
    from squava_b200 import positions
    rng = np.random.Generator(np.random.PCG64(8))
    for _ in range(4):
        sc = [int(v) for v in rng.integers(-5, 6, size=25)]
        boards, _ = positions.midgame_batch(12, marks=(10, 12, 14), seed=int(rng.integers(1 << 30)))
        check(dev, oracle, boards, 5, sc)


def test_boards_that_already_hold_lines_and_full_boards(dev, oracle):
    """unreachable inputs: both colours own lines (the scan order decides), no empty cell"""
    sc = oracle.tables()["scores_ab"]
    rng = np.random.Generator(np.random.PCG64(9))
    rows = []
    while len(rows) < 200:
        trits = rng.integers(0, 3, size=25)
        x = sum(1 << c for c in range(25) if trits[c] == 1); o = sum(1 << c for c in range(25) if trits[c] == 2)
        if bin(x | o).count("1") >= 12:
            rows.append((x, o))
    rows.append((0x1364d93, 0x1ffffff & ~0x1364d93))          # a full board
    boards = np.array(rows, dtype=np.uint32).view(dev.BOARD).reshape(-1)
    r = check(dev, oracle, boards, 4, sc)
    assert int(r["first_best"][-1]) == -1 and int(r["best_mask"][-1]) == 0


def test_batch_order_and_size_do_not_matter(dev, oracle):
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(700, marks=(12, 14, 16, 18), seed=77)
    a = dev.negascout_root(boards, 6, sc)
    b = dev.negascout_root(boards[::-1].copy(), 6, sc)
    assert (a["values"] == b["values"][::-1]).all() and (a["nodes"] == b["nodes"][::-1]).all()
    one = dev.negascout_root(boards[123:124], 6, sc)
    assert (one["values"][0] == a["values"][123]).all() and int(one["nodes"][0]) == int(a["nodes"][123])
    empty = dev.negascout_root(boards[:0], 6, sc)
    assert empty["values"].shape == (0, 25)


def same(a, b):
    return all((a[k] == b[k]).all() for k in ("values", "visit_order", "best_mask", "best_value", "first_best", "nodes"))


def test_split_over_many_warps_changes_nothing(dev, oracle, monkeypatch):
    """NsSplit (abi.cu): positions over their node budget, and short lists, are searched as items under loops replayed on
    the host; an item over ITS budget becomes such a loop.  Whatever the split plies, the budgets and the list length: every output equals the one-warp replay's
    (SQ_NS_SPLIT=0), which the tests above pin to the oracle."""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    cases = [(positions.midgame_batch(40, marks=(8, 10, 12, 14), seed=31)[0], 6),
             (positions.midgame_batch(24, marks=(10, 12), seed=32)[0], 8),
             (positions.midgame_batch(6, marks=(4, 6), seed=33)[0], 5),
             (positions.midgame_batch(30, marks=(16, 18, 20, 21), seed=34)[0], 10),
             (positions.midgame_batch(12, marks=(12,), seed=35)[0], 4)]
    for boards, depth in cases:
        monkeypatch.setenv("SQ_NS_SPLIT", "0")
        plain = dev.negascout_root(boards, depth, sc)
        monkeypatch.delenv("SQ_NS_SPLIT")
        for env in [{}, {"SQ_NS_DIRECT": "1000"}, {"SQ_NS_DIRECT": "0", "SQ_NS_BUDGET": "200"},
                    {"SQ_NS_DIRECT": "0", "SQ_NS_BUDGET": "5000", "SQ_NS_SPLIT_PLY": "1"},
                    {"SQ_NS_DIRECT": "1000", "SQ_NS_SPLIT_PLY": "2", "SQ_NS_SPLIT_PLY_NULL": "1"},
                    {"SQ_NS_DIRECT": "1000", "SQ_NS_SPLIT_PLY": "6", "SQ_NS_SPLIT_PLY_NULL": "6"},
                    {"SQ_NS_DIRECT": "1000", "SQ_NS_SPLIT_PLY": "4", "SQ_NS_SPLIT_PLY_NULL": "3", "SQ_NS_ITEM_BUDGET": "0"},
                    {"SQ_NS_DIRECT": "1000", "SQ_NS_SPLIT_PLY": "1", "SQ_NS_ITEM_BUDGET": "40"},
                    {"SQ_NS_DIRECT": "0", "SQ_NS_BUDGET": "1000", "SQ_NS_ITEM_BUDGET": "300"},
                    {"SQ_NS_DIRECT": "1000", "SQ_NS_ITEM_BUDGET": "200", "SQ_NS_ITEM_BUDGET_WIDE": "50"},
                    {"SQ_NS_DIRECT": "1000", "SQ_NS_ITEM_BUDGET": "2000", "SQ_NS_FLIGHTS": "2"}]:
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = dev.negascout_root(boards, depth, sc)
            for k in env:
                monkeypatch.delenv(k)
            assert same(plain, got), (depth, env)
    # and straight against the oracle, split on
    boards, _ = positions.midgame_batch(6, marks=(10,), seed=36)
    check(dev, oracle, boards, 7, sc)


def test_split_searches_from_several_host_threads(dev, oracle):
    """every call context carries its own flights: searches arriving on several OS threads at once (split and one-warp
    passes mixed) return what the same calls return one after the other"""
    import threading
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    jobs = [(positions.midgame_batch(3, marks=(6,), seed=41)[0], 7), (positions.midgame_batch(80, marks=(10, 12), seed=42)[0], 8),
            (positions.midgame_batch(1, marks=(4,), seed=43)[0], 6), (positions.midgame_batch(40, marks=(14,), seed=44)[0], 10),
            (positions.midgame_batch(2, marks=(8,), seed=45)[0], 7), (positions.midgame_batch(200, marks=(12, 16), seed=46)[0], 6)]
    want = [dev.negascout_root(b, d, sc) for b, d in jobs]
    got, errs = [None] * len(jobs), []

    def worker(k):
        try:
            for _ in range(2):
                got[k] = dev.negascout_root(jobs[k][0], jobs[k][1], sc)
        except Exception as e:        # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(len(jobs))]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    for k in range(len(jobs)):
        assert same(want[k], got[k]), k
