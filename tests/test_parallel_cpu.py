"""N > 1 path on CPU: world_size-2 gloo processes, columns dealt round-robin, results
all-gathered; every rank must end with the single-process cv_results_."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sklearn.linear_model import LogisticRegression
    from skdist.distribute.search import DistGridSearchCV
    from skdist_b200 import engine, parallel
    from skdist_b200.datasets import make_g1_classification
    from tests.fake_engine import FakeEngine
    engine.set_engine_factory(FakeEngine)
    X, y = make_g1_classification(1500, 8, seed=9)
    gs = DistGridSearchCV(LogisticRegression(), {"C": [0.01, 0.1, 1.0, 10.0, 100.0]}, cv=3).fit(X, y)
    eng = engine.get_engine()
    n_fit = sum(b for kind, b in eng.calls if kind == "fit")
    # multiclass target: multinomial problems dealt over the ranks, confusion-count scorer
    from skdist_b200.datasets import make_multiclass
    Xm, ym = make_multiclass(600, 6, 4, seed=5)
    gm = DistGridSearchCV(LogisticRegression(max_iter=40), {"C": [0.1, 1.0, 10.0]}, cv=3, scoring="f1_macro").fit(Xm, ym)
    np.savez(os.path.join(out_dir, "multi%d.npz" % rank), mean=gm.cv_results_["mean_test_score"],
             coef=gm.best_estimator_.coef_, n_fit=sum(b for kind, b in eng.calls if kind == "fit_multinomial"))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mean=gs.cv_results_["mean_test_score"],
             s0=gs.cv_results_["split0_test_score"], best=gs.best_index_, coef=gs.best_estimator_.coef_,
             n_fit=n_fit, shard=parallel.shard_indices(15, rank, world))
    # gather helper on a ragged split (7 items over 2 ranks)
    idx = parallel.shard_indices(7, rank, world)
    full = parallel.all_gather_columns(np.stack([idx * 10.0, idx * 1.0], 1), 7, rank, world)
    assert np.array_equal(full[:, 0], np.arange(7) * 10.0)
    # block-cyclic deal over a permuted order (what the search uses: fold-major blocks)
    order = np.random.RandomState(0).permutation(301)
    mine = parallel.shard_blocks(301, rank, world, order, block=64)
    back = parallel.all_gather_blocks(mine.astype(np.float64) * 3.0, 301, rank, world, order, block=64)
    assert np.array_equal(back, np.arange(301) * 3.0)
    assert len(parallel.shard_blocks(5, rank, world)) in (2, 3)      # small problems still use every rank
    # forests: the trees of the other rank arrive as raw node / value arrays in pieces (a tiny piece size
    # forces several all-gathers), both collection modes; regressors have one "class"
    from sklearn.ensemble import RandomForestClassifier, RandomForestRegressor
    from sklearn.utils import check_random_state
    from skdist.distribute.ensemble import DistRandomForestClassifier, DistRandomForestRegressor
    from skdist_b200.distribute.ensemble import MAX_RAND_SEED, _tree_inputs
    Xq = np.round(Xm * 8).astype(np.float32)
    seeds = check_random_state(4).randint(MAX_RAND_SEED, size=7)
    eng.seed_of_rand_r = {int(_tree_inputs(s_, len(ym), False)[1]): int(s_) for s_ in seeds}
    real_gather = parallel.all_gather_trees
    parallel.all_gather_trees = lambda *a, **k: real_gather(*a, piece_bytes=4096, **k)
    try:
        os.environ["SKDIST_B200_FOREST_GATHER"] = "all"
        rf = DistRandomForestClassifier(n_estimators=7, random_state=4).fit(Xq, ym)
        ref = RandomForestClassifier(n_estimators=7, random_state=4).fit(Xq, ym)
        assert len(rf.estimators_) == 7
        for a_, b_ in zip(rf.estimators_, ref.estimators_):
            assert a_.random_state == b_.random_state
            np.testing.assert_array_equal(a_.tree_.threshold, b_.tree_.threshold)
            np.testing.assert_array_equal(a_.tree_.value, b_.tree_.value)
            np.testing.assert_array_equal(a_.tree_.children_right, b_.tree_.children_right)
        np.testing.assert_array_equal(rf.predict_proba(Xq), ref.predict_proba(Xq))
        yr = (Xq[:, 0] * 2 + Xq[:, 1]).astype(np.float64)
        rr = DistRandomForestRegressor(n_estimators=7, random_state=4).fit(Xq, yr)
        rref = RandomForestRegressor(n_estimators=7, random_state=4).fit(Xq, yr)
        np.testing.assert_allclose(rr.predict(Xq), rref.predict(Xq), rtol=1e-12)
        os.environ["SKDIST_B200_FOREST_GATHER"] = "rank0"
        r0 = DistRandomForestClassifier(n_estimators=7, random_state=4).fit(Xq, ym)
        assert len(r0.estimators_) == (7 if rank == 0 else len(parallel.shard_indices(7, rank, world)))
        if rank == 0:
            np.testing.assert_array_equal(r0.predict_proba(Xq), ref.predict_proba(Xq))
        os.environ.pop("SKDIST_B200_FOREST_GATHER")            # default "auto": a small forest is replicated
        ra = DistRandomForestClassifier(n_estimators=7, random_state=4).fit(Xq, ym)
        assert len(ra.estimators_) == 7
    finally:
        parallel.all_gather_trees = real_gather
        os.environ.pop("SKDIST_B200_FOREST_GATHER", None)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_grid_search_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    for k in ("mean", "s0", "coef"):
        np.testing.assert_array_equal(r0[k], r1[k])
    assert int(r0["best"]) == int(r1["best"])
    # each rank fitted only its own shard of the 15 columns (+ 1 refit)
    assert int(r0["n_fit"]) == 8 + 1 and int(r1["n_fit"]) == 7 + 1
    assert np.array_equal(r0["shard"], np.arange(0, 15, 2)) and np.array_equal(r1["shard"], np.arange(1, 15, 2))
    m0, m1 = np.load(tmp_path / "multi0.npz"), np.load(tmp_path / "multi1.npz")
    np.testing.assert_array_equal(m0["mean"], m1["mean"])
    np.testing.assert_array_equal(m0["coef"], m1["coef"])
    assert int(m0["n_fit"]) + int(m1["n_fit"]) == 9 + 2          # 9 (candidate, fold) problems + one refit per rank
    from sklearn.model_selection import GridSearchCV
    from skdist_b200.datasets import make_multiclass
    Xm, ym = make_multiclass(600, 6, 4, seed=5)
    from sklearn.linear_model import LogisticRegression as _LR
    sk = GridSearchCV(_LR(max_iter=40), {"C": [0.1, 1.0, 10.0]}, cv=3, scoring="f1_macro").fit(Xm, ym)
    np.testing.assert_allclose(m0["mean"], sk.cv_results_["mean_test_score"], rtol=1e-12)
    # single-process reference through the same host code
    from sklearn.linear_model import LogisticRegression
    from skdist.distribute.search import DistGridSearchCV
    from skdist_b200 import engine
    from skdist_b200.datasets import make_g1_classification
    from tests.fake_engine import FakeEngine
    engine.set_engine_factory(FakeEngine)
    try:
        X, y = make_g1_classification(1500, 8, seed=9)
        gs = DistGridSearchCV(LogisticRegression(), {"C": [0.01, 0.1, 1.0, 10.0, 100.0]}, cv=3).fit(X, y)
    finally:
        engine.set_engine_factory(None)
    np.testing.assert_array_equal(gs.cv_results_["mean_test_score"], r0["mean"])
    np.testing.assert_array_equal(gs.best_estimator_.coef_, r0["coef"])


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_block_dealing_is_a_partition_and_balances_rounds(world):
    """The headline search (512 C x 5 folds, fold-major order) dealt over `world` ranks: every column goes to
    exactly one rank, every rank gets whole 128-column blocks of one fold, and with the longest-first deal a
    rank that takes one block more than the others gets the blocks that stop early."""
    from skdist_b200 import parallel
    n_c, n_f = 512, 5
    Cs = np.logspace(-4, 4, n_c)
    cols = np.arange(n_c * n_f)
    order = np.concatenate([cols[cols % n_f == f] for f in range(n_f)])     # fold-major: what the search passes
    cost = np.repeat(parallel.logreg_column_cost(Cs), n_f)                   # column = candidate * n_f + fold
    parts = [parallel.shard_blocks(len(cols), r, world, order, 128, cost) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(parts)), cols)
    loads, blocks = [], []
    for p in parts:
        assert len(p) % 128 == 0
        blk = p.reshape(-1, 128)
        assert all(len(set(b % n_f)) == 1 for b in blk)                      # one fold per block
        blocks.append(len(blk))
        loads.append(sum(cost[b].max() for b in blk))                        # a block runs as long as its slowest column
    assert max(blocks) - min(blocks) <= 1
    plain = [parallel.shard_blocks(len(cols), r, world, order, 128) for r in range(world)]
    plain_load = max(sum(cost[b].max() for b in p.reshape(-1, 128)) for p in plain)
    assert max(loads) <= plain_load + 1e-12                                   # never worse than round-robin
    if world == 8:              # 20 blocks over 8 ranks: the third block of a rank is one that stops early
        assert plain_load == pytest.approx(3.0) and max(loads) < 2.7
