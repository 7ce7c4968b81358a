"""CPU tests of the oracle (the parity checker): known-answer tests, the committed golden vectors, and
hand-derivable properties of the restated LightGBM 3.2.x rules.  No GPU needed."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))
DS_PARAMS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle
    return oracle


# ---- reference's own known-answer tests -----------------------------------------------------------
def test_count_cardinality_reference_kats(O):
    """VerifyLightGBMRanker.scala:127-137 — the only true KATs the reference holds for this path."""
    assert O.count_cardinality([1, 1, 2, 2, 2, 3]) == [2, 3, 1]
    assert O.count_cardinality([1, 1, 1, 1]) == [4]
    assert O.count_cardinality([5]) == [1]
    assert O.count_cardinality([]) == []
    from mmlspark_b200.lightgbm import count_cardinality
    for ids in ([1, 1, 2, 2, 2, 3], [3, 3, 1, 1, 1, 2, 2], [7]):
        assert count_cardinality(ids) == O.count_cardinality(ids)


# ---- LCG sampler: independent restatement in Python ---------------------------------------------------
class _Lcg:
    def __init__(self, seed): self.x = seed & 0xFFFFFFFF
    def step(self): self.x = (214013 * self.x + 2531011) & 0xFFFFFFFF; return self.x
    def next_float(self): return np.float32((self.step() >> 16) & 0x7FFF) / np.float32(32768.0)
    def next_int(self, lo, hi): return (self.step() & 0x7FFFFFFF) % (hi - lo) + lo

    def sample(self, n, k):
        if k > n or k <= 0: return []
        if k == n: return list(range(n))
        if k > 1 and k > n / np.log2(k):
            out = []
            for i in range(n):
                if self.next_float() < (k - len(out)) / float(n - i):
                    out.append(i)
            return out
        s = set()
        for r in range(n - k, n):
            v = self.next_int(0, r)
            if v in s: s.add(r)
            else: s.add(v)
        return sorted(s)


@pytest.mark.parametrize("seed,n,k", [(1, 20, 5), (1, 100, 80), (7, 1000, 10), (1, 50, 50), (3, 250000, 200000), (1, 12, 1)])
def test_lcg_sampler(O, seed, n, k):
    got = O.random_sample(seed, n, k).tolist()
    assert got == _Lcg(seed).sample(n, k)
    key = "%d_%d_%d" % (seed, n, k)
    if key in GOLDEN["lcg_sample"]:
        assert got == GOLDEN["lcg_sample"][key]
    assert got == sorted(set(got)) and all(0 <= v < n for v in got)


# ---- bin finder -------------------------------------------------------------------------------------
def test_bins_few_distinct_values_midpoints(O):
    """<= max_bin distinct values: boundaries at midpoints once >= min_data_in_bin(3) samples accumulated; zero has its own bin."""
    X = np.repeat(np.array([[1.0], [2.0], [3.0], [4.0]]), 25, axis=0)
    ds = O.OracleDataset(X, DS_PARAMS)
    ub = ds.upper_bounds(0)
    info = ds.feature_info(0)
    assert info["missing_type"] == 0 and not info["is_trivial"]
    # values are all positive: first bound is the zero bin edge 1e-35, then midpoints nudged up by one ulp
    assert ub[0] == 1e-35
    assert ub[1] == np.nextafter(1.5, np.inf) and ub[2] == np.nextafter(2.5, np.inf) and ub[3] == np.nextafter(3.5, np.inf)
    assert np.isinf(ub[-1]) and info["num_bin"] == len(ub) == 5
    assert np.array_equal(np.unique(ds.bins()[:, 0]), [1, 2, 3, 4])


def test_bins_nan_gets_last_bin_and_zero_own_bin(O):
    rng = np.random.default_rng(0)
    v = rng.standard_normal(5000)
    v[:500] = np.nan
    v[500:1500] = 0.0
    ds = O.OracleDataset(v[:, None], DS_PARAMS)
    info = ds.feature_info(0)
    ub = ds.upper_bounds(0)
    assert info["missing_type"] == 2 and np.isnan(ub[-1]) and np.isinf(ub[-2])
    b = ds.bins()[:, 0]
    assert (b[:500] == info["num_bin"] - 1).all()
    assert len(set(b[500:1500])) == 1 and b[500] == info["default_bin"]
    zb = info["default_bin"]
    assert ub[zb] == 1e-35 and ub[zb - 1] == -1e-35
    assert info["num_bin"] <= 255


def test_bins_monotone_and_value_to_bin_is_lower_bound(O):
    rng = np.random.default_rng(2)
    X = rng.exponential(1.0, (30000, 3)) * np.array([1.0, -1.0, 100.0])
    ds = O.OracleDataset(X, DS_PARAMS)
    bins = ds.bins()
    for f in range(3):
        ub = ds.upper_bounds(f)
        assert (np.diff(ub) > 0).all()
        assert np.array_equal(bins[:, f], np.searchsorted(ub[:-1], X[:, f], side="left"))
        order = np.argsort(X[:, f])
        assert (np.diff(bins[order, f].astype(int)) >= 0).all()


def test_trivial_and_prefiltered_features(O):
    rng = np.random.default_rng(3)
    n = 4000
    X = np.stack([np.full(n, 2.5), (np.arange(n) < 5).astype(float), rng.standard_normal(n)], axis=1)
    ds = O.OracleDataset(X, DS_PARAMS)
    assert ds.feature_info(0)["is_trivial"]                      # one distinct value -> (after the zero bin) cannot split
    assert ds.feature_info(1)["is_trivial"]                      # 5 ones out of 4000 < min_data_in_leaf(20): feature_pre_filter
    assert not ds.feature_info(2)["is_trivial"]
    assert (ds.bins()[:, :2] == 0).all()


def test_sampling_kicks_in_above_bin_construct_sample_cnt(O):
    rng = np.random.default_rng(4)
    X = rng.standard_normal((30000, 2))
    a = O.OracleDataset(X, DS_PARAMS)
    b = O.OracleDataset(X, DS_PARAMS.replace("200000", "5000"))
    assert a.upper_bounds(0).tobytes() != b.upper_bounds(0).tobytes()
    rows = O.random_sample(1, 30000, 5000)
    c = O.OracleDataset(X[rows], DS_PARAMS)                      # binning the sampled rows alone reproduces the boundaries
    # filter_cnt differs (n changes) but boundaries only depend on the sample
    assert c.upper_bounds(0).tobytes() == b.upper_bounds(0).tobytes()


def test_distributed_bin_finding_ownership(O):
    """Multi-rank rule: rank r finds the bins of feature slice r from ITS OWN rows."""
    rng = np.random.default_rng(5)
    n, F = 6000, 5
    X = rng.standard_normal((n, F))
    X[3000:] *= 3.0                                               # the two shards have different distributions
    ds2 = O.OracleDataset(X, DS_PARAMS, rank_rows=[3000, 3000])
    r0 = O.OracleDataset(X[:3000], DS_PARAMS)
    r1 = O.OracleDataset(X[3000:], DS_PARAMS)
    step = 3                                                      # ceil(5/2)
    for f in range(F):
        owner = r0 if f < step else r1
        assert ds2.upper_bounds(f).tobytes() == owner.upper_bounds(f).tobytes()


def test_golden_bins(O):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    X, s = mg.dataset(123, 2000, 10)
    ds = O.OracleDataset(X, DS_PARAMS)
    g = GOLDEN["bins"]
    for f in range(10):
        assert ds.feature_info(f) == g["feature_info"][f]
        assert [float(v).hex() for v in ds.upper_bounds(f)] == g["upper_bounds_hex"][f]
    assert ds.bins().astype(np.int64).sum(axis=0).tolist() == g["bins_sum_per_feature"]
    assert ds.bins()[:5].tolist() == g["bins_first_rows"]


# ---- split scan -------------------------------------------------------------------------------------
def test_best_split_hand_computed(O):
    case = GOLDEN["best_split_case"]
    hist = np.zeros((256, 2)); hist[:4] = np.array(case["hist4"])
    r = O.best_split(hist, 4, 0, 0, 0, 10.0, 120.0, 120)
    # candidates: thr0 47.78, thr1 68.33, thr2 63.33; shift = 10^2/120
    assert r["threshold"] == 1 and r["default_left"] == 1
    assert abs(r["gain"] - (1600 / 60 + 2500 / 60 - 100 / 120)) < 1e-9
    assert r["left_count"] == 60 and r["right_count"] == 60
    assert abs(r["lout"] - 40 / 60) < 1e-12 and abs(r["rout"] + 50 / 60) < 1e-12
    assert r == pytest.approx(case["result"])


def test_best_split_respects_min_data_and_ties(O):
    hist = np.zeros((256, 2))
    hist[:6] = [[-5, 10], [0, 0], [0, 0], [0, 0], [5, 10], [1, 1]]
    # thresholds 0..3 give identical partitions ({0} | {4,5}) -> identical gains; right-to-left scan with strict '>' keeps the
    # first seen = the HIGHEST threshold (R6)
    r = O.best_split(hist, 6, 0, 0, 0, 1.0, 21.0, 21, min_data_in_leaf=1)
    assert r["threshold"] == 3
    r20 = O.best_split(hist, 6, 0, 0, 0, 1.0, 21.0, 21, min_data_in_leaf=20)
    assert r20["splittable"] == 0 and r20["gain"] == -np.inf
    # l2 regularisation lowers the gain, min_gain_to_split shifts it
    g0 = O.best_split(hist, 6, 0, 0, 0, 1.0, 21.0, 21, min_data_in_leaf=1)["gain"]
    g1 = O.best_split(hist, 6, 0, 0, 0, 1.0, 21.0, 21, min_data_in_leaf=1, l2=5.0)["gain"]
    g2 = O.best_split(hist, 6, 0, 0, 0, 1.0, 21.0, 21, min_data_in_leaf=1, min_gain_to_split=1.0)["gain"]
    assert g1 < g0 and abs((g0 - g2) - 1.0) < 1e-12


def test_best_split_nan_two_way_default_direction(O):
    """With a NaN bin both scan directions run; NaN rows carrying positive gradient are best sent right."""
    hist = np.zeros((256, 2))
    hist[:4] = [[-20, 20], [-18, 20], [2, 20], [40, 20]]      # bin 3 = NaN bin
    r = O.best_split(hist, 4, 2, 1, 1, 4.0, 80.0, 80, min_data_in_leaf=1)
    # forward scan (NaN goes right): thr0 29.6, thr1 80.2, thr2 = 36^2/60 + 40^2/20 = 101.6 -> {bins 0..2} | {NaN}
    assert r["default_left"] == 0 and r["threshold"] == 2
    assert abs(r["gain"] - (36.0 ** 2 / 60 + 40.0 ** 2 / 20 - 16.0 / 80)) < 1e-9
    hist[3] = [-40, 20]                                        # now NaN behaves like the low bins -> default left
    r = O.best_split(hist, 4, 2, 1, 1, -76.0, 80.0, 80, min_data_in_leaf=1)
    assert r["default_left"] == 1


# ---- whole training runs ------------------------------------------------------------------------------
def test_golden_models(O):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    X, s = mg.dataset(123, 2000, 10)
    labels = {"regression": s.astype(np.float32), "binary": (s > 0).astype(np.float32), "multiclass": np.clip(np.floor(s + 1.5), 0, 2).astype(np.float32)}
    from mmlspark_b200.modeltext import parse_model, compare_models
    for name, g in GOLDEN["models"].items():
        Xc = mg.variant(X, name)
        d = O.OracleDataset(Xc, DS_PARAMS + (" " + name.split("|")[1] if "|" in name else "")).set_field("label", labels[name.split("_")[0].split("|")[0]])
        b = O.OracleBooster(d, g["params"])
        b.train(5)
        compare_models(parse_model(b.model_string()), parse_model(g["model"]), value_tol=1e-12, gain_tol=1e-6)
        np.testing.assert_allclose(b.predict_raw(Xc[:8]), np.array(g["raw_pred_first8"]), rtol=1e-12)


def test_training_scores_equal_model_predictions_and_loss_decreases(O):
    rng = np.random.default_rng(8)
    n, F = 20000, 12
    X = rng.standard_normal((n, F))
    y = (X[:, 0] + X[:, 1] * X[:, 2] > 0).astype(np.float32)
    d = O.OracleDataset(X, DS_PARAMS).set_field("label", y)
    b = O.OracleBooster(d, "objective=binary num_leaves=15 verbosity=-1")
    losses = []
    for _ in range(15):
        b.update()
        p = 1 / (1 + np.exp(-b.scores()))
        losses.append(-np.mean(np.where(y > 0, np.log(p), np.log(1 - p))))
    assert all(a > c for a, c in zip(losses, losses[1:]))
    np.testing.assert_allclose(b.predict_raw(X)[:, 0], b.scores(), rtol=1e-10, atol=1e-10)
    tr = b.trace()
    assert (tr[:, 5] + tr[:, 6] > 0).all() and (tr[:, 8] > 0).all()       # counts, gains
    assert (tr[:, 5] >= 20).all() and (tr[:, 6] >= 20).all()               # min_data_in_leaf


def test_structure_agrees_with_sklearn_hgb_first_split(O):
    """Secondary sanity (not parity): on an easy target sklearn's histogram GBDT picks the same root split feature."""
    from sklearn.ensemble import HistGradientBoostingRegressor
    rng = np.random.default_rng(9)
    n = 5000
    X = rng.standard_normal((n, 6))
    y = (3 * (X[:, 2] > 0.3) + 0.1 * rng.standard_normal(n)).astype(np.float32)
    d = O.OracleDataset(X, DS_PARAMS).set_field("label", y)
    b = O.OracleBooster(d, "objective=regression num_leaves=7 verbosity=-1")
    b.update()
    tr = b.trace()
    h = HistGradientBoostingRegressor(max_iter=1, max_leaf_nodes=7, min_samples_leaf=20, learning_rate=0.1, max_bins=255, early_stopping=False).fit(X, y)
    nodes = h._predictors[0][0].nodes
    assert int(tr[0, 3]) == int(nodes[0]["feature_idx"]) == 2
    thr = d.upper_bounds(2)[int(tr[0, 4])]
    assert abs(thr - 0.3) < 0.05 and abs(nodes[0]["num_threshold"] - 0.3) < 0.05


def test_multirank_emulation_counts_come_from_hessians(O):
    """Data-parallel learner: recorded leaf counts are hessian-reconstructed, so with non-constant hessians they may differ
    from the true row counts while the serial learner records true counts (R4)."""
    rng = np.random.default_rng(10)
    n, F = 8000, 6
    X = rng.standard_normal((n, F))
    y = (X[:, 0] + 0.5 * rng.standard_normal(n) > 0).astype(np.float32)
    from mmlspark_b200.modeltext import parse_model
    p = "objective=binary num_leaves=7 verbosity=-1"
    s = O.OracleBooster(O.OracleDataset(X, DS_PARAMS).set_field("label", y), p); s.train(3)
    m = O.OracleBooster(O.OracleDataset(X, DS_PARAMS, rank_rows=[4000, 4000]).set_field("label", y), p); m.train(3)
    ts, tm = parse_model(s.model_string())["trees"], parse_model(m.model_string())["trees"]
    assert ts[0]["leaf_count"].sum() == n                    # serial: true counts
    assert abs(int(tm[0]["leaf_count"].sum()) - n) <= 7      # emulation: rounded hessian shares, close but need not be exact


def test_dataset_from_precomputed_bins_trains_the_same_model(O):
    """OracleDataset.from_bins (used by the mid-scale GPU parity test, where the raw matrix never exists on the host) is the same
    dataset as the one built from the raw matrix."""
    rng = np.random.default_rng(41)
    n, F = 6000, 9
    X = rng.standard_normal((n, F))
    X[:, 2] = np.where(rng.random(n) < 0.3, np.nan, X[:, 2])
    X[:, 4] = 1.0
    y = (X[:, 0] + np.nan_to_num(X[:, 2]) * X[:, 1] > 0).astype(np.float32)
    params = "objective=binary num_leaves=15 min_data_in_leaf=20 learning_rate=0.1 verbosity=-1"
    a = O.OracleDataset(X, "max_bin=63").set_field("label", y)
    infos = [a.feature_info(f) for f in range(F)]
    uppers = [a.upper_bounds(f) for f in range(F)]
    minmax = [(float(np.nanmin(X[:, f])), float(np.nanmax(X[:, f]))) for f in range(F)]
    b = O.OracleDataset.from_bins(a.bins(), infos, uppers, minmax, "max_bin=63").set_field("label", y)
    ba, bb = O.OracleBooster(a, params), O.OracleBooster(b, params)
    ba.train(5); bb.train(5)
    ta = ba.model_string().split("feature_infos=")[1].split("\n", 1)[1]
    tb = bb.model_string().split("feature_infos=")[1].split("\n", 1)[1]
    assert ta == tb
    assert np.array_equal(ba.scores(), bb.scores())


def test_multiclassova_is_k_independent_binary_problems(O):
    """[UPSTREAM MulticlassOVA]: class k's trees equal the trees of a binary booster trained on (label == k)."""
    rng = np.random.default_rng(43)
    n, F, K = 5000, 6, 3
    X = rng.standard_normal((n, F))
    y = np.argmax(X[:, :K] + 0.3 * rng.standard_normal((n, K)), axis=1).astype(np.float32)
    base = "num_leaves=7 min_data_in_leaf=20 learning_rate=0.2 verbosity=-1 sigmoid=1.3 "
    ds = O.OracleDataset(X, "max_bin=63").set_field("label", y)
    ova = O.OracleBooster(ds, base + "objective=multiclassova num_class=3")
    ova.train(4)
    sc = ova.scores().reshape(K, n)
    assert "objective=multiclassova num_class:3 sigmoid:1.3" in ova.model_string()
    for k in range(K):
        dk = O.OracleDataset(X, "max_bin=63").set_field("label", (y == k).astype(np.float32))
        bk = O.OracleBooster(dk, base + "objective=binary")
        bk.train(4)
        np.testing.assert_allclose(sc[k], bk.scores(), rtol=1e-12, atol=1e-12)


def test_cross_entropy_on_hard_labels_equals_binary_logloss(O):
    """with labels in {0, 1} and sigmoid = 1 the cross-entropy gradients are the binary log-loss gradients"""
    rng = np.random.default_rng(47)
    n, F = 4000, 5
    X = rng.standard_normal((n, F))
    y = (X[:, 0] - X[:, 1] > 0).astype(np.float32)
    base = "num_leaves=7 min_data_in_leaf=20 learning_rate=0.2 verbosity=-1 "
    a = O.OracleBooster(O.OracleDataset(X, "max_bin=63").set_field("label", y), base + "objective=cross_entropy")
    b = O.OracleBooster(O.OracleDataset(X, "max_bin=63").set_field("label", y), base + "objective=binary")
    a.train(5); b.train(5)
    np.testing.assert_allclose(a.scores(), b.scores(), rtol=1e-9, atol=1e-9)
