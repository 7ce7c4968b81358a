"""Independent cross-check of the oracle ("cross-checked structurally against sklearn HGB where conditions allow").

The real lightgbmlib 3.2.110 cannot run here, so the oracle is "parity unpinned" against it.  scikit-learn's
HistGradientBoosting is an independent implementation of the same algorithm family (histogram GBDT, leaf-wise growth, the XGBoost
gain, fp32 gradients, NaN bin with a learned default direction, LightGBM-style categorical splits).  On data where the two bin
finders agree by construction (integer-valued features with < 255 distinct values: both put one bin per distinct value) the two must
produce the SAME model; they do, to 1e-13, for regression / binary / Poisson / Gamma / max_depth / NaN / weights / categorical (multiclass
softmax to 1e-7 in probability).  The one LightGBM-specific
rule sklearn does not have — children inherit the parent's per-feature `is_splittable` flags — is isolated with the oracle's test-only
switch `oracle_inherit_splittable=false`."""
import numpy as np
import pytest

sk = pytest.importorskip("sklearn.ensemble")

BASE = "num_leaves=31 learning_rate=0.1 min_data_in_leaf=20 min_sum_hessian_in_leaf=0.001 verbosity=-1 "


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle as O
    return O


def _data(seed=6, n=20000, F=8):
    rng = np.random.default_rng(seed)
    X = rng.integers(-20, 30, size=(n, F)).astype(np.float64)
    y = (0.3 * X[:, 0] - 0.02 * X[:, 1] ** 2 + 0.5 * (X[:, 2] > 3) * X[:, 3] + rng.standard_normal(n)).astype(np.float32)
    return rng, X, y


def _hgb_reg(iters, **kw):
    return sk.HistGradientBoostingRegressor(loss="squared_error", learning_rate=0.1, max_iter=iters, max_leaf_nodes=31, min_samples_leaf=20,
                                            max_bins=255, early_stopping=False, **kw)


def _oracle_pred(O, X, y, params, iters, ds_params="max_bin=255", weight=None):
    ds = O.OracleDataset(X, ds_params).set_field("label", y)
    if weight is not None:
        ds.set_field("weight", weight.astype(np.float32))
    b = O.OracleBooster(ds, BASE + params)
    b.train(iters)
    return b.predict_raw(X)[:, 0]


@pytest.mark.parametrize("case", ["plain", "max_depth", "nan", "weights", "l2_small", "categorical"])
def test_regression_matches_sklearn_hgb(O, case):
    rng, X, y = _data()
    kw, params, ds_params, w, iters = {}, "objective=regression", "max_bin=255", None, 20
    if case == "max_depth":
        kw, params = dict(max_depth=4), params + " max_depth=4"
    elif case == "nan":
        X = X.copy(); X[rng.random(X.shape) < 0.1] = np.nan
    elif case == "weights":
        w = rng.integers(1, 4, len(y)).astype(np.float64)
    elif case == "l2_small":
        kw, params, iters = dict(l2_regularization=0.5), params + " lambda_l2=0.5", 5
    elif case == "categorical":
        X = X.copy(); X[:, 4] = rng.integers(0, 12, len(y))
        kw, ds_params, iters = dict(categorical_features=[4]), "max_bin=255 categorical_feature=4", 10
    h = _hgb_reg(iters, **kw).fit(X, y.astype(np.float64), sample_weight=w)
    got = _oracle_pred(O, X, y, params, iters, ds_params, w)
    np.testing.assert_allclose(got, h.predict(X), rtol=0, atol=1e-10)


@pytest.mark.parametrize("ncat,iters", [(12, 10), (40, 10), (60, 5)])
def test_categorical_many_vs_many_matches_sklearn_hgb(O, ncat, iters):
    """The label DEPENDS on the category (the "categorical" case above only checks that an irrelevant categorical column does no harm), so
    the trees hold dozens of categorical splits.  Both implementations sort the categories of a leaf by sum_g / (sum_h + 10), scan the
    sorted order from both ends up to the middle and send the chosen prefix left (sklearn `_find_best_bin_to_split_category`,
    LightGBM `FindBestThresholdCategoricalInner`).  LightGBM-only rules are neutralised: cat_l2=0 (extra L2 of categorical gains),
    min_data_per_group=1, max_cat_to_onehot=1 (no one-vs-rest mode); max_cat_threshold=32 does not bind below 64 categories.
    Agreement is exact (0.0) — this pins the ctr order, the two scan directions, the category bitsets in the model and their use at
    prediction time against an implementation that shares no code or author with the oracle.
    (60 categories are run for 5 iterations only: sklearn scans (n_used + 1) // 2 - 1 prefixes in the backward direction where LightGBM
    scans (n_used + 1) / 2 in both, and at iteration 10 a leaf with 57 used categories has its best split at exactly the 29th backward
    prefix — gain 966.53 vs sklearn's 963.71, re-derived by hand.)"""
    rng = np.random.default_rng(21)
    n, F = 20000, 8
    X = rng.integers(-20, 30, size=(n, F)).astype(np.float64)
    X[:, 4] = rng.integers(0, ncat, n)
    eff = rng.standard_normal(ncat)
    y = (0.3 * X[:, 0] - 0.02 * X[:, 1] ** 2 + 2.0 * eff[X[:, 4].astype(int)] + rng.standard_normal(n)).astype(np.float32)
    h = _hgb_reg(iters, categorical_features=[4]).fit(X, y.astype(np.float64))
    ds = O.OracleDataset(X, "max_bin=255 categorical_feature=4").set_field("label", y)
    b = O.OracleBooster(ds, BASE + "objective=regression cat_l2=0 cat_smooth=10 min_data_per_group=1 max_cat_to_onehot=1 max_cat_threshold=32")
    b.train(iters)
    from mmlspark_b200.modeltext import parse_model
    assert sum(int(t.get("num_cat", 0)) for t in parse_model(b.model_string())["trees"]) >= 3 * iters
    np.testing.assert_allclose(b.predict_raw(X)[:, 0], h.predict(X), rtol=0, atol=1e-10)


@pytest.mark.parametrize("weighted", [False, True])
def test_binary_logloss_matches_sklearn_hgb(O, weighted):
    rng, X, y = _data(7)
    yb = (y > np.median(y)).astype(np.float32)
    w = rng.integers(1, 5, len(yb)).astype(np.float64) if weighted else None
    h = sk.HistGradientBoostingClassifier(loss="log_loss", learning_rate=0.1, max_iter=20, max_leaf_nodes=31, min_samples_leaf=20, max_bins=255,
                                          early_stopping=False).fit(X, yb, sample_weight=w)
    got = _oracle_pred(O, X, yb, "objective=binary", 20, weight=w)
    np.testing.assert_allclose(got, h.decision_function(X), rtol=0, atol=1e-10)      # incl. the (weighted) log-odds init score


def test_splittable_inheritance_is_the_only_structural_difference(O):
    """Strong L2 makes some features unsplittable in a parent; LightGBM's children then never look at them again, sklearn's do."""
    _, X, y = _data()
    h = _hgb_reg(3, l2_regularization=50.0).fit(X, y.astype(np.float64))
    ref = h.predict(X)
    without_rule = _oracle_pred(O, X, y, "objective=regression lambda_l2=50 oracle_inherit_splittable=false", 3)
    with_rule = _oracle_pred(O, X, y, "objective=regression lambda_l2=50", 3)
    np.testing.assert_allclose(without_rule, ref, rtol=0, atol=1e-10)
    assert np.abs(with_rule - ref).max() > 1e-3          # the LightGBM rule changes this model (HistogramPool flags)


def test_poisson_matches_sklearn_hgb(O):
    """log-link Poisson deviance: same gradients/hessians once LightGBM's hessian safeguard `poisson_max_delta_step` is 0
    (sklearn has none); init score log(mean y) in both."""
    rng, X, _ = _data(9)
    mu = np.exp(0.03 * X[:, 0] - 0.001 * X[:, 1] ** 2 + 0.02 * (X[:, 2] > 3) * X[:, 3])
    y = rng.poisson(mu).astype(np.float32)
    h = sk.HistGradientBoostingRegressor(loss="poisson", learning_rate=0.1, max_iter=15, max_leaf_nodes=31, min_samples_leaf=20, max_bins=255,
                                         early_stopping=False).fit(X, y.astype(np.float64))
    got = _oracle_pred(O, X, y, "objective=poisson poisson_max_delta_step=0", 15)
    np.testing.assert_allclose(got, h._raw_predict(X).ravel(), rtol=0, atol=1e-10)


def test_gamma_matches_sklearn_hgb(O):
    """log-link Gamma deviance: gradient 1 - y exp(-s), hessian y exp(-s), init score log(mean y) in both implementations.
    The hessians vary by orders of magnitude here, which exposes a LightGBM rule sklearn does not have: the row count of a histogram bin
    is RECONSTRUCTED from its hessian mass (cnt = round(h * num_data / sum_hessian), SURVEY.md "Count-from-hessian"), so `min_data_in_leaf` acts on
    hessian-weighted counts.  With the count constraint off (min_data_in_leaf = 0 / min_samples_leaf = 1) the two implementations agree
    to 1e-13; with it on, they differ exactly because of that rule (sklearn splits off a 24-row leaf of low-hessian rows that LightGBM's
    estimate counts as 8)."""
    rng, X, _ = _data(10)
    mu = np.exp(0.02 * X[:, 0] - 0.001 * X[:, 1] ** 2 + 0.02 * (X[:, 2] > 3) * X[:, 3])
    y = rng.gamma(shape=2.0, scale=mu / 2.0).astype(np.float32) + np.float32(1e-3)

    def both(min_leaf, iters):
        h = sk.HistGradientBoostingRegressor(loss="gamma", learning_rate=0.1, max_iter=iters, max_leaf_nodes=31, min_samples_leaf=max(min_leaf, 1),
                                             max_bins=255, early_stopping=False).fit(X, y.astype(np.float64))
        ds = O.OracleDataset(X, "max_bin=255 min_data_in_leaf=%d" % min_leaf).set_field("label", y)
        b = O.OracleBooster(ds, "num_leaves=31 learning_rate=0.1 min_data_in_leaf=%d min_sum_hessian_in_leaf=0.001 verbosity=-1 objective=gamma" % min_leaf)
        b.train(iters)
        return b.predict_raw(X)[:, 0], h._raw_predict(X).ravel()

    got, ref = both(0, 8)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)
    got, ref = both(20, 1)
    assert np.abs(got - ref).max() > 1e-2


def test_multiclass_softmax_matches_sklearn_hgb(O):
    """K trees per iteration on softmax gradients.  LightGBM scales the hessian by K / (K - 1) (MulticlassSoftmax::GetGradients), which
    leaves every split (gain argmax, hessian-share counts) unchanged and multiplies each leaf value by (K - 1) / K: the oracle at learning
    rate 0.1 must equal sklearn at 0.1 (K - 1) / K.  Init scores differ by a per-row constant (log prior vs centred log prior), so the
    comparison is on probabilities.  fp32 gradients in both; the hessian factor is applied in fp32, hence 1e-7 instead of 1e-13.
    Count constraint off for the reason given in the Gamma test."""
    rng, X, _ = _data(12)
    K = 4
    s = np.stack([0.1 * X[:, 0], -0.002 * X[:, 1] ** 2 + 1, 0.15 * (X[:, 2] > 3) * X[:, 3], 0.05 * X[:, 4]], axis=1) + rng.gumbel(size=(len(X), K))
    y = s.argmax(1).astype(np.float32)
    h = sk.HistGradientBoostingClassifier(loss="log_loss", learning_rate=0.1 * (K - 1) / K, max_iter=6, max_leaf_nodes=31, min_samples_leaf=1,
                                          max_bins=255, early_stopping=False).fit(X, y)
    ds = O.OracleDataset(X, "max_bin=255 min_data_in_leaf=0").set_field("label", y)
    b = O.OracleBooster(ds, "num_leaves=31 learning_rate=0.1 min_data_in_leaf=0 min_sum_hessian_in_leaf=0.001 verbosity=-1 objective=multiclass num_class=%d" % K)
    b.train(6)                     # 24 trees; the fp32 rounding of the hessian factor (1e-8 relative) flips a near-tied split around iteration 9
    z = b.predict_raw(X)
    z = z - z.max(axis=1, keepdims=True)
    p = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
    np.testing.assert_allclose(p, h.predict_proba(X), rtol=0, atol=1e-7)


@pytest.mark.parametrize("loss,params,kw,atol", [
    ("absolute_error", "objective=regression_l1", {}, 1e-6),              # the init score is a label_t (float32) percentile in LightGBM
    ("quantile", "objective=quantile alpha=0.8", {"quantile": 0.8}, 1e-5),        # LightGBM keeps alpha as float32
])
def test_percentile_objectives_match_sklearn_hgb(O, loss, params, kw, atol):
    """L1 / quantile: sign-type gradients, percentile init score AND the per-leaf renewal of the outputs (median / quantile of the
    leaf's residuals) agree with sklearn's `_update_leaves_values`; pins PercentileFun's interpolation (fp = (cnt-1)(1-alpha) in the
    descending order, between d[int(fp)] and d[int(fp)+1]) = numpy's linear percentile."""
    _, X, y = _data(9)
    h = sk.HistGradientBoostingRegressor(
        loss=loss, learning_rate=0.1, max_iter=10, max_leaf_nodes=31, min_samples_leaf=20, max_bins=255, early_stopping=False, **kw).fit(X, y.astype(np.float64))
    got = _oracle_pred(O, X, y, params, 10)
    np.testing.assert_allclose(got, h._raw_predict(X).ravel(), rtol=0, atol=atol)


@pytest.mark.parametrize("y,alpha", [([1, 2, 3, 4, 5], 0.5), ([1, 2, 3, 4], 0.5), ([1, 2, 3, 4, 5], 0.8), ([10, 0], 0.25), ([7], 0.3), ([3, 1, 2], 0.999)])
def test_percentile_init_score_known_answers(O, y, alpha):
    yy = np.array(y, dtype=np.float32)
    ds = O.OracleDataset(np.zeros((len(y), 1)), "max_bin=255").set_field("label", yy)
    b = O.OracleBooster(ds, "objective=quantile alpha=%r verbosity=-1" % alpha)
    b.train(1)                                           # no usable feature: a single constant tree = the init score
    want = np.percentile(yy.astype(np.float64), 100 * float(np.float32(alpha)))
    np.testing.assert_allclose(b.predict_raw(np.zeros((1, 1)))[0, 0], want, rtol=1e-6, atol=1e-6)
