"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): bin indices and tree structure bit-exact; leaf values
and split gains within 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DS_PARAMS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"


def _classifier_params(objective="binary", extra="", iters=100, leaves=31, machines=1):
    # byte-for-byte what TrainParams.toString emits (TrainParams.scala:47-63,83-88)
    return ("metric= boost_from_average=true is_pre_partition=True boosting_type=gbdt tree_learner=data_parallel top_k=20 "
            "num_iterations=%d learning_rate=0.1 num_leaves=%d max_bin=255 bagging_fraction=1.0 pos_bagging_fraction=1.0 "
            "neg_bagging_fraction=1.0 bagging_freq=0 bagging_seed=3 early_stopping_round=0 feature_fraction=1.0 max_depth=-1 "
            "min_sum_hessian_in_leaf=0.001 num_machines=%d verbosity=-1 lambda_l1=0.0 lambda_l2=0.0 metric= "
            "min_gain_to_split=0.0 max_delta_step=0.0 min_data_in_leaf=20 objective=%s num_threads=0 %s" % (iters, leaves, machines, objective, extra))


def _without(params, *keys):
    """drops whole key=value tokens (so that 'bagging_fraction' does not also hit 'pos_bagging_fraction')"""
    return " ".join(t for t in params.split(" ") if t.split("=")[0] not in keys)


def _make(X, y, ds_params=DS_PARAMS, weight=None, group=None, init_score=None):
    from mmlspark_b200 import capi
    from oracle import oracle as O
    ds = capi.Dataset.from_mat(X, ds_params)
    ods = O.OracleDataset(X, ds_params)
    for name, arr in (("label", y), ("weight", weight), ("group", group), ("init_score", init_score)):
        if arr is not None:
            ds.set_field(name, arr)
            ods.set_field(name, arr)
    return ds, ods


def _train_both(ds, ods, params, iters):
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model
    from oracle import oracle as O
    b = capi.Booster(ds, params)
    ob = O.OracleBooster(ods, params)
    for it in range(iters):
        f1 = b.update_one_iter()
        f2 = ob.update()
        assert f1 == f2, "is_finished differs at iteration %d" % it
        if f1:
            break
    return b, ob, parse_model(b.save_model_to_string()), parse_model(ob.model_string())


def _feature_matrix(rng, n):
    """Columns that exercise the bin finder: gaussian, many zeros, few distinct values, negatives only,
    constant (trivial), heavy ties, NaNs, tiny magnitudes, integers."""
    cols = [
        rng.standard_normal(n),
        rng.standard_normal(n) * (rng.random(n) < 0.2),
        rng.integers(0, 5, n).astype(np.float64),
        -np.abs(rng.standard_normal(n)) - 0.5,
        np.full(n, 3.25),
        np.round(rng.standard_normal(n), 1),
        np.where(rng.random(n) < 0.1, np.nan, rng.standard_normal(n)),
        rng.standard_normal(n) * 1e-30,
        rng.integers(-50, 50, n).astype(np.float64),
        rng.exponential(2.0, n),
        np.where(rng.random(n) < 0.85, 0.0, rng.random(n)),
        np.where(rng.random(n) < 0.5, np.nan, rng.integers(0, 3, n).astype(np.float64)),
    ]
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("n,max_bin", [(3000, 255), (50000, 255), (20000, 63), (250000, 255)])
def test_bins_bit_exact(built, n, max_bin):
    from mmlspark_b200 import capi
    from oracle import oracle as O
    rng = np.random.default_rng(n + max_bin)
    X = _feature_matrix(rng, n)
    params = DS_PARAMS.replace("max_bin=255", "max_bin=%d" % max_bin)
    ds = capi.Dataset.from_mat(X, params)
    ods = O.OracleDataset(X, params)
    for f in range(X.shape[1]):
        assert ds.feature_info(f) == ods.feature_info(f), "feature %d meta differs" % f
        ub, oub = ds.upper_bounds(f), ods.upper_bounds(f)
        assert ub.tobytes() == oub.tobytes(), "feature %d bin upper bounds differ (bitwise)" % f   # NaN-safe bitwise compare
    assert np.array_equal(ds.get_bins(), ods.bins())
    # float32 column-major input goes through the same kernel
    ds32 = capi.Dataset.from_mat(X.astype(np.float32), params, row_major=False)
    ods32 = O.OracleDataset(X.astype(np.float32).astype(np.float64), params)
    assert np.array_equal(ds32.get_bins(), ods32.bins())
    for d in (ds, ds32):
        d.free()


def test_reference_dataset_uses_train_bins(built):
    from mmlspark_b200 import capi
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    X = _feature_matrix(rng, 20000)
    Xv = _feature_matrix(rng, 5000)
    ds = capi.Dataset.from_mat(X, DS_PARAMS)
    dv = capi.Dataset.from_mat(Xv, DS_PARAMS, reference=ds)
    ods = O.OracleDataset(X, DS_PARAMS)
    # bin the validation rows with the training mappers on the host
    want = np.zeros(Xv.shape, dtype=np.uint8)
    for f in range(X.shape[1]):
        info = ods.feature_info(f)
        if info["is_trivial"]:
            continue
        ub = ods.upper_bounds(f)
        nb = info["num_bin"] - (1 if info["missing_type"] == 2 else 0)
        v = Xv[:, f].copy()
        nan = np.isnan(v)
        v[nan] = 0.0
        b = np.searchsorted(ub[:nb - 1], v, side="left")
        if info["missing_type"] == 2:
            b[nan] = info["num_bin"] - 1
        want[:, f] = b
    assert np.array_equal(dv.get_bins(), want)


@pytest.mark.parametrize("n,F", [(1000, 7), (40000, 70), (300000, 33)])
def test_histogram_kernel_vs_oracle(built, n, F):
    """K4 (fixed-point shared-memory atomics) against the fp64 oracle histogram: root pass and gathered leaves."""
    from mmlspark_b200 import capi
    from oracle import oracle as O
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, F))
    X[:, 1] = np.where(rng.random(n) < 0.3, np.nan, X[:, 1])
    ds = capi.Dataset.from_mat(X, DS_PARAMS)
    bins = ds.get_bins()
    g = (rng.standard_normal(n) * np.exp(rng.standard_normal(n) * 3)).astype(np.float32)   # wide dynamic range
    h = rng.random(n).astype(np.float32)
    for idx in (None, np.sort(rng.choice(n, n // 3, replace=False)).astype(np.int32), np.arange(0, min(n, 37), dtype=np.int32), np.zeros(0, dtype=np.int32)):
        got = ds.histogram(g, h, idx)
        want = O.histogram(bins, g, h, idx)
        cnt = n if idx is None else len(idx)
        # fixed point: every element is rounded to 2^-35 of the maximum magnitude => |err| <= cnt * 2^-36 * max
        tol_g = max(cnt, 1) * 2.0 ** -35 * float(np.abs(g).max())
        tol_h = max(cnt, 1) * 2.0 ** -35 * float(np.abs(h).max())
        assert np.abs(got[:, :, 0] - want[:, :, 0]).max() <= tol_g
        assert np.abs(got[:, :, 1] - want[:, :, 1]).max() <= tol_h
        np.testing.assert_allclose(got.sum(axis=1), want.sum(axis=1), rtol=1e-9, atol=tol_g)
    ds.free()


def test_regression_tree_sequence_identical(built):
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(1)
    n, F = 60000, 50
    X = rng.standard_normal((n, F))
    y = (2 * X[:, 0] + np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + 0.3 * rng.standard_normal(n)).astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params("regression", "alpha=0.9 tweedie_variance_power=1.5").replace("is_unbalance=false", "")
    b, ob, m, om = _train_both(ds, ods, params, 30)
    compare_models(m, om)
    assert len(m["trees"]) == 30
    np.testing.assert_allclose(b.get_scores(), ob.scores(), rtol=1e-7, atol=1e-7)


def test_binary_config1_100_iters(built):
    """BASELINE config 1: LightGBMClassifier binary, synthetic 50k x 28 dense, 100 iterations."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(42)
    n, F = 50000, 28
    X = rng.standard_normal((n, F))
    w = np.random.default_rng(7).standard_normal(F)
    y = (X @ w + 0.5 * rng.standard_normal(n) > 0).astype(np.float32)
    ds, ods = _make(X, y)
    b, ob, m, om = _train_both(ds, ods, _classifier_params("binary", "is_unbalance=false"), 100)
    compare_models(m, om)
    assert len(m["trees"]) == 100
    p = b.predict_for_mat(X[:2000], predict_type=1)[:, 0]
    np.testing.assert_allclose(p, ob.predict_raw(X[:2000])[:, 0], rtol=1e-6, atol=1e-6)


def test_binary_weighted_unbalanced_regularised(built):
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(3)
    n, F = 30000, 20
    X = rng.standard_normal((n, F))
    y = (X[:, 0] + X[:, 1] ** 2 + rng.standard_normal(n) > 1.5).astype(np.float32)
    wt = rng.random(n).astype(np.float32) + 0.5
    ds, ods = _make(X, y, weight=wt)
    params = _classifier_params("binary", "is_unbalance=true", leaves=15).replace("lambda_l1=0.0", "lambda_l1=0.5").replace(
        "lambda_l2=0.0", "lambda_l2=2.0").replace("min_gain_to_split=0.0", "min_gain_to_split=0.1").replace("max_delta_step=0.0", "max_delta_step=0.7")
    b, ob, m, om = _train_both(ds, ods, params, 25)
    compare_models(m, om)


def test_missing_values_two_way_scan(built):
    """NaN features exercise the left-to-right + right-to-left scan and default_left."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(9)
    n, F = 40000, 12
    X = rng.standard_normal((n, F))
    for f in (0, 3, 5):
        X[rng.random(n) < 0.25, f] = np.nan
    X[:, 7] = np.where(rng.random(n) < 0.6, 0.0, X[:, 7])          # most_freq_bin == default bin (offset path)
    y = (np.where(np.isnan(X[:, 0]), 1.5, X[:, 0]) + np.nan_to_num(X[:, 3]) * 0.5 + X[:, 7] + 0.2 * rng.standard_normal(n)).astype(np.float32)
    ds, ods = _make(X, y)
    b, ob, m, om = _train_both(ds, ods, _classifier_params("regression", ""), 20)
    compare_models(m, om)
    dts = np.concatenate([t["decision_type"] for t in m["trees"] if t["num_leaves"] > 1])
    assert (dts >= 8).any(), "expected at least one split on a NaN-missing feature"


def test_max_depth_and_small_leaves(built):
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(11)
    n, F = 5000, 10
    X = rng.standard_normal((n, F))
    y = (X[:, 0] > 0).astype(np.float32) + 0.1 * rng.standard_normal(n).astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params("regression", "", leaves=63).replace("max_depth=-1", "max_depth=4").replace("min_data_in_leaf=20", "min_data_in_leaf=200")
    b, ob, m, om = _train_both(ds, ods, params, 15)
    compare_models(m, om)


def test_multiclass_softmax(built):
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(21)
    n, F, K = 30000, 16, 4
    X = rng.standard_normal((n, F))
    W = rng.standard_normal((F, K))
    y = np.argmax(X @ W + rng.standard_normal((n, K)), axis=1).astype(np.float32)
    ds, ods = _make(X, y)
    b, ob, m, om = _train_both(ds, ods, _classifier_params("multiclass", "num_class=%d" % K, leaves=15), 10)
    compare_models(m, om)
    assert len(m["trees"]) == 10 * K
    prob = b.predict_for_mat(X[:100])
    np.testing.assert_allclose(prob.sum(axis=1), 1.0, atol=1e-9)          # VerifyLightGBMClassifier.scala:91-99


def test_lambdarank(built):
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(31)
    sizes = rng.integers(5, 40, 600).astype(np.int32)
    n, F = int(sizes.sum()), 12
    X = rng.standard_normal((n, F))
    rel = np.clip(np.round(X[:, 0] + 0.5 * X[:, 1] + rng.standard_normal(n) * 0.5 + 1.5), 0, 4).astype(np.float32)
    ds, ods = _make(X, rel, group=sizes)
    params = _classifier_params("lambdarank", "max_position=20 eval_at=1,2,3,4,5", leaves=15).replace("boost_from_average=true", "")
    b, ob, m, om = _train_both(ds, ods, params, 10)
    # per-document lambdas are accumulated by one thread per document in the reference's own pair order (same fp32 additions), so the
    # general 1e-5 bar holds (round 1 needed 1e-4 with shared-memory float atomics)
    compare_models(m, om)
    np.testing.assert_allclose(b.get_scores(), ob.scores(), rtol=1e-7, atol=1e-9)


def test_custom_objective_and_init_score(built):
    """LGBM_BoosterUpdateOneIterCustom (fobj path, LightGBMBooster.scala:368-388) equals the built-in L2 objective
    when fed the same gradients; init_score shifts the starting point."""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    rng = np.random.default_rng(41)
    n, F = 20000, 8
    X = rng.standard_normal((n, F))
    y = (X[:, 0] - X[:, 1] + 0.1 * rng.standard_normal(n)).astype(np.float32)
    init = np.full(n, 0.25)
    ds, ods = _make(X, y, init_score=init)
    params = _classifier_params("regression", "")
    b1, ob, m1, om = _train_both(ds, ods, params, 5)
    compare_models(m1, om)
    b2 = capi.Booster(ds, params)
    for _ in range(5):
        s = b2.get_scores()
        b2.update_one_iter_custom((s - y).astype(np.float32), np.ones(n, dtype=np.float32))
    compare_models(parse_model(b2.save_model_to_string()), m1)


def test_finishes_when_no_split_possible(built):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((30, 3))
    y = rng.standard_normal(30).astype(np.float32)
    ds, ods = _make(X, y)
    b, ob, m, om = _train_both(ds, ods, _classifier_params("regression", ""), 3)
    assert len(m["trees"]) == len(om["trees"]) == 1 and m["trees"][0]["num_leaves"] == 1
    np.testing.assert_allclose(m["trees"][0]["leaf_value"], om["trees"][0]["leaf_value"], rtol=1e-12)


def test_determinism_run_to_run(built):
    """Integer (fixed-point) accumulation makes the engine bit-reproducible despite atomics."""
    from mmlspark_b200 import capi
    rng = np.random.default_rng(77)
    n, F = 100000, 40
    X = rng.standard_normal((n, F))
    y = (X[:, 0] * X[:, 1] + rng.standard_normal(n)).astype(np.float32)
    ds, _ = _make(X, y)
    strs = []
    for _ in range(2):
        b = capi.Booster(ds, _classifier_params("regression", ""))
        for _ in range(5):
            b.update_one_iter()
        strs.append(b.save_model_to_string())
        b.free()
    assert strs[0] == strs[1]


@pytest.mark.parametrize("objective", ["huber", "fair", "poisson", "gamma", "tweedie"])
def test_regression_objective_variants(built, objective):
    """LightGBMRegressor.objective values without leaf renewal (LightGBMRegressor.scala:46, TrainParams.scala:165-179)."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(51)
    n, F = 30000, 14
    X = rng.standard_normal((n, F))
    mu = 0.8 * X[:, 0] + 0.5 * np.sin(2 * X[:, 1]) + 0.3 * X[:, 2] * X[:, 3]
    if objective in ("poisson", "tweedie"):
        y = rng.poisson(np.exp(mu)).astype(np.float32)
    elif objective == "gamma":
        y = (rng.gamma(2.0, np.exp(mu) / 2.0) + 1e-3).astype(np.float32)
    else:
        y = (mu + 0.3 * rng.standard_t(3, n)).astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params(objective, "alpha=0.9 tweedie_variance_power=1.5", leaves=15)
    b, ob, m, om = _train_both(ds, ods, params, 12)
    compare_models(m, om)
    assert m["header"]["objective"] == objective
    raw = b.predict_for_mat(X[:50], predict_type=1)[:, 0]
    norm = b.predict_for_mat(X[:50])[:, 0]
    if objective in ("poisson", "gamma", "tweedie"):
        np.testing.assert_allclose(norm, np.exp(raw), rtol=1e-12)          # ConvertOutput = exp
    else:
        np.testing.assert_array_equal(norm, raw)
    ev = b.get_eval(0)
    assert b.eval_names() == [objective] and np.isfinite(ev).all()


def test_unsupported_objective_fails_loudly(built):
    from mmlspark_b200 import capi
    rng = np.random.default_rng(0)
    X = rng.standard_normal((500, 3)); y = rng.standard_normal(500).astype(np.float32)
    ds, _ = _make(X, y)
    for obj in ("cross_entropy", "rank_xendcg", "multiclassova", "not_an_objective"):
        with pytest.raises(capi.LightGBMError):
            capi.Booster(ds, "objective=%s" % obj)
    for bad in ("boosting_type=not_a_booster", "boosting_type=goss bagging_fraction=0.5 bagging_freq=1"):
        with pytest.raises(capi.LightGBMError):
            capi.Booster(ds, "objective=regression " + bad)


def test_feature_fraction_column_sampling(built):
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(61)
    n, F = 20000, 24
    X = rng.standard_normal((n, F))
    y = (X[:, :6].sum(axis=1) + 0.5 * rng.standard_normal(n)).astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params("regression", "", leaves=15).replace("feature_fraction=1.0", "feature_fraction=0.4")
    b, ob, m, om = _train_both(ds, ods, params, 10)
    compare_models(m, om)
    used = [set(t["split_feature"].tolist()) for t in m["trees"]]
    assert all(len(u) <= 10 for u in used) and len(set().union(*used)) > 10      # <= round(24*0.4) features per tree, different per tree


def test_batched_gpu_prediction_bit_exact_with_host_predictor(built):
    """B200GBM_BoosterPredictForMatDevice against the host single-row predictor the reference's UDFs use."""
    import json
    import os
    from mmlspark_b200 import capi
    rng = np.random.default_rng(71)
    n, F = 50000, 20
    X = rng.standard_normal((n, F))
    X[rng.random((n, F)) < 0.05] = np.nan
    y = (np.nan_to_num(X[:, 0]) + np.nan_to_num(X[:, 1]) ** 2 + 0.3 * rng.standard_normal(n) > 0.8).astype(np.float32)
    ds, _ = _make(X, y)
    b = capi.Booster(ds, _classifier_params("binary", "is_unbalance=false", leaves=31))
    for _ in range(20):
        b.update_one_iter()
    for pt in (capi.PREDICT_RAW_SCORE, capi.PREDICT_NORMAL, capi.PREDICT_LEAF_INDEX):
        dev = b.predict_device(X, pt)
        host = b.predict_for_mat(X, pt)
        assert dev.shape == host.shape
        np.testing.assert_array_equal(dev, host)
    np.testing.assert_array_equal(b.predict_device(X, capi.PREDICT_RAW_SCORE, 3, 5), b.predict_for_mat(X, capi.PREDICT_RAW_SCORE, 3, 5))
    # float32 input and a prediction-only booster loaded from the model text
    b2 = capi.Booster(model_str=b.save_model_to_string())
    X32 = X.astype(np.float32)
    np.testing.assert_array_equal(b2.predict_device(X32, capi.PREDICT_RAW_SCORE), b.predict_for_mat(X32.astype(np.float64), capi.PREDICT_RAW_SCORE))
    # multiclass golden model
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))["models"]["multiclass"]["model"]
    mc = capi.Booster(model_str=g)
    Xm = rng.standard_normal((3000, 10))
    np.testing.assert_array_equal(mc.predict_device(Xm, capi.PREDICT_RAW_SCORE), mc.predict_for_mat(Xm, capi.PREDICT_RAW_SCORE))
    np.testing.assert_array_equal(mc.predict_device(Xm), mc.predict_for_mat(Xm))
    big = np.tile(X32, (20, 1))
    out, ms = b.predict_device(big, capi.PREDICT_RAW_SCORE, return_ms=True)
    print("GPU batch predict: %d rows x %d feats x 20 trees in %.2f ms (incl. H2D/D2H) = %.1f Mrows/s" % (big.shape[0], F, ms, big.shape[0] / ms / 1e3))


def _categorical_matrix(rng, n):
    c12 = rng.integers(0, 12, n).astype(np.float64)                               # many-vs-many
    c3 = rng.integers(0, 3, n).astype(np.float64)                                 # one-hot (num_bin <= 4)
    c3[rng.random(n) < 0.05] = np.nan
    zipf = np.minimum(rng.zipf(1.5, n), 400).astype(np.float64) * 3 - 2           # ~150 categories, long tail, a negative value
    x = rng.standard_normal((n, 3))
    eff12 = np.array([0.5, -1, 2, 0.1, -0.3, 1.5, -2, 0, 0.7, -0.9, 1.1, -1.4])
    s = eff12[c12.astype(int)] + np.where(np.nan_to_num(c3) == 1, 1.0, 0.0) + 0.8 * np.sin(zipf) + x[:, 0] + 0.3 * rng.standard_normal(n)
    X = np.column_stack([c12, x[:, 0], c3, x[:, 1], zipf, x[:, 2]])
    return X, s


@pytest.mark.parametrize("objective", ["regression", "binary"])
def test_categorical_features(built, objective):
    """categoricalSlotIndexes -> categorical_feature=... (LightGBMBase.scala:168-199, 265-272; 'num_cat=' check VerifyLightGBMClassifier.scala:463-495):
    categorical bin finder, one-hot and many-vs-many split search, bitset splits in partition / model text / predictors."""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import compare_models
    from oracle import oracle as O
    rng = np.random.default_rng(81)
    n = 40000
    X, s = _categorical_matrix(rng, n)
    y = (s > 0.3).astype(np.float32) if objective == "binary" else s.astype(np.float32)
    dsp = DS_PARAMS + " categorical_feature=0,2,4"
    ds, ods = _make(X, y, ds_params=dsp)
    for f in range(X.shape[1]):
        assert ds.feature_info(f) == ods.feature_info(f)
    assert np.array_equal(ds.get_bins(), ods.bins())
    params = _classifier_params(objective, "categorical_feature=0,2,4" + (" is_unbalance=false" if objective == "binary" else ""), leaves=15)
    b, ob, m, om = _train_both(ds, ods, params, 15)
    compare_models(m, om)
    assert sum(t.get("num_cat", 0) for t in m["trees"]) > 5 and "num_cat=" in b.save_model_to_string()
    dts = np.concatenate([t["decision_type"] for t in m["trees"]])
    assert (dts & 1).any() and ((dts & 1) == 0).any()
    # predictors agree: host single/batch, device batch, oracle, and the training scores
    raw_host = b.predict_for_mat(X[:3000], capi.PREDICT_RAW_SCORE)[:, 0]
    np.testing.assert_array_equal(b.predict_device(X[:3000], capi.PREDICT_RAW_SCORE)[:, 0], raw_host)
    np.testing.assert_allclose(raw_host, ob.predict_raw(X[:3000])[:, 0], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(raw_host, b.get_scores()[:3000], rtol=1e-9, atol=1e-9)
    b2 = capi.Booster(model_str=b.save_model_to_string())              # model text round trip keeps cat_boundaries / cat_threshold
    np.testing.assert_array_equal(b2.predict_for_mat(X[:3000], capi.PREDICT_RAW_SCORE)[:, 0], raw_host)
    unseen = X[:5].copy(); unseen[:, 0] = 999; unseen[:, 4] = -7        # unseen / negative categories go right
    np.testing.assert_allclose(b.predict_for_mat(unseen, capi.PREDICT_RAW_SCORE), ob.predict_raw(unseen), rtol=1e-6, atol=1e-6)


def test_categorical_too_many_bins_fails_loudly(built):
    """features beyond 256 bins (categorical, or numerical with max_bin > 255) are supported up to 16384 bins (tests/test_gpu_wide.py);
    past that: -1 with a clear message"""
    from mmlspark_b200 import capi
    rng = np.random.default_rng(0)
    X = np.column_stack([rng.integers(0, 30000, 190000).astype(np.float64), rng.standard_normal(190000)])      # ~6 rows per category
    with pytest.raises(capi.LightGBMError) as e:
        capi.Dataset.from_mat(X, DS_PARAMS + " categorical_feature=0")
    assert "16384" in str(e.value)
    with pytest.raises(capi.LightGBMError) as e:
        capi.Dataset.from_mat(X, DS_PARAMS.replace("max_bin=255", "max_bin=20000"))
    assert "max_bin" in str(e.value)


def _sampling_case(rng, n=30000, F=12):
    X = rng.standard_normal((n, F))
    X[:, 3] = np.where(rng.random(n) < 0.1, np.nan, X[:, 3])
    s = 1.2 * X[:, 0] + np.sin(2 * X[:, 1]) + X[:, 2] * np.nan_to_num(X[:, 3]) + 0.4 * rng.standard_normal(n)
    return X, s


@pytest.mark.parametrize("objective,extra", [
    ("regression", "bagging_fraction=0.6 bagging_freq=2"),
    ("binary", "bagging_fraction=0.35 bagging_freq=1 bagging_seed=11 is_unbalance=false"),
    ("multiclass", "num_class=3 bagging_fraction=0.8 bagging_freq=3"),
    ("binary", "pos_bagging_fraction=0.9 neg_bagging_fraction=0.3 bagging_freq=1 is_unbalance=false"),       # balanced bagging
])
def test_bagging(built, objective, extra):
    """Row bagging: per-1024-row-block LCG draws, the in-bag list is the tree's root, out-of-bag rows are scored
    by the binned tree walk.  n is not a multiple of 1024 so the last block is partial."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(81)
    X, s = _sampling_case(rng)
    y = {"regression": s, "binary": s > 0, "multiclass": np.digitize(s, [-0.8, 0.8])}[objective].astype(np.float32)
    ds, ods = _make(X, y)
    params = _without(_classifier_params(objective, "", leaves=15), "bagging_fraction", "pos_bagging_fraction", "neg_bagging_fraction", "bagging_freq", "bagging_seed")
    params += " " + extra
    b, ob, m, om = _train_both(ds, ods, params, 9)
    compare_models(m, om)
    assert len(m["trees"]) == 9 * (3 if objective == "multiclass" else 1)
    np.testing.assert_allclose(b.get_scores(0).ravel(), ob.scores().ravel(), rtol=0, atol=1e-9)
    root_counts = [int(t["internal_count"][0]) for t in m["trees"]]
    assert max(root_counts) < 0.95 * len(y)        # trees really saw a subsample


@pytest.mark.parametrize("objective,extra", [
    ("binary", "bagging_fraction=0.7 bagging_freq=1 is_unbalance=false"),
    ("regression", "feature_fraction=0.5"),                       # rf without row bagging: column sampling only
    ("multiclass", "num_class=3 bagging_fraction=0.5 bagging_freq=1 feature_fraction=0.8"),
])
def test_random_forest(built, objective, extra):
    """boosting_type=rf: gradients fixed at the init score, no shrinkage, average_output model, running-average scores."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(83)
    X, s = _sampling_case(rng, n=20000)
    y = {"regression": s + 3.0, "binary": s > 0.5, "multiclass": np.digitize(s, [-0.8, 0.8])}[objective].astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params(objective, "", leaves=15).replace("boosting_type=gbdt", "boosting_type=rf")
    params = _without(params, "bagging_fraction", "bagging_freq", "feature_fraction") + " " + extra
    b, ob, m, om = _train_both(ds, ods, params, 8)
    compare_models(m, om)
    text = b.save_model_to_string()
    assert "\naverage_output\n" in text
    np.testing.assert_allclose(b.get_scores(0).ravel(), ob.scores().ravel(), rtol=0, atol=1e-9)
    # the model's prediction is the average over the iterations == the training score
    Xs = X[:2000]
    raw = b.predict_for_mat(Xs, predict_type=1).reshape(len(Xs), -1)
    K = raw.shape[1]
    np.testing.assert_allclose(raw, b.get_scores(0).reshape(K, -1)[:, :2000].T, rtol=0, atol=1e-9)
    np.testing.assert_allclose(b.predict_device(Xs, predict_type=1).reshape(len(Xs), -1), raw, rtol=0, atol=0)


def test_random_forest_needs_subsampling(built):
    from mmlspark_b200 import capi
    rng = np.random.default_rng(5)
    X = rng.standard_normal((2000, 4))
    ds, _ = _make(X, X[:, 0].astype(np.float32))
    with pytest.raises(capi.LightGBMError, match="bagging_freq"):
        capi.Booster(ds, "objective=regression boosting_type=rf verbosity=-1")


@pytest.mark.parametrize("objective,extra", [
    ("binary", "learning_rate=0.25 is_unbalance=false"),
    ("multiclass", "num_class=3 learning_rate=0.5 top_rate=0.3 other_rate=0.2"),
    ("regression", "learning_rate=0.34 top_rate=0.1 other_rate=0.05"),
])
def test_goss(built, objective, extra):
    """boosting_type=goss: full data for the first 1/learning_rate iterations, then top_rate by |g*h| + other_rate sampled with
    amplified gradients, per 1024-row chunk."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(85)
    X, s = _sampling_case(rng, n=25000)
    # no NaN column here: in a leaf whose subsample holds no NaN row the two scan directions have mathematically EQUAL gains, and
    # with float-amplified gradients the winner (default_left) is decided by summation-order rounding noise (ties)
    X[:, 3] = np.nan_to_num(X[:, 3])
    y = {"regression": s, "binary": s > 0, "multiclass": np.digitize(s, [-0.8, 0.8])}[objective].astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params(objective, "", leaves=15).replace("boosting_type=gbdt", "boosting_type=goss").replace("learning_rate=0.1 ", "") + " " + extra
    b, ob, m, om = _train_both(ds, ods, params, 9)
    compare_models(m, om)
    np.testing.assert_allclose(b.get_scores(0).ravel(), ob.scores().ravel(), rtol=0, atol=1e-9)
    K = 3 if objective == "multiclass" else 1
    root_counts = [int(t["internal_count"][0]) for t in m["trees"]]
    assert root_counts[0] == len(y) and root_counts[-1] < 0.6 * len(y)       # warm-up on all rows, then the GOSS subsample


def test_cuda_path_reproduces_committed_golden_models(built):
    """The committed fixtures (tests/golden/oracle_golden.json, generated by make_golden.py) are reproduced by the CUDA path
    without running the oracle: regression / binary / multiclass / bagging / rf / goss / categorical, 5 iterations each."""
    import importlib.util
    import json
    import os
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_data", os.path.join(here, "golden", "make_golden.py"))
    golden = json.load(open(os.path.join(here, "golden", "oracle_golden.json")))["models"]
    # make_golden imports the oracle at module level only to GENERATE; here just its seeded dataset() is used
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    X, s = mg.dataset(123, 2000, 10)
    labels = {"regression": s.astype(np.float32), "binary": (s > 0).astype(np.float32), "multiclass": np.clip(np.floor(s + 1.5), 0, 2).astype(np.float32)}
    assert len(golden) >= 7
    for name, g in golden.items():
        Xc = mg.variant(X, name)
        ds = capi.Dataset.from_mat(Xc, DS_PARAMS + (" " + name.split("|")[1] if "|" in name else ""))
        ds.set_field("label", labels[name.split("_")[0].split("|")[0]])
        b = capi.Booster(ds, g["params"])
        for _ in range(5):
            b.update_one_iter()
        compare_models(parse_model(b.save_model_to_string()), parse_model(g["model"]))
        raw = b.predict_for_mat(Xc[:8], predict_type=1).reshape(8, -1)
        np.testing.assert_allclose(raw, np.array(g["raw_pred_first8"]).reshape(8, -1), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("objective,extra", [
    ("regression", ""),                                                              # native defaults: drop_rate 0.1, skip_drop 0.5, max_drop 50
    ("binary", "uniform_drop=true drop_rate=0.3 skip_drop=0.2 is_unbalance=false"),
    ("binary", "xgboost_dart_mode=true drop_rate=0.5 skip_drop=0.0 max_drop=3 is_unbalance=false"),
    ("multiclass", "num_class=3 drop_rate=0.4 skip_drop=0.1 bagging_fraction=0.7 bagging_freq=1"),
])
def test_dart(built, objective, extra):
    """boosting_type=dart: dropped trees are negated and re-applied to the binned training rows before the gradients are taken,
    the new tree is shrunk by lr/(1+k), then the dropped trees are re-normalised (their stored leaf values change in the model)."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(87)
    X, s = _sampling_case(rng, n=20000)
    y = {"regression": s, "binary": s > 0, "multiclass": np.digitize(s, [-0.8, 0.8])}[objective].astype(np.float32)
    ds, ods = _make(X, y)
    params = _classifier_params(objective, "", leaves=15).replace("boosting_type=gbdt", "boosting_type=dart")
    if "bagging_fraction" in extra:
        params = _without(params, "bagging_fraction", "bagging_freq")
    params += " " + extra
    b, ob, m, om = _train_both(ds, ods, params, 25)
    compare_models(m, om)
    np.testing.assert_allclose(b.get_scores(0).ravel(), ob.scores().ravel(), rtol=0, atol=1e-8)
    shr = np.array([float(t["shrinkage"]) for t in m["trees"]])
    assert len(np.unique(np.round(shr, 12))) > 2        # trees were dropped and re-normalised
    # model prediction == training score (drops are fully accounted for), host and device predictors
    Xs = X[:3000]
    raw = b.predict_for_mat(Xs, predict_type=1).reshape(len(Xs), -1)
    K = raw.shape[1]
    np.testing.assert_allclose(raw, b.get_scores(0).reshape(K, -1)[:, :3000].T, rtol=0, atol=1e-8)
    np.testing.assert_allclose(b.predict_device(Xs, predict_type=1).reshape(len(Xs), -1), raw, rtol=0, atol=0)


@pytest.mark.parametrize("objective,extra,weighted", [
    ("regression_l1", "", False),
    ("quantile", "alpha=0.8", False),
    ("quantile", "alpha=0.2", True),
    ("mape", "", False),
    ("regression_l1", "bagging_fraction=0.7 bagging_freq=1", True),
])
def test_percentile_objectives_renew_leaf_outputs(built, objective, extra, weighted):
    """regression_l1 / quantile / mape: sign-type gradients, init score = (weighted) percentile of the labels and, after every
    tree, leaf outputs re-fitted as the percentile of the leaf's residuals (SURVEY §8 a9 RenewTreeOutput) — device sort path."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(91)
    n, F = 30000, 8
    X = rng.standard_normal((n, F))
    y = (X[:, 0] + 0.5 * X[:, 1] + (1 + 0.5 * np.abs(X[:, 2])) * rng.standard_normal(n)).astype(np.float32)
    if objective == "mape":
        y = (y + 6).astype(np.float32)
    w = rng.uniform(0.5, 2.0, n).astype(np.float32) if weighted else None
    ds, ods = _make(X, y, weight=w)
    params = _classifier_params(objective, "", leaves=15).replace("learning_rate=0.1", "learning_rate=0.2")
    if "bagging" in extra:
        params = _without(params, "bagging_fraction", "bagging_freq")
    params += " " + extra
    b, ob, m, om = _train_both(ds, ods, params, 12)
    compare_models(m, om)
    assert m["header"]["objective"] == objective
    np.testing.assert_allclose(b.get_scores(0).ravel(), ob.scores().ravel(), rtol=0, atol=1e-9)
    p = b.predict_for_mat(X, predict_type=1).ravel()
    if objective == "quantile":
        a = float(extra.split("=")[1])
        ww = w if w is not None else np.ones(n)
        assert abs(float(np.sum(ww * (y < p)) / np.sum(ww)) - a) < 0.03          # calibrated quantile
    names = b.eval_names()
    assert names == [{"regression_l1": "l1"}.get(objective, objective)]
    assert np.isfinite(b.get_eval(0)[0])


@pytest.mark.parametrize("n,F,case", [
    (45, 3, "tiny"),                 # fewer rows than a warp; min_data_in_leaf=20 leaves room for exactly one split
    (1000, 1, "one_feature"),
    (5000, 33, "tile_boundary"),     # 33 used features = 2 tiles, the second holding one feature
    (5000, 40, "trivial_columns"),   # constant / all-NaN / all-zero columns are dropped by the bin finder (used-feature map)
    (4000, 6, "duplicates"),         # heavy ties: few distinct values per feature, duplicated rows
    (1537, 5, "ragged_blocks"),      # not a multiple of 32 / 512 / 1024 / 2048 (stage, bagging block, partition chunk sizes)
])
def test_edge_shapes(built, n, F, case):
    """Ragged and degenerate inputs through the whole path (binning, K4 staging tails, partition chunks, model text)."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(1000 + n)
    X = rng.standard_normal((n, F))
    if case == "trivial_columns":
        X[:, 3] = 7.5; X[:, 10] = np.nan; X[:, 17] = 0.0; X[:, 25] = np.where(rng.random(n) < 0.999, 0.0, 1.0)
    if case == "duplicates":
        X = np.round(X[rng.integers(0, 200, n)] * 2) / 2
    y = (X[:, 0] + (np.nan_to_num(X[:, min(1, F - 1)]) > 0.3) + 0.2 * rng.standard_normal(n)).astype(np.float32)
    ds, ods = _make(X, y)
    assert np.array_equal(ds.get_bins(), ods.bins())
    b, ob, m, om = _train_both(ds, ods, _classifier_params("regression", "", leaves=15), 6)
    compare_models(m, om)
    np.testing.assert_allclose(b.get_scores(0), ob.scores(), rtol=0, atol=1e-9)
    # bagging on a ragged row count exercises the partial last LCG block
    b2, ob2, m2, om2 = _train_both(ds, ods, _without(_classifier_params("regression", "", leaves=7), "bagging_fraction", "bagging_freq") + " bagging_fraction=0.7 bagging_freq=1", 4)
    compare_models(m2, om2)


def test_no_usable_feature_gives_constant_model(built):
    """every column trivial => no tree can be grown: one constant tree holding the label mean, is_finished on the first call."""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model
    X = np.ones((500, 4)); X[:, 1] = np.nan
    y = np.linspace(0, 1, 500).astype(np.float32)
    ds, ods = _make(X, y)
    b = capi.Booster(ds, "objective=regression verbosity=-1")
    from oracle import oracle as O
    ob = O.OracleBooster(ods, "objective=regression verbosity=-1")
    assert b.update_one_iter() == ob.update()
    m, om = parse_model(b.save_model_to_string()), parse_model(ob.model_string())
    assert len(m["trees"]) == len(om["trees"])
    if m["trees"]:
        np.testing.assert_allclose(m["trees"][0]["leaf_value"], om["trees"][0]["leaf_value"], rtol=1e-12)
    np.testing.assert_allclose(b.predict_for_mat(X[:3], predict_type=1).ravel(), float(np.mean(y)), rtol=1e-6)


def test_zero_and_extreme_weights_and_gradients(built):
    """weights spanning 12 orders of magnitude and zero-weight rows: the per-tree fixed-point scale follows max|g| and the
    reconstructed counts follow the hessians, as in the oracle."""
    from mmlspark_b200.modeltext import compare_models
    rng = np.random.default_rng(77)
    n, F = 20000, 6
    X = rng.standard_normal((n, F))
    y = (X[:, 0] - X[:, 1] + 0.3 * rng.standard_normal(n) > 0).astype(np.float32)
    w = np.exp(rng.uniform(-6, 6, n)).astype(np.float32)
    w[rng.random(n) < 0.05] = 0.0
    ds, ods = _make(X, y, weight=w)
    b, ob, m, om = _train_both(ds, ods, _classifier_params("binary", "is_unbalance=false min_sum_hessian_in_leaf=1e-3", leaves=15), 8)
    compare_models(m, om, check_counts=True)


@pytest.mark.parametrize("case", ["regression_nan", "binary", "categorical_weights", "categorical_effects"])
def test_cuda_path_matches_sklearn_hist_gradient_boosting(built, case):
    """The CUDA path against an INDEPENDENT implementation (scikit-learn's HistGradientBoosting), no oracle involved: on
    integer-valued features both bin finders put one bin per distinct value, so the fitted models must be the same function
    (tests/test_oracle_vs_sklearn_cpu.py explains the construction).  Tolerance 1e-6: the north star's leaf-value bar."""
    sk = pytest.importorskip("sklearn.ensemble")
    from mmlspark_b200 import capi
    rng = np.random.default_rng(16)
    n, F = 20000, 8
    X = rng.integers(-20, 30, size=(n, F)).astype(np.float64)
    y = (0.3 * X[:, 0] - 0.02 * X[:, 1] ** 2 + 0.5 * (X[:, 2] > 3) * X[:, 3] + rng.standard_normal(n)).astype(np.float32)
    base = "num_leaves=31 learning_rate=0.1 min_data_in_leaf=20 min_sum_hessian_in_leaf=0.001 verbosity=-1 "
    common = dict(learning_rate=0.1, max_iter=15, max_leaf_nodes=31, min_samples_leaf=20, max_bins=255, early_stopping=False)
    w, ds_params = None, DS_PARAMS
    if case == "regression_nan":
        X[rng.random(X.shape) < 0.1] = np.nan
        h = sk.HistGradientBoostingRegressor(loss="squared_error", **common).fit(X, y.astype(np.float64))
        params, want = base + "objective=regression", h.predict(X)
    elif case == "binary":
        y = (y > np.median(y)).astype(np.float32)
        h = sk.HistGradientBoostingClassifier(loss="log_loss", **common).fit(X, y)
        params, want = base + "objective=binary", h.decision_function(X)
    elif case == "categorical_effects":      # the label depends on the category: dozens of many-vs-many categorical splits (LightGBM-only rules neutralised)
        X[:, 4] = rng.integers(0, 40, n)
        y = (y + 2.0 * rng.standard_normal(40)[X[:, 4].astype(int)]).astype(np.float32)
        h = sk.HistGradientBoostingRegressor(loss="squared_error", categorical_features=[4], **common).fit(X, y.astype(np.float64))
        params = base + "objective=regression cat_l2=0 cat_smooth=10 min_data_per_group=1 max_cat_to_onehot=1 max_cat_threshold=32"
        want, ds_params = h.predict(X), DS_PARAMS + " categorical_feature=4"
    else:
        X[:, 4] = rng.integers(0, 12, n)
        w = rng.integers(1, 4, n).astype(np.float32)
        h = sk.HistGradientBoostingRegressor(loss="squared_error", categorical_features=[4], **common).fit(X, y.astype(np.float64), sample_weight=w.astype(np.float64))
        params, want, ds_params = base + "objective=regression", h.predict(X), DS_PARAMS + " categorical_feature=4"
    ds = capi.Dataset.from_mat(X, ds_params)
    ds.set_field("label", y)
    if w is not None:
        ds.set_field("weight", w)
    b = capi.Booster(ds, params)
    for _ in range(15):
        assert not b.update_one_iter()
    got = b.predict_device(X, predict_type=1).ravel()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


@pytest.mark.parametrize("objective", ["binary", "multiclass", "regression_categorical"])
def test_batched_gpu_treeshap_matches_host_predictor(built, objective):
    """C_API_PREDICT_CONTRIB in one batched kernel (explicit-stack TreeSHAP, one thread per row) against the host single-row
    predictor the reference's featuresShap UDF uses (LightGBMBooster.scala:412-423); contributions sum to the raw score
    (VerifyLightGBMRanker.scala:124 'predict == sum of SHAP')."""
    from mmlspark_b200 import capi
    rng = np.random.default_rng(93)
    n, F = 6000, 9
    X = rng.standard_normal((n, F))
    X[rng.random((n, F)) < 0.05] = np.nan
    s = np.nan_to_num(X[:, 0]) + np.nan_to_num(X[:, 1]) * np.nan_to_num(X[:, 2]) + 0.3 * rng.standard_normal(n)
    ds_params, params = DS_PARAMS, "num_leaves=31 min_data_in_leaf=5 verbosity=-1 "
    if objective == "binary":
        y, params = (s > 0).astype(np.float32), params + "objective=binary"
    elif objective == "multiclass":
        y, params = np.digitize(s, [-0.7, 0.7]).astype(np.float32), params + "objective=multiclass num_class=3"
    else:
        X[:, 5] = rng.integers(0, 9, n)
        y, params, ds_params = (s + (X[:, 5] % 3)).astype(np.float32), params + "objective=regression", DS_PARAMS + " categorical_feature=5"
    ds = capi.Dataset.from_mat(X, ds_params)
    ds.set_field("label", y)
    b = capi.Booster(ds, params)
    for _ in range(12):
        b.update_one_iter()
    K = 3 if objective == "multiclass" else 1
    Xs = X[:700]
    dev = b.predict_device(Xs, predict_type=capi.PREDICT_CONTRIB)
    assert dev.shape == (700, K * (F + 1))
    host = np.stack([b.predict_for_mat_single(r, capi.PREDICT_CONTRIB) for r in Xs[:150]])
    np.testing.assert_allclose(dev[:150], host, rtol=0, atol=1e-12)
    raw = b.predict_device(Xs, predict_type=capi.PREDICT_RAW_SCORE).reshape(700, K)
    np.testing.assert_allclose(dev.reshape(700, K, F + 1).sum(axis=2), raw, rtol=0, atol=1e-9)
    # float32 input and an iteration window
    dev32 = b.predict_device(Xs.astype(np.float32), predict_type=capi.PREDICT_CONTRIB, start_iteration=2, num_iteration=5)
    raw32 = b.predict_device(Xs.astype(np.float32), predict_type=capi.PREDICT_RAW_SCORE, start_iteration=2, num_iteration=5).reshape(700, K)
    np.testing.assert_allclose(dev32.reshape(700, K, F + 1).sum(axis=2), raw32, rtol=0, atol=1e-9)


def test_partition_column_copy_is_transparent(built, monkeypatch):
    """The booster keeps an optional [feature][row] copy of the training tiles so that k_partition reads one byte per row instead of a
    32-byte sector.  With and without it (B200GBM_COLUMN_COPY=0) the models must be identical byte for byte, on a row count that is not a
    multiple of the 256-row transposition block, with categorical + NaN features, and both must match the oracle."""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    from oracle import oracle as O
    rng = np.random.default_rng(77)
    n, F = 70_001, 37                                   # two tiles, the second one partly filled
    X = rng.standard_normal((n, F))
    X[:, 3] = rng.integers(0, 30, n)
    X[rng.random(n) < 0.05, 5] = np.nan
    y = (X[:, 0] + np.sin(2 * X[:, 1]) + (X[:, 3] % 3) + X[:, 36] * 0.5 + 0.3 * rng.standard_normal(n) > 0.8).astype(np.float32)
    dsp = DS_PARAMS + " categorical_feature=3"
    params = _classifier_params("binary", "is_unbalance=false")
    texts, copies = [], []
    for env in ("1", "0"):
        monkeypatch.setenv("B200GBM_COLUMN_COPY", env)
        ds = capi.Dataset.from_mat(X, dsp)
        ds.set_field("label", y)
        b = capi.Booster(ds, params)
        for _ in range(8):
            assert not b.update_one_iter()
        texts.append(b.save_model_to_string())
        copies.append(b.get_memory_info()["partition_column_copy_bytes"])
        b.free(); ds.free()
    assert copies[0] == 2 * 32 * 70_144 and copies[1] == 0
    assert texts[0] == texts[1]
    ods = O.OracleDataset(X, dsp)
    ods.set_field("label", y)
    ob = O.OracleBooster(ods, params)
    ob.train(8)
    compare_models(parse_model(texts[0]), parse_model(ob.model_string()))
