"""Behavioural tests of the estimator layer on the GPU, ported from the reference's suites
(lightgbm/src/test/scala/.../split1/VerifyLightGBMClassifier.scala, split2/VerifyLightGBMRegressor.scala,
split2/VerifyLightGBMRanker.scala): each parameter moves the metric the right way, model strings carry the
parameter block the reference greps, save/load round-trips, edge cases don't hang."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ")
    except Exception:
        return 0


def _auc(y, s):
    from sklearn.metrics import roc_auc_score
    return roc_auc_score(y, s)


def _binary_frame(seed=0, n=20000, F=12, imbalance=None):
    from mmlspark_b200.lightgbm import Frame
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, F))
    s = X[:, 0] + 0.8 * X[:, 1] * X[:, 2] + 0.5 * np.sin(3 * X[:, 3]) + 0.7 * rng.standard_normal(n)
    y = (s > (np.quantile(s, imbalance) if imbalance else 0)).astype(np.float64)
    return Frame({"features": X, "label": y})


def test_classifier_fit_transform_outputs(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame()
    model = LightGBMClassifier(numIterations=30, numLeaves=15, numTasks=1).fit(df)
    out = model.transform(df)
    raw, prob, pred = out["rawPrediction"], out["probability"], out["prediction"]
    assert raw.shape == prob.shape == (20000, 2)
    np.testing.assert_allclose(raw[:, 0], -raw[:, 1])                       # [-s, s]
    np.testing.assert_allclose(prob.sum(axis=1), 1.0, atol=1e-12)           # [1-p, p]
    np.testing.assert_array_equal(pred, (prob[:, 1] > 0.5).astype(np.float64))
    assert _auc(df["label"], prob[:, 1]) > 0.85
    assert model.getBoosterNumTotalIterations() == 30 and model.getBoosterNumFeatures() == 12 and model.getBoosterNumClasses() == 1
    imp = model.getFeatureImportances("split")
    assert len(imp) == 12 and np.argmax(imp) in (0, 1, 2, 3)
    # more iterations -> better training AUC (VerifyLightGBMClassifier.scala:385-397 uses numIterations at predict time)
    few = model.setNumIterations(3).transform(df)["probability"][:, 1]
    assert _auc(df["label"], few) < _auc(df["label"], prob[:, 1])


def test_native_model_save_load_roundtrip(built, tmp_path):
    from mmlspark_b200.lightgbm import LightGBMClassificationModel, LightGBMClassifier
    df = _binary_frame(1)
    model = LightGBMClassifier(numIterations=10, numTasks=1).fit(df)
    path = str(tmp_path / "model.txt")
    model.saveNativeModel(path)
    m2 = LightGBMClassificationModel.loadNativeModelFromFile(path)
    m3 = LightGBMClassificationModel.loadNativeModelFromString(model.getNativeModel())
    a, b, c = model.transform(df)["probability"], m2.transform(df)["probability"], m3.transform(df)["probability"]
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, c)
    # continue training from the saved model (modelString), like verifySaveBooster (:712-755)
    cont = LightGBMClassifier(numIterations=5, numTasks=1, modelString=model.getNativeModel()).fit(df)
    assert cont.getBoosterNumTotalIterations() == 15


def test_model_string_parameter_block_and_delegate_lr(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier, LightGBMDelegate
    df = _binary_frame(2, n=5000)
    m = LightGBMClassifier(numIterations=3, numTasks=1, lambdaL1=0.1, lambdaL2=0.5).fit(df)
    s = m.getNativeModel()
    assert "[lambda_l1: 0.1]" in s and "[lambda_l2: 0.5]" in s                # :273-275

    class D(LightGBMDelegate):
        def getLearningRate(self, batchIndex, partitionId, curIters, trainParams, previousLearningRate):
            return 0.005 if curIters >= 1 else previousLearningRate
    m = LightGBMClassifier(numIterations=3, numTasks=1, delegate=D()).fit(df)
    assert "learning_rate: 0.005" in m.getNativeModel()                        # :509
    from mmlspark_b200.modeltext import parse_model
    assert [t["shrinkage"] for t in parse_model(m.getNativeModel())["trees"]] == ["1", "0.005", "0.005"]
    named = LightGBMClassifier(numIterations=2, numTasks=1, slotNames=["f%d" % i for i in range(12)]).fit(df)
    assert "feature_names=f0 f1 f2" in named.getNativeModel()                  # :569-592


def test_params_move_the_metric_the_right_way(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(3)
    base = LightGBMClassifier(numIterations=20, numTasks=1).fit(df)
    trees = lambda m: sum(1 for l in m.getNativeModel().split("\n") if l.startswith("num_leaves=") and l != "num_leaves=1")   # noqa: E731
    leaves = lambda m: sum(int(l.split("=")[1]) for l in m.getNativeModel().split("\n") if l.startswith("num_leaves="))      # noqa: E731
    strict = LightGBMClassifier(numIterations=20, numTasks=1, minGainToSplit=50.0).fit(df)                                   # :345-350
    assert leaves(strict) < leaves(base)
    shallow = LightGBMClassifier(numIterations=20, numTasks=1, maxDepth=2).fit(df)
    assert leaves(shallow) <= 20 * 4 and trees(shallow) == 20
    big_leaf = LightGBMClassifier(numIterations=20, numTasks=1, minDataInLeaf=2000).fit(df)
    assert leaves(big_leaf) < leaves(base)
    # maxDeltaStep clips leaf outputs (:375-383)
    clipped = LightGBMClassifier(numIterations=5, numTasks=1, maxDeltaStep=0.1, learningRate=1.0).fit(df)
    from mmlspark_b200.modeltext import parse_model
    t1 = parse_model(clipped.getNativeModel())["trees"][1]
    assert np.abs(t1["leaf_value"]).max() <= 0.1 + 1e-12


def test_is_unbalance_and_weight_column(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(4, imbalance=0.95)
    y = df["label"]
    plain = LightGBMClassifier(numIterations=20, numTasks=1).fit(df).transform(df)["probability"][:, 1]
    unb = LightGBMClassifier(numIterations=20, numTasks=1, isUnbalance=True).fit(df).transform(df)["probability"][:, 1]
    assert unb[y == 1].mean() > plain[y == 1].mean() + 0.1                     # rare class gets more mass (:421-427)
    w = np.where(y == 1, 20.0, 1.0)
    wdf = df.with_column("w", w)
    wt = LightGBMClassifier(numIterations=20, numTasks=1, weightCol="w").fit(wdf).transform(df)["probability"][:, 1]
    assert wt[y == 1].mean() > plain[y == 1].mean() + 0.1                      # :399-419


def test_validation_early_stopping(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(5, n=12000)
    rng = np.random.default_rng(0)
    vdf = df.with_column("valid", rng.random(12000) < 0.3)
    for metric in ("auc", "binary_logloss", "binary_error"):                    # :429-461
        m = LightGBMClassifier(numIterations=300, numTasks=1, learningRate=0.3, numLeaves=63, validationIndicatorCol="valid",
                               earlyStoppingRound=5, metric=metric).fit(vdf)
        assert 0 < m.getBoosterNumTotalIterations() < 300, metric
        assert m.getBoosterBestIteration() >= 0 and m.getBoosterBestIteration() <= m.getBoosterNumTotalIterations()
        out = m.transform(df)                                                   # predicts with numIterations = best iteration (Appendix D)
        assert out["probability"].shape == (12000, 2)


def test_leaf_and_shap_columns(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(6, n=3000)
    m = LightGBMClassifier(numIterations=7, numLeaves=9, numTasks=1, leafPredictionCol="leaves", featuresShapCol="shap").fit(df)
    out = m.transform(df)
    assert out["leaves"].shape == (3000, 7) and (out["leaves"] == np.round(out["leaves"])).all() and out["leaves"].max() < 9
    assert out["shap"].shape == (3000, 13)
    np.testing.assert_allclose(out["shap"].sum(axis=1), out["rawPrediction"][:, 1], rtol=1e-8, atol=1e-8)
    assert len(m.getFeatureShaps(df["features"][0])) == 13


def test_multiclass_classifier(built):
    from mmlspark_b200.lightgbm import Frame, LightGBMClassifier
    rng = np.random.default_rng(7)
    n, F, K = 15000, 10, 4
    X = rng.standard_normal((n, F))
    y = np.argmax(X[:, :K] + 0.5 * rng.standard_normal((n, K)), axis=1).astype(np.float64)
    df = Frame({"features": X, "label": y})
    m = LightGBMClassifier(objective="multiclass", numIterations=20, numTasks=1).fit(df)
    out = m.transform(df)
    assert out["probability"].shape == (n, K)
    np.testing.assert_allclose(out["probability"].sum(axis=1), 1.0, atol=1e-9)
    assert (out["prediction"] == y).mean() > 0.7
    assert m.getBoosterNumClasses() == K and m.getBoosterNumTotalModel() == 20 * K


def test_regressor_and_custom_objective(built):
    from mmlspark_b200.lightgbm import Frame, LightGBMRegressor
    rng = np.random.default_rng(8)
    n, F = 20000, 10
    X = rng.standard_normal((n, F))
    y = 3 * X[:, 0] + np.sin(2 * X[:, 1]) * 2 + X[:, 2] * X[:, 3] + 0.3 * rng.standard_normal(n)
    df = Frame({"features": X, "label": y})
    rmse = lambda m: float(np.sqrt(np.mean((m.transform(df)["prediction"] - y) ** 2)))   # noqa: E731
    m10 = LightGBMRegressor(numIterations=10, numTasks=1).fit(df)
    m50 = LightGBMRegressor(numIterations=50, numTasks=1).fit(df)
    assert rmse(m50) < rmse(m10) < float(np.std(y))
    assert abs(m50.predict(X[0]) - m50.transform(df)["prediction"][0]) < 1e-12

    class L2(object):                                                          # FObjTrait.getGradient (params/FObjTrait.scala:16)
        def getGradient(self, preds, labels):
            return preds - labels, np.ones_like(preds)
    mc = LightGBMRegressor(numIterations=10, numTasks=1, fobj=L2(), boostFromAverage=False).fit(df)
    mb = LightGBMRegressor(numIterations=10, numTasks=1, boostFromAverage=False).fit(df)
    np.testing.assert_allclose(mc.transform(df)["prediction"], mb.transform(df)["prediction"], rtol=1e-6, atol=1e-6)
    bad = LightGBMRegressor(numIterations=2, numTasks=1, slotNames=["a"] * 3)   # wrong number of slot names fails early (VerifyLightGBMRegressor.scala:142-146)
    with pytest.raises(Exception):
        bad.fit(df)


def test_ranker(built):
    from mmlspark_b200.lightgbm import Frame, LightGBMRanker
    rng = np.random.default_rng(9)
    sizes = rng.integers(5, 30, 500)
    q = np.repeat(np.arange(500), sizes)
    n = len(q)
    X = rng.standard_normal((n, 8))
    rel = np.clip(np.round(X[:, 0] + 0.5 * X[:, 1] + 0.5 * rng.standard_normal(n) + 1.5), 0, 4)
    perm = rng.permutation(n)
    df = Frame({"features": X[perm], "label": rel[perm], "query": q[perm]})
    m = LightGBMRanker(groupCol="query", numIterations=20, numTasks=1, minDataInLeaf=5).fit(df)
    pred = m.transform(df)["prediction"]
    assert np.corrcoef(pred, df["label"])[0, 1] > 0.6
    assert "objective=lambdarank" in m.getNativeModel()
    shap = m.getFeatureShaps(df["features"][0])
    assert abs(sum(shap) - m.predict(df["features"][0])) < 1e-9               # VerifyLightGBMRanker.scala:124


def test_ranker_validation_rows_are_grouped_before_counting(built):
    """preprocessData (sortWithinPartitions(groupCol)) also runs on the validation frame (LightGBMBase.scala:465-468): with shuffled
    validation rows the query sizes handed to the validation dataset must still be the true group sizes (a fragmented group list
    still sums to num_data, so nothing else would notice), and validation NDCG must not collapse to the trivial single-document value."""
    from mmlspark_b200.lightgbm import Frame, LightGBMRanker
    from mmlspark_b200.lightgbm import train_utils as tu
    rng = np.random.default_rng(19)
    sizes = rng.integers(8, 25, 400)
    q = np.repeat(np.arange(400), sizes)
    n = len(q)
    X = rng.standard_normal((n, 6))
    rel = np.clip(np.round(X[:, 0] + 0.5 * X[:, 1] + 0.4 * rng.standard_normal(n) + 1.5), 0, 4)
    is_valid = np.isin(q, np.arange(0, 400, 4))            # every 4th query is validation data
    seen = []

    class Spy(LightGBMRanker):
        def _make_dataset(self, part, params_str, reference=None):
            if reference is not None:
                seen.append(tu.count_cardinality(part[self.get("groupCol")].tolist()))
            return super()._make_dataset(part, params_str, reference=reference)

    order = rng.permutation(n)
    df = Frame({"features": X[order], "label": rel[order], "query": q[order], "valid": is_valid[order]})
    m = Spy(groupCol="query", validationIndicatorCol="valid", numIterations=10, numTasks=1, minDataInLeaf=5, evalAt=[1, 3]).fit(df)
    assert len(seen) == 1
    assert seen[0] == [int(s) for s in sizes[0::4]]         # whole groups, in group-id order
    assert 1 <= m.getBoosterNumTotalIterations() <= 10      # earlyStoppingRound = 0: trainCore stops at the first non-improving round (TrainUtils.scala:139-143)


def test_batches_and_empty_partition_do_not_hang(built):
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(10, n=9000)
    m = LightGBMClassifier(numIterations=5, numTasks=1, numBatches=3).fit(df)   # :278-281
    assert m.getBoosterNumTotalIterations() == 15

    class OneEmpty(LightGBMClassifier):                                         # empty partition -> "ignore" status (:594-606)
        def _partitions(self, df, num_tasks):
            n = df.num_rows()
            return [slice(0, n), slice(n, n)]
    m = OneEmpty(numIterations=4, numTasks=2).fit(df)
    assert m.getBoosterNumTotalIterations() == 4


def test_two_tasks_two_gpus(built):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(11)
    m2 = LightGBMClassifier(numIterations=20, numTasks=2, defaultListenPort=24400).fit(df)     # numPartitions = 2 (:126)
    m1 = LightGBMClassifier(numIterations=20, numTasks=1).fit(df)
    a2 = _auc(df["label"], m2.transform(df)["probability"][:, 1])
    a1 = _auc(df["label"], m1.transform(df)["probability"][:, 1])
    assert abs(a1 - a2) < 0.01 and a2 > 0.85


BOOSTING_TYPES = ["gbdt", "rf", "dart", "goss"]          # VerifyLightGBMClassifier.scala:135


def _with_boosting(est, boosting_type):
    est.setBoostingType(boosting_type)
    if boosting_type == "rf":                             # VerifyLightGBMClassifier.scala:655-658
        est.setBaggingFraction(0.9)
        est.setBaggingFreq(1)
    return est


def test_classifier_all_boosting_types(built):
    """'can be trained and scored on <file>' sweeps boostingTypes (VerifyLightGBMClassifier.scala:644-668): probabilities are
    well-formed, importances have one entry per feature, the model is better than chance."""
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(3)
    aucs = {}
    for bt in BOOSTING_TYPES:
        model = _with_boosting(LightGBMClassifier(numIterations=25, numLeaves=15, numTasks=1), bt).fit(df)
        out = model.transform(df)
        np.testing.assert_allclose(out["probability"].sum(axis=1), 1.0, atol=1e-12)
        assert len(model.getFeatureImportances("split")) == 12 and len(model.getFeatureImportances("gain")) == 12
        aucs[bt] = _auc(df["label"], out["probability"][:, 1])
        assert ("boosting: %s]" % bt) in model.getNativeModel()
    assert all(a > 0.8 for a in aucs.values()), aucs
    assert "average_output" in _with_boosting(LightGBMClassifier(numIterations=3, numTasks=1), "rf").fit(df).getNativeModel()


def test_multiclassova_classifier_and_high_cardinality_categorical_slots(built):
    """objective=multiclassova (LightGBMParams.scala:296-300): one sigmoid per class from the native output transform (probabilities do
    NOT sum to 1); categoricalSlotIndexes with a high-cardinality column (LightGBMBase.scala:168-199) takes the > 256-bin path, and a saved
    model reloaded through loadNativeModelFromString predicts the same."""
    from mmlspark_b200.lightgbm import Frame, LightGBMClassifier, LightGBMClassificationModel
    rng = np.random.default_rng(17)
    n, K = 40000, 3
    cat = np.floor(2000.0 ** rng.random(n)) - 1                      # ~1000 distinct categories, log-uniform
    eff = rng.standard_normal(2000)
    X = np.column_stack([rng.standard_normal((n, 5)), cat])
    s = X[:, 0] + 0.8 * eff[cat.astype(int)] + 0.3 * rng.standard_normal(n)
    y = np.digitize(s, [-0.7, 0.7]).astype(np.float64)
    df = Frame({"features": X, "label": y})
    m = LightGBMClassifier(objective="multiclassova", numIterations=25, numLeaves=15, numTasks=1, categoricalSlotIndexes=[5]).fit(df)
    out = m.transform(df)
    p = out["probability"]
    assert p.shape == (n, K) and (p > 0).all() and (p < 1).all()
    assert np.abs(p.sum(axis=1) - 1.0).max() > 1e-3                  # independent sigmoids, not a softmax
    np.testing.assert_allclose(p, 1.0 / (1.0 + np.exp(-out["rawPrediction"])), rtol=1e-9, atol=1e-12)
    assert (out["prediction"] == y).mean() > 0.75
    text = m.getNativeModel()
    assert "objective=multiclassova num_class:3 sigmoid:1" in text
    assert any(int(l.split("=")[1]) > 0 for l in text.split("\n") if l.startswith("num_cat=")), "expected splits on the categorical slot"
    m2 = LightGBMClassificationModel.loadNativeModelFromString(text)
    np.testing.assert_allclose(m2.transform(df)["probability"], p, rtol=1e-12)


def test_multiclass_and_regressor_all_boosting_types(built):
    """VerifyLightGBMClassifier.scala:676-705 (multiclass sweep) and VerifyLightGBMRegressor.scala:188-207 (regression sweep)."""
    from mmlspark_b200.lightgbm import Frame, LightGBMClassifier, LightGBMRegressor
    rng = np.random.default_rng(12)
    n, F = 15000, 8
    X = rng.standard_normal((n, F))
    s = 2 * X[:, 0] + X[:, 1] * X[:, 2] + 0.4 * rng.standard_normal(n)
    ycls = np.digitize(s, [-1.0, 1.0]).astype(np.float64)
    dfc, dfr = Frame({"features": X, "label": ycls}), Frame({"features": X, "label": s})
    for bt in BOOSTING_TYPES:
        mc = _with_boosting(LightGBMClassifier(objective="multiclass", numIterations=15, numLeaves=15, numTasks=1), bt).fit(dfc)
        out = mc.transform(dfc)
        assert out["probability"].shape == (n, 3)
        np.testing.assert_allclose(out["probability"].sum(axis=1), 1.0, atol=1e-12)
        assert float(np.mean(out["prediction"] == ycls)) > 0.7, bt
        mr = _with_boosting(LightGBMRegressor(numIterations=30, numLeaves=15, numTasks=1), bt).fit(dfr)
        rmse = float(np.sqrt(np.mean((mr.transform(dfr)["prediction"] - s) ** 2)))
        assert rmse < 0.75 * float(np.std(s)), (bt, rmse)
        assert len(mr.getFeatureImportances("split")) == F


def test_dart_mode_parameters(built):
    """'Verify LightGBM Classifier with dart mode parameters' (VerifyLightGBMClassifier.scala:352-368): the dart knobs are accepted,
    reach the engine (parameter block) and change the model."""
    from mmlspark_b200.lightgbm import LightGBMClassifier
    df = _binary_frame(4)
    m1 = LightGBMClassifier(numIterations=40, numLeaves=15, numTasks=1).setBoostingType("dart").setSkipDrop(1.0).fit(df)
    m2 = (LightGBMClassifier(numIterations=40, numLeaves=15, numTasks=1).setBoostingType("dart").setXGBoostDartMode(True).setDropRate(0.6)
          .setMaxDrop(60).setSkipDrop(0.4).setUniformDrop(True).fit(df))
    s1, s2 = m1.getNativeModel(), m2.getNativeModel()
    assert "[skip_drop: 1]" in s1 and "[xgboost_dart_mode: 1]" in s2 and "[drop_rate: 0.6]" in s2 and "[max_drop: 60]" in s2 and "[uniform_drop: 1]" in s2
    p1, p2 = m1.transform(df)["probability"][:, 1], m2.transform(df)["probability"][:, 1]
    assert _auc(df["label"], p1) > 0.85 and _auc(df["label"], p2) > 0.85
    assert np.abs(p1 - p2).max() > 1e-3
    # skip_drop = 1 never drops => identical to plain gbdt
    m0 = LightGBMClassifier(numIterations=40, numLeaves=15, numTasks=1).fit(df)
    np.testing.assert_allclose(m0.transform(df)["probability"], m1.transform(df)["probability"], rtol=0, atol=1e-12)


def test_continued_training_with_initial_score(built):
    """'continued training with initial score' (VerifyLightGBMClassifier.scala): the raw score of a first model is fed back as
    initScoreCol; the second fit starts from it (no boost-from-average) and improves the training AUC."""
    from mmlspark_b200.lightgbm import Frame, LightGBMClassifier
    df = _binary_frame(5)
    m1 = LightGBMClassifier(numIterations=10, numLeaves=7, numTasks=1).fit(df)
    out1 = m1.transform(df)
    df2 = Frame({"features": df["features"], "label": df["label"], "init": out1["rawPrediction"][:, 1]})
    m2 = LightGBMClassifier(numIterations=10, numLeaves=7, numTasks=1, initScoreCol="init").fit(df2)
    raw2 = m2.transform(df2)["rawPrediction"][:, 1] + df2["init"]          # the model holds only the increment
    assert _auc(df["label"], raw2) > _auc(df["label"], out1["rawPrediction"][:, 1])
    # same as training 20 iterations in one go up to the init-score handling of the first tree
    m20 = LightGBMClassifier(numIterations=20, numLeaves=7, numTasks=1).fit(df)
    assert abs(_auc(df["label"], raw2) - _auc(df["label"], m20.transform(df)["rawPrediction"][:, 1])) < 0.01


def test_max_delta_step_and_tweedie(built):
    """'max delta step parameter' (classifier) and 'tweedie distribution' (regressor, VerifyLightGBMRegressor.scala:147-154)."""
    from mmlspark_b200.lightgbm import Frame, LightGBMClassifier, LightGBMRegressor
    df = _binary_frame(6)
    base = dict(numIterations=30, numLeaves=15, numTasks=1, learningRate=0.9)
    p1 = LightGBMClassifier(**base).fit(df).transform(df)["probability"][:, 1]
    m2 = LightGBMClassifier(maxDeltaStep=0.5, **base).fit(df)
    p2 = m2.transform(df)["probability"][:, 1]
    assert "[max_delta_step: 0.5]" in m2.getNativeModel()
    assert np.abs(p1 - p2).max() > 1e-3
    from mmlspark_b200.modeltext import parse_model
    for t in parse_model(m2.getNativeModel())["trees"][1:]:          # |leaf output| <= learning_rate * max_delta_step (first tree also carries the bias)
        assert np.abs(t["leaf_value"]).max() <= 0.9 * 0.5 + 1e-12
    rng = np.random.default_rng(3)
    X = rng.random((8000, 6))
    y = rng.poisson(np.exp(1.5 * X[:, 0] - X[:, 1])) * rng.gamma(2.0, 1.0, 8000)
    dfr = Frame({"features": X, "label": y})
    mt = LightGBMRegressor(objective="tweedie", tweedieVariancePower=1.5, numIterations=30, numTasks=1).fit(dfr)
    pred = mt.transform(dfr)["prediction"]
    assert "objective=tweedie" in mt.getNativeModel() and "[tweedie_variance_power: 1.5]" in mt.getNativeModel()
    assert (pred > 0).all() and np.corrcoef(pred, np.exp(1.5 * X[:, 0] - X[:, 1]))[0, 1] > 0.8


def test_slot_names_and_categorical_slots(built):
    """'slot names parameter' (renamed slot shows up in the model string) and 'categorical parameter for dense dataset' /
    'Regressor categorical parameter': categoricalSlotNames / categoricalSlotIndexes reach the engine as categorical_feature."""
    from mmlspark_b200.lightgbm import Frame, LightGBMClassifier, LightGBMRegressor
    from mmlspark_b200.modeltext import parse_model
    rng = np.random.default_rng(9)
    n = 12000
    X = rng.standard_normal((n, 5))
    X[:, 3] = rng.integers(0, 8, n)
    effect = np.array([2.0, -1.0, 0.5, 3.0, -2.5, 0.0, 1.0, -0.5])[X[:, 3].astype(int)]       # non-monotone in the category id
    y = (X[:, 0] + effect + 0.3 * rng.standard_normal(n) > 0.5).astype(np.float64)
    df = Frame({"features": X, "label": y})
    names = ["f0", "f1", "f2", "Age_years", "f4"]
    mc = LightGBMClassifier(numIterations=20, numLeaves=15, numTasks=1, slotNames=names, categoricalSlotNames=["Age_years"]).fit(df)
    s = mc.getNativeModel()
    assert "Age_years" in s and "feature_names=f0 f1 f2 Age_years f4" in s
    trees = parse_model(s)["trees"]
    assert any(t.get("num_cat", 0) > 0 for t in trees)                       # bitset splits on the categorical slot
    mnum = LightGBMClassifier(numIterations=20, numLeaves=15, numTasks=1, slotNames=names).fit(df)
    assert all(t.get("num_cat", 0) == 0 for t in parse_model(mnum.getNativeModel())["trees"])
    assert _auc(y, mc.transform(df)["probability"][:, 1]) >= _auc(y, mnum.transform(df)["probability"][:, 1]) - 1e-3
    mr = LightGBMRegressor(numIterations=10, numTasks=1, categoricalSlotIndexes=[3]).fit(Frame({"features": X, "label": effect + X[:, 0]}))
    assert any(t.get("num_cat", 0) > 0 for t in parse_model(mr.getNativeModel())["trees"])
    with pytest.raises(ValueError, match="Invalid slot names"):
        LightGBMRegressor(numIterations=2, numTasks=1, slotNames=["a", "b", "c[", "d", "e"]).fit(df)


def test_ranker_query_column_types_and_shap(built):
    """'Ranker with int, long and string query column', 'Throws error when group column is not long, int or string',
    'Ranker feature shaps' (predict == sum of SHAP, VerifyLightGBMRanker.scala:110-125)."""
    from mmlspark_b200.lightgbm import Frame, LightGBMRanker
    rng = np.random.default_rng(21)
    nq, per = 300, 12
    q = np.repeat(np.arange(nq), per)
    X = rng.standard_normal((nq * per, 4))
    rel = np.clip(np.round(X[:, 0] + 0.5 * rng.standard_normal(nq * per) + 1.5), 0, 4)
    kw = dict(numIterations=8, numLeaves=7, numTasks=1, groupCol="query", evalAt=(1, 2, 3), minDataInLeaf=5)
    preds = []
    for qcol in (q.astype(np.int64), q.astype(np.int32), np.array(["str_%04d" % v for v in q])):
        m = LightGBMRanker(featuresShapCol="shap", **kw).fit(Frame({"features": X, "label": rel, "query": qcol}))
        out = m.transform(Frame({"features": X, "label": rel, "query": qcol}))
        preds.append(out["prediction"])
        np.testing.assert_allclose(out["shap"].sum(axis=1), out["prediction"], rtol=0, atol=1e-9)
    # same groups whatever the key type (zero-padded strings sort like the integers); lambdarank sums fp32 lambdas with shared-memory
    # float atomics, so repeated fits agree to ~1e-9, not bit for bit
    np.testing.assert_allclose(preds[0], preds[1], rtol=0, atol=1e-6)
    np.testing.assert_allclose(preds[0], preds[2], rtol=0, atol=1e-6)
    with pytest.raises(ValueError, match="int, long or string"):
        LightGBMRanker(**kw).fit(Frame({"features": X, "label": rel, "query": q.astype(np.float64) + 0.5}))
