"""Generates tests/golden/*.json from the oracle (seeded).  There is no runnable LightGBM 3.2.110 in this
environment, so these vectors freeze the oracle's behaviour ("parity unpinned" against the
real binary) and pin the CUDA path and the host code against regressions.  Re-run: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O   # noqa: E402

DS_PARAMS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"


def dataset(seed, n, F):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, F))
    X[:, 1] = np.round(X[:, 1], 1)
    X[:, 2] = np.where(rng.random(n) < 0.7, 0.0, X[:, 2])
    X[:, 3] = np.where(rng.random(n) < 0.15, np.nan, X[:, 3])
    X[:, 4] = rng.integers(0, 6, n)
    s = X[:, 0] + np.nan_to_num(X[:, 3]) * 0.7 - 0.5 * X[:, 4] + X[:, 5] * X[:, 6] + 0.2 * rng.standard_normal(n)
    return X, s


def variant(X, name):
    """per-case feature matrix: cases named *widecat* get a high-cardinality categorical column 9 (hundreds of bins: the uint16 path)"""
    if "widecat" not in name:
        return X
    X2 = X.copy()
    X2[:, 9] = np.floor(np.abs(X[:, 9]) * 173.0) % 450.0
    return X2


def main():
    out = {}
    out["lcg_sample"] = {"%d_%d_%d" % (s, n, k): O.random_sample(s, n, k).tolist() for s, n, k in [(1, 20, 5), (1, 100, 80), (7, 1000, 10), (1, 50, 50)]}
    X, s = dataset(123, 2000, 10)
    ods = O.OracleDataset(X, DS_PARAMS)
    out["bins"] = {"seed": 123, "n": 2000, "F": 10,
                   "feature_info": [ods.feature_info(f) for f in range(10)],
                   "upper_bounds_hex": [[float(v).hex() for v in ods.upper_bounds(f)] for f in range(10)],
                   "bins_sum_per_feature": ods.bins().astype(np.int64).sum(axis=0).tolist(),
                   "bins_first_rows": ods.bins()[:5].tolist()}
    models = {}
    for name, params, y in [
        ("regression", "objective=regression num_leaves=7 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1", s.astype(np.float32)),
        ("binary", "objective=binary num_leaves=7 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1 is_unbalance=false", (s > 0).astype(np.float32)),
        ("multiclass", "objective=multiclass num_class=3 num_leaves=5 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1",
         np.clip(np.floor(s + 1.5), 0, 2).astype(np.float32)),
        # row-sampling modes and categorical splits (name prefix = label kind; optional dataset parameters after '|')
        ("regression_bagging", "objective=regression num_leaves=7 min_data_in_leaf=20 verbosity=-1 bagging_fraction=0.6 bagging_freq=2", s.astype(np.float32)),
        ("binary_rf", "objective=binary boosting_type=rf num_leaves=7 min_data_in_leaf=20 verbosity=-1 bagging_fraction=0.7 bagging_freq=1 feature_fraction=0.8",
         (s > 0).astype(np.float32)),
        ("binary_goss", "objective=binary boosting_type=goss learning_rate=0.5 num_leaves=7 min_data_in_leaf=20 verbosity=-1", (s > 0).astype(np.float32)),
        ("binary_dart", "objective=binary boosting_type=dart drop_rate=0.5 skip_drop=0.0 num_leaves=7 min_data_in_leaf=20 verbosity=-1", (s > 0).astype(np.float32)),
        ("regression_quantile", "objective=quantile alpha=0.7 num_leaves=7 min_data_in_leaf=20 verbosity=-1 learning_rate=0.3", s.astype(np.float32)),
        ("regression_categorical|categorical_feature=4", "objective=regression num_leaves=7 min_data_in_leaf=20 verbosity=-1 min_data_per_group=50 cat_smooth=5",
         s.astype(np.float32)),
        # round 2: one-vs-all multiclass, cross-entropy on {0,1} labels, a categorical feature with more than 256 bins
        ("multiclass_ova", "objective=multiclassova num_class=3 num_leaves=5 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1 sigmoid=1.0",
         np.clip(np.floor(s + 1.5), 0, 2).astype(np.float32)),
        ("binary_xentropy", "objective=cross_entropy num_leaves=7 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1", (s > 0).astype(np.float32)),
        ("regression_widecat|categorical_feature=9", "objective=regression num_leaves=7 min_data_in_leaf=20 verbosity=-1 min_data_per_group=20 cat_smooth=2 min_data_in_bin=1",
         s.astype(np.float32)),
    ]:
        d = O.OracleDataset(variant(X, name), DS_PARAMS + (" " + name.split("|")[1] if "|" in name else "")).set_field("label", y)
        b = O.OracleBooster(d, params)
        b.train(5)
        models[name] = {"params": params, "model": b.model_string(), "raw_pred_first8": b.predict_raw(variant(X, name)[:8]).tolist()}
    out["models"] = models
    # one hand-checkable split: 4 bins, constant hessian
    hist = np.zeros((256, 2)); hist[0] = [-30, 30]; hist[1] = [-10, 30]; hist[2] = [10, 30]; hist[3] = [40, 30]
    out["best_split_case"] = {"hist4": hist[:4].tolist(), "result": O.best_split(hist, 4, 0, 0, 0, 10.0, 120.0, 120)}
    json.dump(out, open(os.path.join(HERE, "oracle_golden.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "oracle_golden.json"))


if __name__ == "__main__":
    main()
