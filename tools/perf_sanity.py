"""Throughput sanity of the non-headline estimator shapes (BASELINE.json configs 2, 4, 5 scaled to fit a short run):
regression L2, lambdarank with ~100-row query groups, 10-class softmax with categorical columns.  Prints iterations/s."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mmlspark_b200 import capi  # noqa: E402

DS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"
BASE = "num_leaves=31 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1 "


def run(name, X, y, params, ds_params=DS, group=None, iters=12, warm=3):
    t0 = time.perf_counter()
    ds = capi.Dataset.from_mat(X, ds_params)
    ds.set_field("label", y)
    if group is not None:
        ds.set_field("group", group)
    t_ds = time.perf_counter() - t0
    b = capi.Booster(ds, BASE + params)
    for _ in range(warm):
        b.update_one_iter()
    b.get_timing(reset=True)
    t0 = time.perf_counter()
    for _ in range(iters):
        b.update_one_iter()
    dt = time.perf_counter() - t0
    tm = b.get_timing()
    print(json.dumps({"case": name, "rows": int(X.shape[0]), "features": int(X.shape[1]), "iters_per_s": round(iters / dt, 3), "ms_per_iter": round(1000 * dt / iters, 2),
                      "dataset_create_s": round(t_ds, 2), "launches_per_iter": tm["launches"] / iters}))
    b.free(); ds.free()


rng = np.random.default_rng(1)
n, F = 5_000_000, 256
X = rng.random((n, F), dtype=np.float32)
y = (np.sin(6.28 * X[:, :16]).sum(axis=1) + 0.1 * rng.standard_normal(n)).astype(np.float32)
run("cfg2-like regression L2 (constant hessian, NATOM=3)", X, y, "objective=regression")
del X
n, F = 2_000_000, 136
X = rng.standard_normal((n, F), dtype=np.float32)
sizes = rng.integers(50, 151, size=n // 100 + 10)
sizes = sizes[np.cumsum(sizes) <= n]
sizes = np.append(sizes, n - sizes.sum()).astype(np.int32)
rel = np.clip(np.round(X[:, 0] + 0.5 * X[:, 1] + rng.standard_normal(n) + 2), 0, 4).astype(np.float32)
run("cfg4-like lambdarank (%d groups)" % len(sizes), X, rel, "objective=lambdarank", group=sizes)
n, F = 2_000_000, 128
X = rng.standard_normal((n, F), dtype=np.float32)
X[:, 112:] = np.floor(np.abs(rng.standard_normal((n, 16))) ** 3 * 20) % 150           # 16 skewed categorical columns (cardinality 150: bins must fit uint8)
W = rng.standard_normal((112, 10))
ycls = np.argmax(X[:, :112] @ W + 2 * rng.standard_normal((n, 10)), axis=1).astype(np.float32)
run("cfg5-like multiclass(10) with 16 categorical columns", X, ycls, "objective=multiclass num_class=10",
    ds_params=DS + " categorical_feature=" + ",".join(str(c) for c in range(112, 128)))
