"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): every hand-synchronised kernel of the engine at shapes a
sanitizer finishes in minutes — K4 with TMA bulk copies (root) and cp.async gathers (leaves), the ticket-elected pick step of k_scan, the
software grid barriers of k_partition, the threshold selection + short bitonic sorts of k_scan_wide, the column-major copy (k_tiles_to_columns), k4_hist_wide, bagging, lambdarank and the device metrics."""
import sys

import numpy as np

sys.path.insert(0, ".")
from mmlspark_b200 import capi  # noqa: E402

DS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"
BASE = "num_leaves=15 learning_rate=0.1 min_data_in_leaf=20 verbosity=-1 "
rng = np.random.default_rng(0)


def run(name, X, y, params, ds_params=DS, group=None, iters=2, metric=True):
    ds = capi.Dataset.from_mat(X, ds_params).set_field("label", y)
    if group is not None:
        ds.set_field("group", group)
    b = capi.Booster(ds, BASE + params)
    for _ in range(iters):
        b.update_one_iter()
    ev = b.get_eval(0) if metric else []
    print(name, "ok", len(b.save_model_to_string()), list(np.round(ev, 6)))
    b.free(); ds.free()


n, F = 40000, 70                       # 3 feature tiles, > 2^14 rows: K4 flushes mid-pass; 20 partition chunks
X = rng.standard_normal((n, F))
X[:, 3] = np.where(rng.random(n) < 0.2, np.nan, X[:, 3])
s = X[:, 0] + np.sin(2 * X[:, 1]) + X[:, 2] * np.nan_to_num(X[:, 3]) + 0.3 * rng.standard_normal(n)
run("binary+auc", X, (s > 0).astype(np.float32), "objective=binary metric=auc,binary_logloss")
run("regression(const hessian)+bagging", X, s.astype(np.float32), "objective=regression bagging_fraction=0.5 bagging_freq=1 metric=l2")
run("multiclass", X[:15000], np.digitize(s[:15000], [-1, 0, 1]).astype(np.float32), "objective=multiclass num_class=4 metric=multi_logloss", iters=1)
sizes = rng.integers(5, 60, 400).astype(np.int32)
m = int(sizes.sum())
rel = np.clip(np.round(X[:m, 0] + 1.5), 0, 4).astype(np.float32)
run("lambdarank+ndcg", X[:m], rel, "objective=lambdarank metric=ndcg,map eval_at=1,3 min_data_in_leaf=5", group=sizes)
Xc = X[:30000].copy()
Xc[:, 5] = np.floor(3000.0 ** rng.random(30000)) - 1      # a wide categorical column (hundreds of bins)
Xc[:, 6] = rng.integers(0, 30, 30000)
yc = (s[:30000] + 0.5 * (Xc[:, 5] % 3) > 0.5).astype(np.float32)
run("wide categorical", Xc, yc, "objective=binary metric=binary_error", ds_params=DS + " categorical_feature=5,6")
print("sanitize smoke done")
