"""GPU parity tests around the path the benchmark itself runs:
  * LGBM_DatasetCreateFromSampledColumn + LGBM_DatasetPushRows (replaces DatasetAggregator.scala:87-95,173-175 chunk feeding) against
    LGBM_DatasetCreateFromMat and the oracle: f32/f64, >= 3 chunks with a ragged tail, numpy / pinned-host / device sources;
  * LGBM_DatasetCreateFromCSR (DatasetAggregator.scala:438-459) without densifying, incl. training on it;
  * mid-scale training parity (5M x 128, 5 iterations) where K4 takes its multi-flush (2^14 rows) and multi-chunk-per-tile paths;
  * size-independent properties at a size whose tile offsets cross 2^32 bytes (histogram conservation, row-sample binning)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DS_PARAMS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"
SEED = 2025


def _params(objective, extra="", leaves=31):
    return ("metric= boost_from_average=true is_pre_partition=True boosting_type=gbdt tree_learner=data_parallel top_k=20 num_iterations=100 "
            "learning_rate=0.1 num_leaves=%d max_bin=255 bagging_fraction=1.0 bagging_freq=0 feature_fraction=1.0 max_depth=-1 "
            "min_sum_hessian_in_leaf=0.001 num_machines=1 verbosity=-1 lambda_l1=0.0 lambda_l2=0.0 min_gain_to_split=0.0 max_delta_step=0.0 "
            "min_data_in_leaf=20 objective=%s num_threads=0 %s" % (leaves, objective, extra))


def host_bin(X, info, ub):
    """numpy restatement of BinMapper::ValueToBin for numerical features (independent of both the product and the oracle code)."""
    nb = info["num_bin"] - (1 if info["missing_type"] == 2 else 0)
    v = np.asarray(X, dtype=np.float64).copy()
    nan = np.isnan(v)
    v[nan] = 0.0
    b = np.searchsorted(ub[:nb - 1], v, side="left")
    if info["missing_type"] == 2:
        b[nan] = info["num_bin"] - 1
    return b


def _mixed_matrix(rng, n, F):
    X = rng.standard_normal((n, F))
    X[:, 1] = np.where(rng.random(n) < 0.2, np.nan, X[:, 1])
    X[:, 2] = np.where(rng.random(n) < 0.7, 0.0, X[:, 2])
    X[:, 3] = rng.integers(0, 6, n)
    X[:, 4] = 2.5                                  # trivial column
    X[:, 5] = np.round(X[:, 5], 1)
    return X


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("source", ["numpy", "pinned", "device"])
def test_push_rows_matches_from_mat_and_oracle(built, dtype, source):
    from mmlspark_b200 import capi
    from oracle import oracle as O
    rng = np.random.default_rng(17)
    n, F = 260_003, 37                             # > bin_construct_sample_cnt, so the LCG row sample matters; 37 features = 2 tiles, ragged
    X = _mixed_matrix(rng, n, F).astype(dtype)
    Xd = X.astype(np.float64)
    ref = capi.Dataset.from_mat(X, DS_PARAMS)
    ods = O.OracleDataset(Xd, DS_PARAMS)
    rows = capi.sample_indices(n, 200000, 1)
    ds = capi.Dataset.from_sampled_columns(Xd[rows], n, DS_PARAMS)
    chunk = 70_000                                 # 4 chunks, last one ragged (50 003 rows)
    code = capi.DTYPE_FLOAT32 if dtype == np.float32 else capi.DTYPE_FLOAT64
    esz = np.dtype(dtype).itemsize
    pinned = capi.PinnedBuffer(chunk * F * esz) if source in ("pinned", "device") else None
    dev = capi.DeviceBuffer(chunk * F * esz) if source == "device" else None
    for off in range(0, n, chunk):
        blk = np.ascontiguousarray(X[off:off + chunk])
        if source == "numpy":
            ds.push_rows(blk, off)
            continue
        pinned.as_array(dtype, (chunk, F))[:len(blk)] = blk
        ptr = pinned.ptr
        if source == "device":
            capi.memcpy(dev.ptr, pinned.ptr, len(blk) * F * esz)
            ptr = dev.ptr
        ds.push_rows(ptr, off, nrow=len(blk), ncol=F, dtype_code=code)
    for f in range(F):
        assert ds.feature_info(f) == ref.feature_info(f) == ods.feature_info(f), "feature %d meta differs" % f
        assert ds.upper_bounds(f).tobytes() == ref.upper_bounds(f).tobytes() == ods.upper_bounds(f).tobytes(), "feature %d bounds differ" % f
        assert ds.feature_range(f) == ref.feature_range(f)
    want = ods.bins()
    assert np.array_equal(ds.get_bins(), want), "pushed bins differ from the oracle"
    assert np.array_equal(ref.get_bins(), want)
    pick = np.sort(rng.choice(n, 5000, replace=False)).astype(np.int32)
    assert np.array_equal(ds.get_bins_rows(pick), want[pick].astype(np.uint16))
    for b in (pinned, dev):
        if b is not None:
            b.free()
    ds.free(); ref.free()


def _random_csr(rng, n, F, density, nan_rate=0.02):
    mask = rng.random((n, F)) < density
    D = np.where(mask, rng.standard_normal((n, F)), 0.0)
    D[:, 3] = np.where(mask[:, 3], rng.integers(1, 9, n), 0.0)
    if nan_rate > 0:
        D[mask & (rng.random((n, F)) < nan_rate)] = np.nan      # explicit NaN entries
    D[0, :] = 0.0                                            # an empty row
    stored = mask.copy()
    stored[0, :] = False
    indptr = np.concatenate([[0], np.cumsum(stored.sum(axis=1))]).astype(np.int32)
    indices = np.nonzero(stored)[1].astype(np.int32)
    data = D[stored]
    return D, indptr, indices, data


def test_from_csr_bins_and_training(built):
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    from oracle import oracle as O
    rng = np.random.default_rng(23)
    n, F = 60_000, 45
    D, indptr, indices, data = _random_csr(rng, n, F, 0.15)
    y = (np.nan_to_num(D[:, 0]) * 2 - np.nan_to_num(D[:, 1]) + np.nan_to_num(D[:, 2]) * np.nan_to_num(D[:, 5]) + 0.1 * rng.standard_normal(n) > 0).astype(np.float32)
    ds = capi.Dataset.from_csr(indptr, indices, data, F, DS_PARAMS).set_field("label", y)
    dd = capi.Dataset.from_mat(D, DS_PARAMS).set_field("label", y)
    ods = O.OracleDataset(D, DS_PARAMS).set_field("label", y)
    for f in range(F):
        assert ds.feature_info(f) == dd.feature_info(f) == ods.feature_info(f)
        assert ds.upper_bounds(f).tobytes() == ods.upper_bounds(f).tobytes()
    assert np.array_equal(ds.get_bins(), ods.bins())
    # a validation set created from CSR with reference= reuses the training mappers
    Dv, ipv, ixv, dv = _random_csr(rng, 5000, F, 0.15)
    vs = capi.Dataset.from_csr(ipv, ixv, dv, F, DS_PARAMS, reference=ds)
    vd = capi.Dataset.from_mat(Dv, DS_PARAMS, reference=dd)
    assert np.array_equal(vs.get_bins(), vd.get_bins())
    # training on a CSR dataset.  No NaN entries here: a NaN-missing feature in a leaf without NaN rows has two mathematically equal
    # scan directions whose winner is decided by fp64 rounding in any implementation (DESIGN.md, K5 "Ties"), which is not what this test is about
    D2, ip2, ix2, d2 = _random_csr(rng, n, F, 0.15, nan_rate=0.0)
    y2 = (D2[:, 0] * 2 - D2[:, 1] + D2[:, 2] * D2[:, 5] + 0.1 * rng.standard_normal(n) > 0).astype(np.float32)
    ds2 = capi.Dataset.from_csr(ip2, ix2, d2, F, DS_PARAMS).set_field("label", y2)
    ods2 = O.OracleDataset(D2, DS_PARAMS).set_field("label", y2)
    assert np.array_equal(ds2.get_bins(), ods2.bins())
    params = _params("binary", "is_unbalance=false")
    b = capi.Booster(ds2, params)
    ob = O.OracleBooster(ods2, params)
    for _ in range(8):
        assert b.update_one_iter() == ob.update()
    compare_models(parse_model(b.save_model_to_string()), parse_model(ob.model_string()))
    # malformed input fails loudly instead of writing out of bounds
    bad = indices.copy()
    bad[5] = -1
    with pytest.raises(capi.LightGBMError):
        capi.Dataset.from_csr(indptr, bad, data, F, DS_PARAMS)
    bad[5] = F
    with pytest.raises(capi.LightGBMError):
        capi.Dataset.from_csr(indptr, bad, data, F, DS_PARAMS)


def test_from_csr_wide_sparse_does_not_densify(built):
    """2^18 hashed columns: nrow x num_col doubles would be 105 GB; the CSR path needs O(nnz)."""
    from mmlspark_b200 import capi
    rng = np.random.default_rng(29)
    n, F, per_row = 50_000, 1 << 18, 12
    cols = np.sort(rng.integers(0, 4000, (n, per_row)), axis=1)          # only 4000 columns ever occur
    cols += np.arange(per_row)[None, :] * 4000                           # strictly increasing inside a row, < 48000
    indices = cols.reshape(-1).astype(np.int32)
    indptr = (np.arange(n + 1) * per_row).astype(np.int32)
    data = rng.standard_normal(n * per_row)
    ds = capi.Dataset.from_csr(indptr, indices, data, F, DS_PARAMS + " min_data_in_leaf=5")
    assert ds.num_data() == n and ds.num_feature() == F
    r = np.array([0, 1, n - 1], dtype=np.int32)
    got = ds.get_bins_rows(r)
    for i, row in enumerate(r):
        for k in range(per_row):
            f = int(cols[row, k])
            info = ds.feature_info(f)
            if info["is_trivial"]:
                continue
            want = host_bin(np.array([data[row * per_row + k]]), info, ds.upper_bounds(f))[0]
            assert got[i, f] == want
    ds.free()


def _synth_dataset(capi, n, F, kind, chunk_rows):
    sample_rows = capi.sample_indices(n, 200000, 1)
    sample, _ = capi.synthetic_rows(sample_rows, F, SEED, kind)
    ds = capi.Dataset.from_sampled_columns(sample, n, DS_PARAMS)
    dev_x = capi.DeviceBuffer(chunk_rows * F * 4)
    dev_y = capi.DeviceBuffer(chunk_rows * 4)
    label = np.empty(n, dtype=np.float32)
    for off in range(0, n, chunk_rows):
        rows = min(chunk_rows, n - off)
        capi.synthetic_fill(dev_x.ptr, dev_y.ptr, off, rows, F, SEED, kind)
        capi.memcpy(label[off:off + rows].ctypes.data, dev_y.ptr, rows * 4)
        ds.push_rows(dev_x.ptr, off, nrow=rows, ncol=F, dtype_code=capi.DTYPE_FLOAT32)
    dev_x.free(); dev_y.free()
    ds.set_field("label", label)
    return ds, label


def _check_row_sample(capi, ds, n, F, kind, rng, count=4000, extra=()):
    rows = np.unique(np.concatenate([rng.integers(0, n, count), np.array([0, 1, n - 2, n - 1] + list(extra))])).astype(np.int32)
    Xs, _ = capi.synthetic_rows(rows, F, SEED, kind)
    got = ds.get_bins_rows(rows)
    for f in range(F):
        info = ds.feature_info(f)
        if info["is_trivial"]:
            continue
        want = host_bin(Xs[:, f].astype(np.float32).astype(np.float64), info, ds.upper_bounds(f))
        assert np.array_equal(got[:, f], want), "feature %d: device bins of sampled rows differ from host binning" % f


@pytest.mark.parametrize("kind,objective", [(1, "binary"), (0, "regression")])
def test_mid_scale_training_parity_5m_x_128(built, kind, objective):
    """5M x 128: every K4 root pass needs > 1 row chunk per feature tile and flushes its sub-histogram every 2^14 rows; deep leaves take the
    gathered path.  The oracle trains on the product's downloaded bins (checked row-sample-wise against independent host binning)."""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    from oracle import oracle as O
    n, F = 5_000_000, 128
    ds, label = _synth_dataset(capi, n, F, kind, chunk_rows=1_900_000)      # 3 chunks, ragged tail
    rng = np.random.default_rng(3)
    _check_row_sample(capi, ds, n, F, kind, rng, extra=(1_899_999, 1_900_000, 3_799_999, 3_800_000))
    params = _params(objective, "is_unbalance=false" if objective == "binary" else "")
    b = capi.Booster(ds, params)
    assert b.get_info()["constant_hessian"] == (objective == "regression")
    for _ in range(5):
        assert not b.update_one_iter()
    bins = ds.get_bins()
    infos = [ds.feature_info(f) for f in range(F)]
    ods = O.OracleDataset.from_bins(bins, infos, [ds.upper_bounds(f) for f in range(F)], [ds.feature_range(f) for f in range(F)], DS_PARAMS)
    ods.set_field("label", label)
    del bins
    ob = O.OracleBooster(ods, params)
    ob.train(5)
    compare_models(parse_model(b.save_model_to_string()), parse_model(ob.model_string()))
    np.testing.assert_allclose(b.get_scores(), ob.scores(), rtol=1e-9, atol=1e-9)
    b.free(); ds.free()


def test_offsets_beyond_4gib_conservation(built):
    """20M x 256 = 5.1 GB of bins: tile base offsets and PushRows row offsets exceed 2^32 bytes.  Size-independent properties:
    (1) sampled rows bin exactly as the host binning says, incl. rows around every chunk boundary; (2) histogram conservation — with
    integer-valued gradients every feature's bins sum to EXACTLY (sum g, n) for the full pass and for an index-list leaf."""
    from mmlspark_b200 import capi
    n, F = 20_000_000, 256
    chunk = 1 << 20
    ds, _ = _synth_dataset(capi, n, F, 1, chunk_rows=chunk)
    rng = np.random.default_rng(4)
    _check_row_sample(capi, ds, n, F, 1, rng, extra=(chunk - 1, chunk, 17 * chunk - 1, 17 * chunk))
    g = ((np.arange(n) % 7) - 3).astype(np.float32)
    h = np.ones(n, dtype=np.float32)
    H = ds.histogram(g, h)
    used = [f for f in range(F) if not ds.feature_info(f)["is_trivial"]]
    assert len(used) == F
    sg = float(g.astype(np.float64).sum())
    assert np.array_equal(H[:, :, 0].sum(axis=1), np.full(F, sg)), "gradient mass is not conserved"
    assert np.array_equal(H[:, :, 1].sum(axis=1), np.full(F, float(n))), "row count is not conserved"
    idx = np.sort(rng.choice(n, 3_000_000, replace=False)).astype(np.int32)
    Hl = ds.histogram(g, h, idx)
    assert np.array_equal(Hl[:, :, 0].sum(axis=1), np.full(F, float(g[idx].astype(np.float64).sum())))
    assert np.array_equal(Hl[:, :, 1].sum(axis=1), np.full(F, float(len(idx))))
    # and the per-bin counts of one feature against a bincount of host-binned synthetic rows of a small slice
    rows = np.arange(5_000_000, 5_200_000, dtype=np.int32)
    Xs, _ = capi.synthetic_rows(rows, F, SEED, 1)
    f = 200
    want = np.bincount(host_bin(Xs[:, f].astype(np.float32).astype(np.float64), ds.feature_info(f), ds.upper_bounds(f)), minlength=256)
    Hs = ds.histogram(g, h, rows)
    assert np.array_equal(Hs[f, :, 1], want.astype(np.float64))
    ds.free()


@pytest.mark.parametrize("objective,df", [("regression", 1.5), ("regression", 3.0), ("huber", 1.5)])
def test_heavy_tailed_gradients_at_scale(built, objective, df):
    """2M x 32 with Student-t labels (df = 1.5: infinite variance, max |g| ~ 10^4 x the typical gradient).  K4 accumulates 36-bit
    fixed point scaled by max |g| (a deliberate deviation from the reference's fp64 accumulation), so a heavy tail is the worst
    case for its resolution.  Bar: identical tree structure, leaf values / gains within the north-star 1e-5, training scores within 1e-5
    relative of the label scale."""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    from oracle import oracle as O
    rng = np.random.default_rng(int(df * 10))
    n, F = 2_000_000, 32
    X = rng.normal(size=(n, F)).astype(np.float32)
    y = (X[:, 0] * 2 + np.sin(X[:, 1] * 3) + X[:, 2] * X[:, 3] + 0.5 * rng.standard_t(df, size=n)).astype(np.float32)
    ds = capi.Dataset.from_mat(X, DS_PARAMS)
    ds.set_field("label", y)
    params = _params(objective)
    b = capi.Booster(ds, params)
    for _ in range(5):
        assert not b.update_one_iter()
    ods = O.OracleDataset(X, DS_PARAMS)
    ods.set_field("label", y)
    ob = O.OracleBooster(ods, params)
    ob.train(5)
    compare_models(parse_model(b.save_model_to_string()), parse_model(ob.model_string()))
    np.testing.assert_allclose(b.get_scores(), ob.scores(), rtol=1e-5, atol=1e-5)
    b.free(); ds.free()


@pytest.mark.parametrize("objective", ["multiclassova", "cross_entropy"])
def test_multiclassova_and_cross_entropy_match_oracle(built, objective):
    """objectives the reference advertises (LightGBMParams.scala:296-300, fobj/metric docs :425-437)"""
    from mmlspark_b200 import capi
    from mmlspark_b200.modeltext import parse_model, compare_models
    from oracle import oracle as O
    rng = np.random.default_rng(31)
    n, F = 40_000, 20
    X = rng.standard_normal((n, F))
    w = (0.5 + rng.random(n)).astype(np.float32)
    if objective == "multiclassova":
        y = np.argmax(X[:, :4] + 0.5 * rng.standard_normal((n, 4)), axis=1).astype(np.float32)
        extra = "num_class=4 is_unbalance=true sigmoid=1.5"
    else:
        y = (1.0 / (1.0 + np.exp(-(X[:, 0] - X[:, 1] * X[:, 2])))).astype(np.float32)        # probabilistic labels
        extra = ""
    params = _params(objective, extra)
    ds = capi.Dataset.from_mat(X, DS_PARAMS).set_field("label", y).set_field("weight", w)
    ods = O.OracleDataset(X, DS_PARAMS).set_field("label", y).set_field("weight", w)
    b = capi.Booster(ds, params)
    ob = O.OracleBooster(ods, params)
    for _ in range(6):
        assert b.update_one_iter() == ob.update()
    ma, mb = parse_model(b.save_model_to_string()), parse_model(ob.model_string())
    compare_models(ma, mb)
    assert ma["header"]["objective"].startswith(objective)
    np.testing.assert_allclose(b.get_scores(), ob.scores(), rtol=1e-9, atol=1e-9)
    # normal predictions apply the objective's output transform (per-class sigmoid / sigmoid), on the GPU batch path and the host single-row path
    raw = ob.predict_raw(X[:200])
    sig = 1.5 if objective == "multiclassova" else 1.0
    want = 1.0 / (1.0 + np.exp(-sig * raw))
    np.testing.assert_allclose(b.predict_device(X[:200]), want, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(b.predict_for_mat_single(X[7]), want[7], rtol=1e-6, atol=1e-9)
    loaded = capi.Booster(model_str=b.save_model_to_string())
    np.testing.assert_allclose(loaded.predict_for_mat_single(X[7]), want[7], rtol=1e-6, atol=1e-9)
    ev = b.get_eval(0)
    assert len(ev) == 1 and np.isfinite(ev[0])


def test_unknown_objective_in_model_string_fails_at_load(built):
    from mmlspark_b200 import capi
    rng = np.random.default_rng(2)
    X = rng.standard_normal((2000, 5))
    y = X[:, 0].astype(np.float32)
    ds = capi.Dataset.from_mat(X, DS_PARAMS).set_field("label", y)
    b = capi.Booster(ds, _params("regression"))
    b.update_one_iter()
    text = b.save_model_to_string().replace("objective=regression", "objective=some_future_objective")
    with pytest.raises(capi.LightGBMError):
        capi.Booster(model_str=text)
    loaded = capi.Booster(model_str=b.save_model_to_string())
    with pytest.raises(capi.LightGBMError):       # no training data behind a loaded booster: -1, not a segfault
        loaded.get_eval(0)
    with pytest.raises(capi.LightGBMError):
        loaded.get_predict(0)
