"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module (see the header of gbm_oracle.cpp).  Parity unpinned: there is no runnable LightGBM
3.2.110 in this environment.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gbm_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def effective_cpus():
    """Host threads this process may really use: min(affinity mask, cgroup CPU quota).  The GPU box reports
    128 logical CPUs but its container quota is 16 cores — 128 OpenMP threads there run ~100x slower."""
    n = len(os.sched_getaffinity(0))
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, p = f.read().split()
            if q != "max":
                n = max(1, min(n, int(int(q) / int(p))))
    except Exception:
        pass
    env = os.environ.get("OMP_NUM_THREADS")
    if env:
        n = max(1, min(n, int(env)))
    return n


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_dataset_create.restype = C.c_void_p
        L.orc_dataset_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        L.orc_dataset_create_from_bins.restype = C.c_void_p
        L.orc_dataset_create_from_bins.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        L.orc_dataset_free.argtypes = [C.c_void_p]
        L.orc_dataset_num_used.argtypes = [C.c_void_p]
        L.orc_dataset_bins.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dataset_bins16.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dataset_bin_to_cat.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_dataset_feature_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_dataset_upper_bounds.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_dataset_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.orc_booster_create.restype = C.c_void_p
        L.orc_booster_create.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_booster_free.argtypes = [C.c_void_p]
        L.orc_booster_update.argtypes = [C.c_void_p]
        L.orc_booster_reset_learning_rate.argtypes = [C.c_void_p, C.c_double]
        L.orc_booster_num_trees.argtypes = [C.c_void_p]
        L.orc_booster_model_string.restype = C.c_char_p
        L.orc_booster_model_string.argtypes = [C.c_void_p]
        L.orc_booster_trace_len.argtypes = [C.c_void_p]
        L.orc_booster_trace.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_booster_scores.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_booster_gradients.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_booster_predict_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_booster_hist_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_histogram.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_best_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.orc_random_sample.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_count_cardinality.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_set_num_threads(C.c_int(effective_cpus()))
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


TRACE_COLS = ["tree", "split", "leaf", "feature", "threshold_bin", "left_count", "right_count", "default_left",
              "gain", "left_sum_g", "left_sum_h", "right_sum_g", "right_sum_h", "left_out", "right_out", "smaller_rows"]


class OracleDataset:
    def __init__(self, X, params="max_bin=255", rank_rows=None):
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.F = X.shape
        rr = None
        nr = 1
        if rank_rows is not None:
            rr = np.ascontiguousarray(rank_rows, dtype=np.int32)
            nr = len(rr)
            assert int(rr.sum()) == self.n
        self.h = lib().orc_dataset_create(_p(X), self.n, self.F, params.encode(), nr, _p(rr))
        self._keep = []

    @classmethod
    def from_bins(cls, bins, infos, uppers, minmax, params="max_bin=255", rank_rows=None):
        """Dataset over pre-computed bins (row-major uint8 [n][F]); infos[f] = feature_info dict, uppers[f] = bin upper bounds,
        minmax[f] = (min, max) of the sampled values.  Mid-scale parity tests build it from the product's downloaded bins."""
        self = cls.__new__(cls)
        bins = np.ascontiguousarray(bins, dtype=np.uint8)
        self.n, self.F = bins.shape
        meta = np.zeros((self.F, 5), dtype=np.int32)
        off = np.zeros(self.F + 1, dtype=np.int32)
        for f, inf in enumerate(infos):
            meta[f] = [inf["num_bin"], inf["missing_type"], inf["default_bin"], inf["most_freq_bin"], int(inf["is_trivial"])]
            off[f + 1] = off[f] + len(uppers[f])
        up = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.float64) for u in uppers]) if self.F else np.zeros(0))
        mm = np.ascontiguousarray(np.asarray(minmax, dtype=np.float64).reshape(self.F, 2))
        rr, nr = None, 1
        if rank_rows is not None:
            rr = np.ascontiguousarray(rank_rows, dtype=np.int32)
            nr = len(rr)
            assert int(rr.sum()) == self.n
        self.h = lib().orc_dataset_create_from_bins(_p(bins), self.n, self.F, _p(meta), _p(up), _p(off), _p(mm), params.encode(), nr, _p(rr))
        self._keep = []
        return self

    def set_field(self, name, arr):
        dt = {"label": np.float32, "weight": np.float32, "init_score": np.float64, "group": np.int32}[name]
        a = np.ascontiguousarray(arr, dtype=dt)
        rc = lib().orc_dataset_set_field(self.h, name.encode(), _p(a), len(a))
        assert rc == 0
        return self

    def bins(self):
        out = np.zeros((self.n, self.F), dtype=np.uint8)
        lib().orc_dataset_bins(self.h, _p(out))
        return out

    def bins16(self):
        out = np.zeros((self.n, self.F), dtype=np.uint16)
        lib().orc_dataset_bins16(self.h, _p(out))
        return out

    def bin_to_cat(self, f):
        out = np.zeros(65536, dtype=np.int32)
        k = lib().orc_dataset_bin_to_cat(self.h, f, _p(out))
        return out[:k].copy()

    def feature_info(self, f):
        info = np.zeros(5, dtype=np.int32)
        lib().orc_dataset_feature_info(self.h, f, _p(info))
        return dict(num_bin=int(info[0]), missing_type=int(info[1]), default_bin=int(info[2]),
                    most_freq_bin=int(info[3]), is_trivial=bool(info[4]))

    def upper_bounds(self, f):
        out = np.zeros(32768, dtype=np.float64)
        k = lib().orc_dataset_upper_bounds(self.h, f, _p(out))
        return out[:k].copy()

    def close(self):
        if self.h:
            lib().orc_dataset_free(self.h)
            self.h = None


class OracleBooster:
    def __init__(self, ds, params):
        self.ds = ds
        self.h = lib().orc_booster_create(ds.h, params.encode())
        if not self.h:
            raise ValueError("oracle: unsupported objective in: " + params)
        self.K = 1
        for tok in params.split():
            if tok.startswith("num_class=") and any(("objective=" + o) in params for o in ("multiclass", "softmax", "multiclassova", "ova", "ovr")):
                self.K = int(tok.split("=")[1])

    def update(self):
        return bool(lib().orc_booster_update(self.h))

    def train(self, iters):
        done = 0
        for _ in range(iters):
            if self.update():
                break
            done += 1
        return done

    def reset_learning_rate(self, lr):
        lib().orc_booster_reset_learning_rate(self.h, float(lr))

    def model_string(self):
        return lib().orc_booster_model_string(self.h).decode()

    def trace(self):
        n = lib().orc_booster_trace_len(self.h)
        out = np.zeros((n, 16), dtype=np.float64)
        if n:
            lib().orc_booster_trace(self.h, _p(out))
        return out

    def scores(self):
        out = np.zeros(self.K * self.ds.n, dtype=np.float64)
        lib().orc_booster_scores(self.h, _p(out))
        return out

    def gradients(self):
        g = np.zeros(self.K * self.ds.n, dtype=np.float32)
        h = np.zeros(self.K * self.ds.n, dtype=np.float32)
        lib().orc_booster_gradients(self.h, _p(g), _p(h))
        return g, h

    def predict_raw(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.zeros((X.shape[0], self.K), dtype=np.float64)
        lib().orc_booster_predict_raw(self.h, _p(X), X.shape[0], X.shape[1], _p(out))
        return out

    def hist_stats(self):
        s = C.c_double(0)
        c = C.c_longlong(0)
        lib().orc_booster_hist_stats(self.h, C.byref(s), C.byref(c))
        return s.value, c.value

    def close(self):
        if self.h:
            lib().orc_booster_free(self.h)
            self.h = None


def histogram(bins, g, h, idx=None):
    bins = np.ascontiguousarray(bins, dtype=np.uint8)
    n, F = bins.shape
    g = np.ascontiguousarray(g, dtype=np.float32)
    h = np.ascontiguousarray(h, dtype=np.float32)
    out = np.zeros((F, 256, 2), dtype=np.float64)
    if idx is not None:
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        cnt = len(idx)
    else:
        cnt = n
    lib().orc_histogram(_p(bins), n, F, _p(g), _p(h), _p(idx), cnt, _p(out))
    return out


SPLIT_COLS = ["gain", "threshold", "default_left", "left_count", "right_count", "lsg", "lsh", "rsg", "rsh", "lout",
              "rout", "splittable"]


def best_split(hist, num_bin, missing_type, default_bin, most_freq_bin, sum_g, sum_h, num_data, l1=0.0, l2=0.0,
               max_delta_step=0.0, min_gain_to_split=0.0, min_sum_hessian=1e-3, min_data_in_leaf=20):
    hist = np.ascontiguousarray(hist, dtype=np.float64)
    meta = np.array([num_bin, missing_type, default_bin, most_freq_bin], dtype=np.int32)
    cfg = np.array([l1, l2, max_delta_step, min_gain_to_split, min_sum_hessian, min_data_in_leaf], dtype=np.float64)
    out = np.zeros(12, dtype=np.float64)
    lib().orc_best_split(_p(hist), _p(meta), _p(cfg), float(sum_g), float(sum_h), int(num_data), _p(out))
    return dict(zip(SPLIT_COLS, out.tolist()))


def random_sample(seed, N, K):
    out = np.zeros(max(K, 1), dtype=np.int32)
    k = lib().orc_random_sample(seed, N, K, _p(out))
    return out[:k].copy()


def count_cardinality(ids):
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.zeros(max(len(ids), 1), dtype=np.int32)
    k = lib().orc_count_cardinality(_p(ids), len(ids), _p(out))
    return out[:k].tolist()
