"""LightGBM model-text v3 parsing + structural comparison (used by the parity tests and smoke()).

Format: SURVEY.md Appendix B.3; writer = csrc/model.h (HostModel::ToString)."""
import numpy as np

INT_KEYS = ("split_feature", "decision_type", "left_child", "right_child", "leaf_count", "internal_count", "cat_boundaries", "cat_threshold")
FLOAT_KEYS = ("split_gain", "threshold", "leaf_value", "leaf_weight", "internal_value", "internal_weight")


def parse_model(text):
    """-> dict(header=dict, trees=[dict], params=str).  Tree arrays become numpy arrays."""
    lines = text.split("\n")
    header, trees, params = {}, [], ""
    i = 0
    while i < len(lines) and not lines[i].startswith("Tree=") and lines[i] != "end of trees":
        if "=" in lines[i]:
            k, v = lines[i].split("=", 1)
            header[k] = v
        i += 1
    while i < len(lines):
        if lines[i].startswith("Tree="):
            t = {}
            i += 1
            while i < len(lines) and not lines[i].startswith("Tree=") and lines[i] != "end of trees":
                if "=" in lines[i]:
                    k, v = lines[i].split("=", 1)
                    if k in INT_KEYS:
                        t[k] = np.array([int(x) for x in v.split()], dtype=np.int64)
                    elif k in FLOAT_KEYS:
                        t[k] = np.array([float(x) for x in v.split()], dtype=np.float64)
                    elif k in ("num_leaves", "num_cat", "is_linear"):
                        t[k] = int(v)
                    else:
                        t[k] = v
                i += 1
            trees.append(t)
        elif lines[i] == "parameters:":
            j = i + 1
            while j < len(lines) and lines[j] != "end of parameters":
                j += 1
            params = "\n".join(lines[i + 1:j])
            i = j
        else:
            i += 1
    return dict(header=header, trees=trees, params=params)


def compare_models(a, b, value_tol=1e-5, gain_tol=1e-5, check_counts=True, allow_nan_direction_ties=False):
    """Asserts: identical tree sequence (split feature, threshold value, children, decision type, counts)
    and leaf values / split gains within the north-star tolerance (1e-5 relative, gains are printed %g).
    allow_nan_direction_ties: ignore the default_left bit of nodes on NaN-missing features.  In a leaf that holds no NaN row the two scan
    directions of such a feature have mathematically equal gains and any implementation's fp64 rounding picks the winner (DESIGN.md K5
    "Ties"); the direction is then irrelevant for the partition, and thresholds, children and row counts are still compared exactly —
    they would differ if a NaN row had reached the node."""
    assert len(a["trees"]) == len(b["trees"]), "number of trees differs: %d vs %d" % (len(a["trees"]), len(b["trees"]))
    for k in ("num_class", "num_tree_per_iteration", "max_feature_idx", "objective", "feature_infos"):
        assert a["header"].get(k) == b["header"].get(k), "header field %s differs: %r vs %r" % (k, a["header"].get(k), b["header"].get(k))
    for ti, (ta, tb) in enumerate(zip(a["trees"], b["trees"])):
        assert ta["num_leaves"] == tb["num_leaves"], "tree %d: num_leaves %d vs %d" % (ti, ta["num_leaves"], tb["num_leaves"])
        if ta["num_leaves"] > 1:
            for k in ("split_feature", "decision_type", "left_child", "right_child"):
                xa, xb = ta[k], tb[k]
                if k == "decision_type" and allow_nan_direction_ties:
                    nan_num = lambda d: ((d >> 2) & 3 == 2) & (d & 1 == 0)          # NaN-missing numerical nodes
                    xa = np.where(nan_num(xa), xa & ~2, xa)
                    xb = np.where(nan_num(xb), xb & ~2, xb)
                assert np.array_equal(xa, xb), "tree %d: %s differs\n%s\n%s" % (ti, k, ta[k], tb[k])
            assert np.array_equal(ta["threshold"], tb["threshold"]), "tree %d: thresholds differ" % ti
            assert ta.get("num_cat", 0) == tb.get("num_cat", 0), "tree %d: num_cat differs" % ti
            if ta.get("num_cat", 0) > 0:
                for k in ("cat_boundaries", "cat_threshold"):
                    assert np.array_equal(ta[k], tb[k]), "tree %d: %s differs\n%s\n%s" % (ti, k, ta[k], tb[k])
            if check_counts:
                for k in ("leaf_count", "internal_count"):
                    assert np.array_equal(ta[k], tb[k]), "tree %d: %s differs\n%s\n%s" % (ti, k, ta[k], tb[k])
            np.testing.assert_allclose(ta["split_gain"], tb["split_gain"], rtol=max(gain_tol, 2e-6), atol=1e-12, err_msg="tree %d split_gain" % ti)
            np.testing.assert_allclose(ta["leaf_weight"], tb["leaf_weight"], rtol=value_tol, atol=1e-9, err_msg="tree %d leaf_weight" % ti)
        np.testing.assert_allclose(ta["leaf_value"], tb["leaf_value"], rtol=value_tol, atol=1e-9, err_msg="tree %d leaf_value" % ti)
        if "shrinkage" in ta and "shrinkage" in tb:
            np.testing.assert_allclose(float(ta["shrinkage"]), float(tb["shrinkage"]), rtol=1e-9, err_msg="tree %d shrinkage" % ti)
    return True
