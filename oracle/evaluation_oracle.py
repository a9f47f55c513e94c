"""CPU oracle for the post-processing that follows the hot path in the reference's test.py.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT (see oracle/mvsnet_oracle.py): only `tests/` may import it.

Restated in numpy (own formulation) from, relative to /root/reference:
  test.py:41-71     get_depth_prediction_metrics
  test.py:167-186   get_groundtruth_depthmap (ground truth multiplied back by the baseline, depth limits per split)
  test.py:211-214   idepth / baseline, then 1 / x where x > 0
  test.py:222, 236  validity mask: ground truth and estimate strictly inside (min_depth, max_depth)

Parity pin: tests/golden/evaluation_small.npz holds what the reference's own functions return (their source is cut
out of test.py with `ast` by tests/golden/make_golden_evaluation.py -- the file itself needs matplotlib and datasets)
on seeded inputs; tests/test_evaluation.py checks every function here against it.
"""
import numpy as np

METRIC_KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def depth_limits(split):
    """(min_depth, max_depth) of test.py:167-186."""
    if "gta_sfm" in split:
        return 0.0, 1e3
    if "demon" in split:
        return 0.5, 10.0          # "Limits from DPSNet"
    raise ValueError(split)


def idepthmap_to_depthmap(idepthmap, baseline):
    """test.py:211-213 on float32 arrays: (B,1,H,W), (B,) -> idepth_est, depth_est."""
    idepthmap = np.asarray(idepthmap, dtype=np.float32)
    scale = np.asarray(baseline, dtype=np.float32).reshape(-1, 1, 1, 1)
    idepth = idepthmap / scale
    depth = idepth.copy()
    pos = depth > 0
    depth[pos] = np.float32(1.0) / depth[pos]
    return idepth, depth


def depth_metrics(depth_true, depth_est):
    """test.py:41-71 on already-masked 1-D float32 arrays; float64 means of the float32 per-pixel terms."""
    t = np.asarray(depth_true, dtype=np.float32)
    e = np.asarray(depth_est, dtype=np.float32)
    thresh = np.maximum(t / e, e / t)
    diff = t - e
    sq = diff * diff
    dl = np.log(t) - np.log(e)
    mean = lambda x: float(np.mean(x.astype(np.float64)))
    return {"abs_rel": mean(np.abs(diff) / t), "sq_rel": mean(sq / t), "rmse": float(np.sqrt(mean(sq))),
            "rmse_log": float(np.sqrt(mean(dl * dl))), "a1": mean(thresh < np.float32(1.25)),
            "a2": mean(thresh < np.float32(1.25 ** 2)), "a3": mean(thresh < np.float32(1.25 ** 3))}


def evaluate(idepthmap, baseline, depth_true_normalised, split):
    """Per batch item: (metrics dict or None when no pixel is valid, number of valid pixels)."""
    lo, hi = depth_limits(split)
    _, depth = idepthmap_to_depthmap(idepthmap, baseline)
    scale = np.asarray(baseline, dtype=np.float32).reshape(-1, 1, 1, 1)
    truth = np.asarray(depth_true_normalised, dtype=np.float32) * scale
    out = []
    for b in range(depth.shape[0]):
        t, e = truth[b].squeeze(), depth[b].squeeze()
        mask = (t > lo) & (t < hi)
        if mask.sum() <= 0:
            out.append((None, 0))
            continue
        mask = mask & (e > lo) & (e < hi)
        out.append((depth_metrics(t[mask], e[mask]) if mask.sum() > 0 else None, int(mask.sum())))
    return out
