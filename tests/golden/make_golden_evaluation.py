"""Golden fixture for the post-processing row: runs the REFERENCE's own
`get_depth_prediction_metrics` and `get_groundtruth_depthmap` (test.py:41-71, 167-186) and the statements of its test
loop that turn the network output into depth (test.py:211-214, 218-236) on seeded inputs.  test.py cannot be imported
here (matplotlib, datasets), so the two functions are cut out of the file with `ast` and executed unmodified; the loop
statements are reproduced below (per-item broadcast made explicit).

    python tests/golden/make_golden_evaluation.py        -> tests/golden/evaluation_small.npz
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def reference_functions():
    path = os.path.join(REF, "test.py")
    tree = ast.parse(open(path).read())
    want = {"get_depth_prediction_metrics", "get_groundtruth_depthmap"}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return ns["get_depth_prediction_metrics"], ns["get_groundtruth_depthmap"]


def make_case(seed, B, rows, cols):
    g = torch.Generator().manual_seed(seed)
    baseline = torch.rand(B, generator=g) * 0.5 + 0.1                                    # metres
    depth_true = torch.rand(B, 1, rows, cols, generator=g) * 14.0                        # some beyond the DeMoN limits
    depth_true[torch.rand(B, 1, rows, cols, generator=g) < 0.1] = 0.0                    # holes in the ground truth
    depth_true_norm = depth_true / baseline.view(-1, 1, 1, 1)                            # as multi_view_unpack_batch holds it
    noise = 1.0 + 0.3 * (torch.rand(B, 1, rows, cols, generator=g) - 0.5)
    idepth = torch.where(depth_true_norm > 0, 1.0 / (depth_true_norm * noise + 1e-3), torch.zeros(()))
    idepth[torch.rand(B, 1, rows, cols, generator=g) < 0.05] = 0.0                       # relu zeros in the estimate
    if B > 2:
        depth_true_norm[2] = 0.0                                                         # an image without ground truth
    return idepth.float(), baseline.float(), depth_true_norm.float()


def main():
    metrics_fn, truth_fn = reference_functions()
    flat = {}
    for name, split, (seed, B, rows, cols) in (("gta", "gta_sfm_overlap0.5_test.txt", (5, 3, 37, 50)),
                                               ("demon", "demon_test.txt", (6, 3, 48, 64))):
        idepth, baseline, truth_norm = make_case(seed, B, rows, cols)
        flat[f"{name}_idepth"], flat[f"{name}_baseline"], flat[f"{name}_truth_norm"] = idepth, baseline, truth_norm
        # --- test.py:211-213; the reference runs batch size 1, where `/ inputs["baseline"]` is a per-item division ---
        inputs = {"baseline": baseline.clone(), "left_depthmap_true": truth_norm.clone()}
        outputs = {"left_idepthmap_pyr": [idepth.clone()]}
        batch_left_idepthmap_est = outputs["left_idepthmap_pyr"][0] / inputs["baseline"].view(-1, 1, 1, 1)
        batch_left_depthmap_est = outputs["left_idepthmap_pyr"][0] / inputs["baseline"].view(-1, 1, 1, 1)
        batch_left_depthmap_est[batch_left_depthmap_est > 0] = 1.0 / batch_left_depthmap_est[batch_left_depthmap_est > 0]
        flat[f"{name}_idepth_est"], flat[f"{name}_depth_est"] = batch_left_idepthmap_est, batch_left_depthmap_est
        rows_out, counts = [], []
        for idx in range(B):
            # test.py:216-236 (the reference runs with batch size 1: one image per call of get_groundtruth_depthmap)
            one = {"baseline": inputs["baseline"][idx:idx + 1], "left_depthmap_true": inputs["left_depthmap_true"][idx:idx + 1].clone()}
            left_depthmap_true, min_depth, max_depth = truth_fn(split, one, "file")
            mask = (left_depthmap_true > min_depth) & (left_depthmap_true < max_depth)
            if np.sum(mask) <= 0:
                rows_out.append([np.nan] * 7)
                counts.append(0)
                continue
            left_depthmap_est = batch_left_depthmap_est[idx, :, :, :].unsqueeze(0).cpu().numpy().squeeze()
            mask = mask & (left_depthmap_est > min_depth) & (left_depthmap_est < max_depth)
            m = metrics_fn(left_depthmap_true[mask], left_depthmap_est[mask])
            rows_out.append([float(m[k]) for k in ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")])
            counts.append(int(mask.sum()))
        flat[f"{name}_metrics"] = torch.tensor(rows_out, dtype=torch.float64)
        flat[f"{name}_counts"] = torch.tensor(counts)
        print(name, "valid pixels", counts, "abs_rel", [round(r[0], 5) for r in rows_out])
    path = os.path.join(HERE, "evaluation_small.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in flat.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
