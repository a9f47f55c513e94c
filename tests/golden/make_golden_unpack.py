"""Golden fixture for the input preparation: runs the REFERENCE's own `multi_view_unpack_batch`
(multi_view_stereonet/multi_view_stereonet_utils.py:541-641) on a seeded batch.  The module itself cannot be imported
here (it needs matplotlib), so the function's source is cut out of the reference file with `ast` and executed
unmodified against the reference's `utils.image_utils`.

    python tests/golden/make_golden_unpack.py        -> tests/golden/unpack_small.npz
"""
import ast
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from multi_view_stereonet_b200 import synthetic  # noqa: E402
from utils import image_utils  # noqa: E402  (the reference)


def reference_unpack(name="multi_view_unpack_batch"):
    path = os.path.join(REF, "multi_view_stereonet/multi_view_stereonet_utils.py")
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "image_utils": image_utils}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main_two_view():
    """Reference `unpack_batch` (:406-501) -> tests/golden/unpack2_small.npz"""
    out = reference_unpack("unpack_batch")(clone(synthetic.make_raw_batch_two_view()), torch.device("cpu"), 5)
    flat = {k: out[k] for k in ("baseline", "T_right_in_left", "T_left_in_right", "left_depthmap_true",
                                "left_idepthmap_true", "right_depthmap_true", "right_idepthmap_true")}
    for lvl in range(5):
        flat[f"K_pyr{lvl}"] = out["K_pyr"][lvl]
        flat[f"left_image_pyr{lvl}"] = out["left_image_pyr"][lvl]
        flat[f"right_image_pyr{lvl}"] = out["right_image_pyr"][lvl]
    path = os.path.join(HERE, "unpack2_small.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in flat.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def clone(batch):
    return {k: ([t.clone() if torch.is_tensor(t) else t for t in v] if isinstance(v, list) else
                (v.clone() if torch.is_tensor(v) else v)) for k, v in batch.items()}


def main():
    batch = synthetic.make_raw_batch()
    out = reference_unpack()(clone(batch), torch.device("cpu"), 5)
    flat = {}
    flat["baseline"] = out["baseline"]
    flat["left_depthmap_true"] = out["left_depthmap_true"]
    flat["left_idepthmap_true"] = out["left_idepthmap_true"]
    for lvl in range(5):
        flat[f"K_pyr{lvl}"] = out["K_pyr"][lvl]
        flat[f"left_image_pyr{lvl}"] = out["left_image_pyr"][lvl]
        for v in range(len(batch["right_image"])):
            flat[f"right_image_pyr{v}_{lvl}"] = out["right_image_pyr"][v][lvl]
    for v in range(len(batch["right_image"])):
        flat[f"T_right_in_left{v}"] = out["T_right_in_left"][v]
        flat[f"T_left_in_right{v}"] = out["T_left_in_right"][v]
        flat[f"right_idepthmap_true{v}"] = out["right_idepthmap_true"][v]
    path = os.path.join(HERE, "unpack_small.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in flat.items()})
    print("wrote", path, os.path.getsize(path), "bytes;", "baseline", out["baseline"].tolist(),
          "sizes", [tuple(t.shape[-2:]) for t in out["left_image_pyr"]])


if __name__ == "__main__":
    main()
    main_two_view()
