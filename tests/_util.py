"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from multi_view_stereonet_b200 import synthetic, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity bar from BASELINE.json north_star: relative L-inf <= 1e-3 on the inverse
# depth maps, i.e. max|y - y_ref| / max|y_ref|.
REL_LINF_TOL = 1e-3


def rel_linf(y, ref):
    y = torch.as_tensor(y, dtype=torch.float64)
    ref = torch.as_tensor(ref, dtype=torch.float64)
    return float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def load_gta_state():
    """Pretrained GTA-SfM weights extracted from the reference archive by
    tests/golden/make_golden.py, with the shared-extractor aliases restored."""
    sd = weights.load_state_npz(os.path.join(GOLDEN, "gta_sfm_150epochs_state.npz"))
    fe = "left_feature_extractor."
    for k in [k for k in sd if k.startswith(fe)]:
        sd["right_feature_extractor.feature_extractor." + k[len(fe):]] = sd[k]
    return sd


def load_case(name):
    """Returns (golden dict, inputs, hyps, do_cvf, do_refiners) for a fixture."""
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    rows, cols, views, hyps, batch, smooth, cvf = [int(x) for x in z["meta"][:7]]
    refiners = [bool(x) for x in z["meta"][7:12]]
    pitch = float(z["pitch"][0]) if "pitch" in z else None
    inputs = synthetic.make_inputs(rows, cols, views, batch, smooth=bool(smooth), pitch=pitch)
    chk = np.array([float(inputs[0][0].double().sum()), float(inputs[3][-1][0].double().sum()),
                    float(inputs[0][4].double().abs().sum())])
    # The fixtures are only meaningful if this box regenerates identical inputs.
    np.testing.assert_allclose(chk, z["input_checksum"], rtol=1e-9, atol=1e-6)
    return z, inputs, hyps, bool(cvf), refiners


def unpack_mask(packed, shape):
    n = int(np.prod(shape))
    return np.unpackbits(packed)[:n].reshape(shape).astype(bool)
