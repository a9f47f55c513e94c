"""CPU-side checks: the C-ABI library loads and exports every declared symbol,
the module mirrors the reference's interface, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from multi_view_stereonet_b200 import _lib, synthetic
from tests.conftest import REPO


def _ensure_built():
    from multi_view_stereonet_b200 import build
    build.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    header = open(os.path.join(REPO, "include", "b200mvs.h")).read()
    declared = set(re.findall(r"B200MVS_API[^;]*?\b(b200mvs_\w+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.load().b200mvs_version()


def test_state_dict_names_match_reference_layout(gta_state):
    from multi_view_stereonet_b200 import MultiViewStereoNet
    net = MultiViewStereoNet()
    assert net.num_levels == 5
    missing = net.load_state_dict(gta_state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert len(net.state_dict()) == 226
    assert sum(p.numel() for p in net.parameters()) == 608614   # pretrained/gta_sfm_150epochs/logs.txt:4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(gta_state):
    from multi_view_stereonet_b200 import MultiViewStereoNet
    _ensure_built()
    net = MultiViewStereoNet()
    net.load_state_dict(gta_state)
    inputs = synthetic.make_inputs(64, 80, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(*inputs, 8, True, [True] * 5)
    # the C ABI itself refuses too
    handle = ctypes.c_void_p()
    names = (ctypes.c_char_p * 0)()
    rc = _lib.load().b200mvs_create(0, 0, names, _lib.ptr_array([]), (ctypes.c_int64 * 0)(), ctypes.byref(handle))
    assert rc != 0 and handle.value is None


def test_reference_asserts_are_kept(gta_state):
    from multi_view_stereonet_b200 import MultiViewStereoNet
    net = MultiViewStereoNet()
    left, K, T, right = synthetic.make_inputs(64, 80, 1, 1)
    with pytest.raises(AssertionError):
        net(left[:4], K, T, right, 8, True, [True] * 5)      # multi_view_stereonet.py:549
    with pytest.raises(AssertionError):
        net(left, K[:4], T, right, 8, True, [True] * 5)      # multi_view_stereonet.py:548


def test_synthetic_pyramid_and_intrinsics():
    left, K, T, right = synthetic.make_inputs(68, 90, 3, 2)
    assert [tuple(t.shape[-2:]) for t in left] == [(68, 90), (34, 45), (17, 23), (9, 12), (5, 6)]
    assert len(T) == 3 and len(right) == 3 and tuple(T[0].shape) == (2, 4, 4)
    # translations are divided by the FIRST view's baseline (multi_view_stereonet_utils.py:596-604)
    assert abs(float(T[0][0, :3, 3].norm()) - 1.0) < 1e-6
    assert float(T[1][0, :3, 3].norm()) > 1.5
    sx = 45 / 90
    assert abs(float(K[1][0, 0, 2]) - (sx * (float(K[0][0, 0, 2]) + 0.5) - 0.5)) < 1e-6
    # item i depends only on seed + i (ranks build their own shard)
    a = synthetic.make_inputs(64, 80, 1, 4)
    b = synthetic.make_inputs(64, 80, 1, 2, first_item=2)
    assert torch.equal(a[0][0][2:], b[0][0])


def test_algorithmic_work_matches_survey_totals():
    """bench.py's roofline arithmetic (MACs and layerwise-compulsory bytes per depthmap) against the totals SURVEY.md
    8d states for the BASELINE configurations."""
    import bench
    for (rows, cols, views, hyps), (gmac, mbytes) in {
        (64, 80, 1, 8): (0.455, 14.1),
        (512, 640, 1, 64): (39.13, 1081.0),
        (512, 640, 4, 64): (76.58, 1816.0),
        (1024, 1280, 4, 128): (489.8, 10494.0),
    }.items():
        mac, byt, P = bench.algorithmic_work(rows, cols, views, hyps)
        assert abs(mac / 1e9 - gmac) <= 0.002 * gmac + 0.001, (rows, cols, views, hyps, mac / 1e9)
        assert abs(byt / 1e6 - mbytes) <= 0.002 * mbytes + 0.05, (rows, cols, views, hyps, byt / 1e6)
        assert P[0] == rows * cols and len(P) == 5


def test_evaluation_has_no_cpu_path():
    """The post-processing row fails loudly on CPU tensors, like the forward."""
    import pytest as _pytest
    from multi_view_stereonet_b200 import evaluation as ev
    with _pytest.raises(RuntimeError):
        ev.idepthmap_to_depthmap(torch.rand(1, 1, 4, 4), torch.ones(1))
