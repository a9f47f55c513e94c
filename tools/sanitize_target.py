"""One forward at 512x640 (few hypotheses) and one at 64x80 for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic, image_predictor as ip
from multi_view_stereonet_b200 import multi_view_stereonet_utils as snu
sd, _ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net = net.cuda().eval()
with torch.no_grad():
    # CASES=big: only the multi-CTA-cluster shape (the full set takes ~40 s under either tool).
    cases = ((512, 640, 2, 6), (64, 80, 1, 8), (68, 90, 1, 5))
    if os.environ.get("CASES") == "big":
        cases = cases[:1]
    if os.environ.get("WIDE"):
        # the wide sweep (sweep_wide.cu) forced onto a multi-CTA shape: 2 chains x 11 CTAs, 3 dependent steps
        net.set_option("sweep", 1)
        cases = ((512, 640, 2, 4),)
    for rows, cols, views, hyps in cases:
        inp = synthetic.to_device(synthetic.make_inputs(rows, cols, views, 1), "cuda")
        out = net(*inp, hyps, True, [True] * 5)
        torch.cuda.synchronize()
        print(rows, cols, float(out["left_idepthmap_pyr"][0].mean()))
    batch = synthetic.make_raw_batch()
    inputs = snu.multi_view_unpack_batch(batch, torch.device("cuda"), 5)
    pred, mask = ip.ImagePredictor()(inputs["K_pyr"][0], inputs["T_right_in_left"][0],
                                     torch.rand(2, 1, 38, 51, device="cuda") * 3, inputs["right_image_pyr"][0][0])
    torch.cuda.synchronize()
    print("ok", float(pred.mean()))
    # post-processing row: odd pixel count (scalar path) and a 16-byte aligned one (vector path)
    from multi_view_stereonet_b200 import evaluation as ev
    for rows, cols in ((38, 51), (64, 80)):
        est = torch.rand(2, 1, rows, cols, device="cuda") + 0.05
        truth = torch.rand(2, 1, rows, cols, device="cuda") * 20
        _, depth, metrics = ev.evaluate_batch(est, torch.tensor([0.3, 0.5], device="cuda"), truth, "gta_sfm")
        torch.cuda.synchronize()
        print("eval", rows, cols, metrics[0]["num_valid"], round(metrics[0]["abs_rel"], 4))
