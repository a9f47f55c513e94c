"""Runs the CUDA path next to the CPU oracle and reports per-stage errors."""
import torch

from oracle import mvsnet_oracle as oracle
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
from tests._util import rel_linf

# Mask bits can only be expected to agree where the warped coordinate is further
# from the out-of-image threshold than float32 rounding of the coordinate
# (oracle.mask_margins; DESIGN.md "Knife-edge mask pixels").
L0_KNIFE_PX = 2e-4
L4_KNIFE_PX = 2e-5


def make_net(state, device="cuda"):
    net = MultiViewStereoNet()
    net.load_state_dict(state, strict=True)
    return net.to(device).eval()


def run_case(net, state, inputs, hyps, cvf=True, refiners=(True,) * 5, stages=True):
    """Returns (report dict, cuda outputs, oracle outputs)."""
    dev = torch.device("cuda")
    left_pyr, K_pyr, Ts, right_pyrs = inputs
    b = left_pyr[0].shape[0]
    views = len(Ts)
    n = b * views
    h0, w0 = left_pyr[0].shape[-2:]
    h4, w4 = left_pyr[4].shape[-2:]
    net.keep_stages(stages)
    with torch.no_grad():
        out = net(*synthetic.to_device(inputs, dev), hyps, cvf, list(refiners))
        torch.cuda.synchronize()
        ref = oracle.forward(state, *inputs, hyps, cvf, refiners, return_stages=True)
    rep = {}

    # ---- discrete decisions: the two thresholded masks per view ----
    l0 = net.get_stage("l0_mask", torch.uint8).view(b, views, h0, w0).bool().cpu()
    l4 = net.get_stage("l4_mask", torch.uint8).view(b, views, hyps, h4, w4).bool().cpu()
    force = []
    flips0 = flips4 = bad_flips = 0
    for v in range(views):
        st = ref["stages"][f"view{v}"]
        m0, m4 = oracle.mask_margins(Ts[v], K_pyr, hyps, (h0, w0), (h4, w4))
        d0 = l0[:, v] != st["l0_mask"][:, 0]
        d4 = l4[:, v] != st["l4_mask"]
        flips0 += int(d0.sum())
        flips4 += int(d4.sum())
        bad_flips += int((d0 & (m0 > L0_KNIFE_PX)).sum()) + int((d4 & (m4 > L4_KNIFE_PX)).sum())
        force.append({"l0": l0[:, v].unsqueeze(1), "l4": l4[:, v]})
    rep["mask_flips_l0"] = flips0
    rep["mask_flips_l4"] = flips4
    rep["mask_flips_not_knife_edge"] = bad_flips
    if flips0 or flips4:
        # Re-run the oracle with the CUDA path's tie-breaks so that the remaining
        # difference is arithmetic only.
        with torch.no_grad():
            ref = oracle.forward(state, *inputs, hyps, cvf, refiners, return_stages=True, force_masks=force)

    # ---- stages ----
    if stages:
        g = lambda name, dt=torch.float32: net.get_stage(name, dt).cpu()
        samples = g("idepth_samples").view(b, views, hyps)
        base = g("baseline").view(b, views)
        H0 = g("H0").view(b, views, 3, 3)
        H = g("H").view(b, views, hyps, 3, 3)
        warped0 = g("right_image0_warped").view(b, views, 3, h0, w0)
        vol = g("right_feature_volume").view(b, views, hyps, h4, w4, 32)
        costf = g("cost_filtered").view(b, views, hyps, h4, w4)
        rawv = g("idepth4_raw_views").view(b, views, h4, w4)
        for v in range(views):
            st = ref["stages"][f"view{v}"]
            rep[f"v{v}/idepth_samples"] = rel_linf(samples[:, v], st["idepth_samples"])
            rep[f"v{v}/baseline"] = rel_linf(base[:, v], st["baseline"])
            rep[f"v{v}/H0"] = rel_linf(H0[:, v], st["H0"][:, 0])
            rep[f"v{v}/H"] = rel_linf(H[:, v], st["H"])
            rep[f"v{v}/right_image0_warped"] = rel_linf(warped0[:, v], st["right_image0_warped"])
            rep[f"v{v}/right_feature_volume"] = rel_linf(vol[:, v].permute(0, 4, 1, 2, 3),
                                                         st["right_feature_volume_unmasked"])
            rep[f"v{v}/cost_filtered"] = rel_linf(costf[:, v], st["cost_filtered"])
            rep[f"v{v}/idepth4_raw"] = rel_linf(rawv[:, v] / st["baseline"].view(b, 1, 1)
                                                / (1.0 if refiners[4] else st["baseline"].view(b, 1, 1)),
                                                st["idepth4_raw"][:, 0])
        for lvl in range(1, 5):
            f = g(f"left_feature{lvl}").view(b, *left_pyr[lvl].shape[-2:], 32).permute(0, 3, 1, 2)
            rep[f"left_feature{lvl}"] = rel_linf(f, ref["stages"]["left_feature_pyr"][lvl])

    # ---- outputs ----
    for lvl in range(5):
        rep[f"idepth{lvl}"] = rel_linf(out["left_idepthmap_pyr"][lvl].cpu(), ref["left_idepthmap_pyr"][lvl])
        rep[f"raw{lvl}"] = rel_linf(out["left_idepthmap_raw_pyr"][lvl].cpu(), ref["left_idepthmap_raw_pyr"][lvl])
        rep[f"mask_mismatch{lvl}"] = int((out["left_idepthmap_mask_pyr"][lvl].cpu()
                                          != ref["left_idepthmap_mask_pyr"][lvl]).sum())
    return rep, out, ref


def format_report(rep):
    return "\n".join(f"  {k:34s} {v:.3e}" if isinstance(v, float) else f"  {k:34s} {v}" for k, v in rep.items())
