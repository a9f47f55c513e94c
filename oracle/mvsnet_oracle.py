"""CPU oracle for the MultiViewStereoNet depth-inference hot path.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product path (`multi_view_stereonet_b200`) never
does, and fails loudly when its CUDA library is missing.

It is a stage-by-stage restatement, in plain `torch` CPU tensor ops, of the
algorithm in the reference repository (paths relative to /root/reference):

  multi_view_stereonet/multi_view_stereonet.py   (the network, whole file)
  stereo/image_predictor.py:120-209, 400-523     (disparity->idepth, homography, warp)
  utils/resnet.py:10-18, 62-109                  (conv3x3, SimpleBasicBlock)

Parity pin: `tests/golden/*.npz` hold outputs of the reference's own eager
model (imported from /root/reference by `tests/golden/make_golden.py`, with the
reference's pretrained GTA-SfM weights) on the seeded inputs of
`multi_view_stereonet_b200.synthetic`; `tests/test_oracle.py` checks this file
against every one of them.  The reference has no tests or golden vectors of its
own, so those reference-run fixtures are the pin.

Every function takes a `dtype`; float32 mirrors the reference's arithmetic,
float64 gives a higher-precision "truth" used to judge which of two float32
implementations is closer.
"""
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2      # multi_view_stereonet.py:64, 323, 411, 455
GN_GROUPS = 4          # multi_view_stereonet.py:25-31 (32 // 8)
GN_EPS = 1e-5          # torch.nn.GroupNorm default
DILATIONS = (1, 2, 4, 8, 1, 1)   # multi_view_stereonet.py:457


# ----------------------------------------------------------------------------
# Layers
# ----------------------------------------------------------------------------
def _w(sd, name, dtype):
    t = sd.get(name)
    return None if t is None else t.to(dtype)


def gn_lrelu(x, sd, name, dtype):
    """GroupNorm(4, 32) + LeakyReLU(0.2); stats per sample over 8 channels x all
    spatial (x depth for 5-D input).  multi_view_stereonet.py:25-31, 347-350."""
    y = F.group_norm(x, GN_GROUPS, _w(sd, name + ".weight", dtype), _w(sd, name + ".bias", dtype), GN_EPS)
    return F.leaky_relu(y, LRELU_SLOPE)


def res_block(x, sd, name, dilation, dtype):
    """lrelu(GN(conv3x3_dilated(x))) + x -- one conv, no trailing activation.
    utils/resnet.py:93-109 with the overrides at multi_view_stereonet.py:61-66."""
    y = F.conv2d(x, _w(sd, name + ".conv1.weight", dtype), _w(sd, name + ".conv1.bias", dtype),
                 stride=1, padding=dilation, dilation=dilation)
    return gn_lrelu(y, sd, name + ".bn1", dtype) + x


def feature_network(image, sd, dtype, prefix="left_feature_extractor"):
    """FeatureNetwork.forward, multi_view_stereonet.py:109-129: four stride-2 5x5
    convs with no bias and no activation, six residual blocks, a biased 3x3 conv.
    Returns [image, conv0, conv1, conv2, conv_final(...)]."""
    pyr = [image]
    x = image
    for i in range(4):
        x = F.conv2d(x, _w(sd, f"{prefix}.conv{i}.weight", dtype), None, stride=2, padding=2)
        if i < 3:
            pyr.append(x)
    for i in range(6):
        x = res_block(x, sd, f"{prefix}.res{i}", 1, dtype)
    x = F.conv2d(x, _w(sd, f"{prefix}.conv_final.weight", dtype), _w(sd, f"{prefix}.conv_final.bias", dtype), padding=1)
    pyr.append(x)
    return pyr


def feature_refiner(image, features, sd, dtype, prefix="right_feature_extractor.refiner"):
    """FeatureRefiner.forward, multi_view_stereonet.py:424-440."""
    x = torch.cat([image, features], 1)
    x = F.conv2d(x, _w(sd, prefix + ".conv0.weight", dtype), _w(sd, prefix + ".conv0.bias", dtype), padding=1)
    x = gn_lrelu(x, sd, prefix + ".bn0", dtype)
    x = res_block(x, sd, prefix + ".res0", 1, dtype)
    delta = F.conv2d(x, _w(sd, prefix + ".conv_final.weight", dtype), _w(sd, prefix + ".conv_final.bias", dtype), padding=1)
    return features + delta


def idepthmap_refiner(guide, idepth_scaled, sd, prefix, dtype):
    """IDepthmapRefiner.forward, multi_view_stereonet.py:468-484."""
    x = torch.cat([guide, idepth_scaled], 1)
    x = F.conv2d(x, _w(sd, prefix + ".conv0.weight", dtype), _w(sd, prefix + ".conv0.bias", dtype), padding=1)
    x = gn_lrelu(x, sd, prefix + ".bn0", dtype)
    for i, d in enumerate(DILATIONS):
        x = res_block(x, sd, f"{prefix}.res{i}", d, dtype)
    delta = F.conv2d(x, _w(sd, prefix + ".conv_final.weight", dtype), _w(sd, prefix + ".conv_final.bias", dtype), padding=1)
    return F.relu(idepth_scaled + delta)


def cost_volume_filter(volume, sd, dtype, prefix="volume_filter4"):
    """CostVolumeFilter.forward, multi_view_stereonet.py:341-353."""
    x = volume
    for i in range(4):
        x = F.conv3d(x, _w(sd, f"{prefix}.conv{i}.weight", dtype), _w(sd, f"{prefix}.conv{i}.bias", dtype), padding=1)
        x = gn_lrelu(x, sd, f"{prefix}.bn{i}", dtype)
    x = F.conv3d(x, _w(sd, f"{prefix}.conv4.weight", dtype), _w(sd, f"{prefix}.conv4.bias", dtype), padding=1)
    return x[:, 0]


# ----------------------------------------------------------------------------
# Geometry
# ----------------------------------------------------------------------------
def pixel_grid(rows, cols, dtype):
    """Homogeneous integer pixel coordinates (3, rows*cols), x fastest.
    image_predictor.py:139-146, 483-490."""
    y, x = torch.meshgrid(torch.arange(rows, dtype=dtype), torch.arange(cols, dtype=dtype), indexing="ij")
    return torch.stack([x.reshape(-1), y.reshape(-1), torch.ones(rows * cols, dtype=dtype)], 0)


def homography_warp(H, image):
    """HomographyImagePredictor.forward, image_predictor.py:470-523.

    H: (N,3,3) left->right pixel homography; image (N,C,h,w).  Returns the
    bilinear, border-clamped resampling and the out-of-image mask (True =
    invalid).  The coordinate round trip through grid_sample's normalised
    convention is kept (pixel centre convention 2(p+0.5)/size-1, :506-510)."""
    n, _, rows, cols = image.shape
    p = H @ pixel_grid(rows, cols, image.dtype).unsqueeze(0)
    px = p[:, 0] / p[:, 2]
    py = p[:, 1] / p[:, 2]
    u = ((px + 0.5) * 2.0) / cols - 1.0
    v = ((py + 0.5) * 2.0) / rows - 1.0
    mask = (u.abs() > 1.0) | (v.abs() > 1.0)
    grid = torch.stack([u, v], -1).reshape(n, rows, cols, 2)
    out = F.grid_sample(image, grid, mode="bilinear", padding_mode="border", align_corners=False)
    return out, mask.reshape(n, 1, rows, cols)


def plane_sweep_warp(image, H):
    """PlaneSweepWarper.forward, multi_view_stereonet.py:205-235.
    image (B,C,h,w), H (B,D,3,3) -> volume (B,C,D,h,w) zeroed where invalid, mask (B,1,D,h,w)."""
    b, c, rows, cols = image.shape
    d = H.shape[1]
    img = image.unsqueeze(1).expand(b, d, c, rows, cols).reshape(b * d, c, rows, cols)
    out, mask = homography_warp(H.reshape(b * d, 3, 3), img)
    out = out.reshape(b, d, c, rows, cols).permute(0, 2, 1, 3, 4)
    mask = mask.reshape(b, d, 1, rows, cols).permute(0, 2, 1, 3, 4)
    return out * (~mask).to(out.dtype), mask


def normalize_pose(T):
    """Per-view baseline normalisation, multi_view_stereonet.py:566-571."""
    T = T.clone()
    baseline = T[:, :3, 3].pow(2).sum(1).sqrt()
    T[:, :3, 3] = T[:, :3, 3] / baseline.unsqueeze(1)
    return T, baseline


def disparity_to_idepth(K, T_right_in_left, disparity):
    """image_predictor.py:120-209: least-squares idepth along the epipolar line
    for a (non-rectified) disparity.  K (B,4,4), T (B,4,4), disparity (B,1,h,w)."""
    b, _, rows, cols = disparity.shape
    dtype = disparity.dtype
    pix = pixel_grid(rows, cols, dtype).unsqueeze(0).expand(b, 3, rows * cols)
    Kinv = torch.linalg.inv(K)
    T_lr = torch.linalg.inv(T_right_in_left)
    KRK = K[:, :3, :3] @ T_lr[:, :3, :3] @ Kinv[:, :3, :3]
    Kt = (K @ T_lr)[:, :3, 3].unsqueeze(2)                     # (B,3,1)
    disp = disparity.reshape(b, -1)

    p_inf = KRK @ pix
    w = p_inf[:, 2]                                            # == KRK[2] . (x, y, 1), :189-191
    p_inf = p_inf[:, :2] / p_inf[:, 2:3]
    p_far = KRK @ (pix * 1e2) + Kt                             # :172-176
    p_far = p_far[:, :2] / p_far[:, 2:3]
    epi = p_far - p_inf
    norm = epi.pow(2).sum(1).sqrt()
    epi = epi / (norm + 1e-6).unsqueeze(1)
    invalid = norm < 1e-6

    A = Kt[:, :2] - Kt[:, 2:3] * (p_inf + disp.unsqueeze(1) * epi)    # :194-195
    rhs = (w * disp).unsqueeze(1) * epi                               # :197-198
    idepth = (A * rhs).sum(1) / (A * A).sum(1)
    idepth = idepth * (~invalid).to(dtype)
    return idepth.reshape(b, 1, rows, cols)


def idepth_samples(T_norm, K4, rows, cols, num):
    """create_idepth_samples, multi_view_stereonet.py:131-165."""
    b = T_norm.shape[0]
    dtype = T_norm.dtype
    disp = torch.full((b, 1, rows, cols), float(num - 1), dtype=dtype)
    m = disparity_to_idepth(K4, T_norm, disp).reshape(b, -1)
    m = m * (m > 0).to(dtype)
    mx = m.sum(1) / (m > 0).sum(1)
    mx = torch.where(mx > 2.0, torch.full_like(mx, 2.0), mx)
    tz = T_norm[:, 2, 3]
    mx = torch.where(1.0 / mx < tz, 1.0 / tz, mx)
    delta = mx / (num - 1)
    return torch.arange(num, dtype=dtype).unsqueeze(0) * delta.unsqueeze(1)


def plane_sweep_homographies(T_norm, K, samples):
    """create_plane_sweep_homographies + get_fronto_parallel_homography,
    multi_view_stereonet.py:167-194, image_predictor.py:446-459:
    H_d = K3 (R + t idepth_d e3^T) K3^-1 with (R, t) = inverse(T_right_in_left)."""
    T_lr = torch.linalg.inv(T_norm)
    R, t = T_lr[:, :3, :3], T_lr[:, :3, 3]
    K3 = K[:, :3, :3]
    M = R.unsqueeze(1).repeat(1, samples.shape[1], 1, 1)
    M[:, :, :, 2] = M[:, :, :, 2] + t.unsqueeze(1) * samples.unsqueeze(2)
    return K3.unsqueeze(1) @ M @ torch.linalg.inv(K3).unsqueeze(1)


def incremental_right_features(T_norm, K_pyr, right_pyr, samples, sd, dtype, force_masks=None):
    """IncrementalFastGeometryAwareFeatureNetwork.forward, multi_view_stereonet.py:247-300.
    Returns the masked feature volume (B,32,D,h,w), the mask (B,D,h,w) and stage tensors.

    force_masks = {"l0": (B,1,H,W) bool, "l4": (B,D,h,w) bool} replaces the two
    thresholded out-of-image masks by given decisions (test-only: lets a parity
    test separate arithmetic error from knife-edge tie-breaks, see mask_margins)."""
    d = samples.shape[1]
    H0 = plane_sweep_homographies(T_norm, K_pyr[0], samples[:, 0:1])
    img0, mask0 = homography_warp(H0[:, 0], right_pyr[0])
    if force_masks is not None:
        mask0 = force_masks["l0"]
    img0 = (img0 * (~mask0).to(dtype)).unsqueeze(2)
    feats = [feature_network(img0[:, :, 0], sd, dtype)[-1]]
    H = plane_sweep_homographies(T_norm, K_pyr[-1], samples)
    image_volume, mask_volume = plane_sweep_warp(right_pyr[-1], H)
    if force_masks is not None:
        b, c = image_volume.shape[:2]
        raw, _ = homography_warp(H.reshape(-1, 3, 3), right_pyr[-1].unsqueeze(1).expand(b, d, *right_pyr[-1].shape[1:])
                                 .reshape(b * d, *right_pyr[-1].shape[1:]))
        mask_volume = force_masks["l4"].unsqueeze(1)
        image_volume = raw.reshape(b, d, c, *raw.shape[-2:]).permute(0, 2, 1, 3, 4) * (~mask_volume).to(dtype)
    for i in range(1, d):
        H_inc = torch.linalg.inv(H[:, i - 1]) @ H[:, i]
        warped, wmask = homography_warp(H_inc, feats[-1])
        warped = warped * (~wmask).to(dtype)
        feats.append(feature_refiner(image_volume[:, :, i], warped, sd, dtype))
    volume = torch.stack(feats, 2)
    masked = volume * (~mask_volume).to(dtype)
    stages = {"H0": H0, "H": H, "right_image0_warped": img0[:, :, 0], "right_image_volume": image_volume,
              "right_feature_volume_unmasked": volume, "l0_mask": mask0, "l4_mask": mask_volume[:, 0]}
    return masked, mask_volume[:, 0], stages


def mask_margins(T_right_in_left, K_pyr, num, size0, size4):
    """Distance (in pixels, float64) of every warped coordinate from the
    out-of-image threshold of image_predictor.py:513-515, for the two thresholded
    warps of one view: the full-resolution warp (B,H,W) and the 1/16-scale sweep
    (B,D,h,w).  A float32 implementation can only be expected to reproduce the
    reference's mask bit where the margin exceeds its coordinate rounding error
    (~1e-4 px at 640 px, ~1e-5 px at 40 px)."""
    T_norm, _ = normalize_pose(T_right_in_left.double())
    K0, K4 = K_pyr[0].double(), K_pyr[-1].double()
    samples = idepth_samples(T_norm, K4, size4[0], size4[1], num)

    def margin(H, rows, cols):
        p = H @ pixel_grid(rows, cols, torch.float64).unsqueeze(0)
        px, py = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]
        m = torch.stack([(px + 0.5).abs(), (px - (cols - 0.5)).abs(), (py + 0.5).abs(), (py - (rows - 0.5)).abs()])
        return m.min(0).values.reshape(-1, rows, cols)

    b = T_norm.shape[0]
    m0 = margin(plane_sweep_homographies(T_norm, K0, samples[:, :1])[:, 0], *size0)
    m4 = margin(plane_sweep_homographies(T_norm, K4, samples).reshape(-1, 3, 3), *size4).reshape(b, num, *size4)
    return m0, m4


def soft_argmin(cost, samples):
    """extract_idepthmap, multi_view_stereonet.py:486-492 (beta = 1)."""
    probs = F.softmin(cost, dim=1)
    return (probs * samples[:, :, None, None]).sum(1, keepdim=True)


def upsample_bilinear(x, size):
    """Upsampler(refine=False, relu=False), multi_view_stereonet.py:372-380."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def upsample_mask(mask, size, dtype):
    """MaskUpsampler.forward, multi_view_stereonet.py:389-396."""
    return F.interpolate(mask.to(dtype), size=size, mode="bilinear", align_corners=False) > 0.5


# ----------------------------------------------------------------------------
# The hot path
# ----------------------------------------------------------------------------
def forward(sd, left_image_pyr, K_pyr, T_right_in_lefts, right_image_pyrs, num_idepth_samples,
            do_cost_volume_filter=True, do_refiners=(True,) * 5, dtype=torch.float32, return_stages=False,
            force_masks=None):
    """MultiViewStereoNet.forward, multi_view_stereonet.py:538-695.  Same inputs,
    same output dict; `sd` is the reference state dict.  `force_masks` is a
    per-view list for `incremental_right_features` (test-only)."""
    assert len(K_pyr) == 5 and len(left_image_pyr) == 5
    cv = lambda t: t.to(dtype)
    left_image_pyr = [cv(t) for t in left_image_pyr]
    K_pyr = [cv(t) for t in K_pyr]
    T_right_in_lefts = [cv(t) for t in T_right_in_lefts]
    right_image_pyrs = [[cv(t) for t in p] for p in right_image_pyrs]
    num = num_idepth_samples
    views = len(T_right_in_lefts)
    stages = {}

    left_feature_pyr = feature_network(left_image_pyr[0], sd, dtype)
    feat4 = left_feature_pyr[-1]
    b, _, rows4, cols4 = feat4.shape
    fx = [K[:, 0, 0].reshape(b, 1, 1, 1) for K in K_pyr]

    raw_sum = torch.zeros(b, 1, rows4, cols4, dtype=dtype)
    ref_sum = torch.zeros(b, 1, rows4, cols4, dtype=dtype)
    mask_sum = torch.zeros(b, num, rows4, cols4, dtype=dtype)
    for v in range(views):
        T_norm, baseline = normalize_pose(T_right_in_lefts[v])
        samples = idepth_samples(T_norm, K_pyr[-1], left_image_pyr[-1].shape[-2], left_image_pyr[-1].shape[-1], num)
        right_volume, mask, st = incremental_right_features(
            T_norm, K_pyr, right_image_pyrs[v], samples, sd, dtype,
            None if force_masks is None else force_masks[v])
        cost = (feat4.unsqueeze(2) - right_volume).abs() * (~mask).unsqueeze(1).to(dtype)   # :587-592
        if do_cost_volume_filter:
            filtered = cost_volume_filter(cost, sd, dtype)
        else:
            filtered = cost.pow(2).sum(1).sqrt()                                         # :598
        raw = soft_argmin(filtered, samples)
        if do_refiners[4]:
            guide = torch.cat([left_image_pyr[-1], feat4], 1)
            refined = idepthmap_refiner(guide, raw * fx[4], sd, "refiner4", dtype) / fx[4]
            raw = raw / baseline.reshape(b, 1, 1, 1)
            refined = refined / baseline.reshape(b, 1, 1, 1)
        else:
            # Reference quirk (:613-619): the two in-place divisions hit one
            # aliased tensor, so the baseline is divided out twice.
            raw = raw / baseline.reshape(b, 1, 1, 1) / baseline.reshape(b, 1, 1, 1)
            refined = raw
        raw_sum += raw
        ref_sum += refined
        mask_sum += mask.to(dtype)
        if return_stages:
            stages[f"view{v}"] = dict(st, idepth_samples=samples, baseline=baseline, mask=mask,
                                      right_feature_volume=right_volume, cost=cost, cost_filtered=filtered,
                                      idepth4_raw=raw, idepth4=refined)

    raw4 = raw_sum / views
    idepth = ref_sum / views
    mask = (mask_sum / views) > 0.5
    idepths, priors, masks = [idepth], [raw4], [mask]
    for lvl in (3, 2, 1, 0):
        size = left_image_pyr[lvl].shape[-2:]
        prior = upsample_bilinear(idepth, size)
        mask = upsample_mask(mask, size, dtype)
        if do_refiners[lvl]:
            guide = left_image_pyr[0] if lvl == 0 else torch.cat([left_image_pyr[lvl], left_feature_pyr[lvl]], 1)
            idepth = idepthmap_refiner(guide, prior * fx[lvl], sd, f"refiner{lvl}", dtype) / fx[lvl]
        else:
            idepth = prior
        idepths.insert(0, idepth)
        priors.insert(0, prior)
        masks.insert(0, mask)

    out = {"left_idepthmap_pyr": idepths, "left_idepthmap_raw_pyr": priors, "left_idepthmap_mask_pyr": masks}
    if return_stages:
        stages["left_feature_pyr"] = left_feature_pyr
        out["stages"] = stages
    return out
