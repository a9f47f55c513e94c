"""CPU oracle for the reprojection layers of the reference's stereo/image_predictor.py.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT (see oracle/mvsnet_oracle.py): only `tests/` may import it.

Restated per pixel (own formulation, torch CPU ops) from, relative to /root/reference:
  stereo/image_predictor.py:36-118    DepthmapToPointCloud, PointCloudToPixel
  stereo/image_predictor.py:120-218   disparity_to_idepth / DisparityToIDepth
  stereo/image_predictor.py:220-273   IDepthToDisparity
  stereo/image_predictor.py:275-345   RectifiedImagePredictor
  stereo/image_predictor.py:347-398   IDepthImagePredictor
  stereo/image_predictor.py:525-601   IDepthmapProjector, ImagePredictor

Parity pin: tests/golden/image_predictor_small.npz holds the outputs of the reference's own classes (imported from
/root/reference by tests/golden/make_golden_image_predictor.py) on seeded inputs; tests/test_image_predictor_oracle.py
checks every function here against them.
"""
import torch
import torch.nn.functional as F


def _pixel_grid(rows, cols, dtype):
    y, x = torch.meshgrid(torch.arange(rows, dtype=dtype), torch.arange(cols, dtype=dtype), indexing="ij")
    return x.reshape(-1), y.reshape(-1)          # (rows*cols,)


def _mats(K, T_right_in_left, dtype):
    K = K.to(dtype)
    Kinv = torch.inverse(K)
    T_lr = torch.inverse(T_right_in_left.to(dtype))
    KRK = K[:, :3, :3] @ T_lr[:, :3, :3] @ Kinv[:, :3, :3]
    KT = K @ T_lr
    return K, Kinv, T_lr, KRK, KT


def _pix_inf(KRK, x, y):
    """Pixel of the point at infinite depth (image_predictor.py:163-167, 251-255)."""
    pz = KRK[:, 2, 0:1] * x + KRK[:, 2, 1:2] * y + KRK[:, 2, 2:3]
    px = (KRK[:, 0, 0:1] * x + KRK[:, 0, 1:2] * y + KRK[:, 0, 2:3]) / pz
    py = (KRK[:, 1, 0:1] * x + KRK[:, 1, 1:2] * y + KRK[:, 1, 2:3]) / pz
    return px, py, pz


def disparity_to_idepth(K, T_right_in_left, left_disparity, dtype=torch.float32):
    """image_predictor.py:120-209: least-squares inverse depth of a displacement of `disparity` pixels along the
    epipolar line (pointing from the point at depth 100 towards the point at infinity... far to near)."""
    n, rows, cols = K.shape[0], left_disparity.shape[-2], left_disparity.shape[-1]
    x, y = _pixel_grid(rows, cols, dtype)
    _, _, _, KRK, KT = _mats(K, T_right_in_left, dtype)
    Kt = KT[:, :3, 3]
    d = left_disparity.to(dtype).reshape(n, -1)
    ix, iy, pz = _pix_inf(KRK, x, y)
    fz = 100.0 * pz + Kt[:, 2:3]
    fx = (100.0 * (KRK[:, 0, 0:1] * x + KRK[:, 0, 1:2] * y + KRK[:, 0, 2:3]) + Kt[:, 0:1]) / fz
    fy = (100.0 * (KRK[:, 1, 0:1] * x + KRK[:, 1, 1:2] * y + KRK[:, 1, 2:3]) + Kt[:, 1:2]) / fz
    ex, ey = fx - ix, fy - iy
    nrm = torch.sqrt(ex * ex + ey * ey)
    bad = nrm < 1e-6
    ex = ex / (nrm + 1e-6)
    ey = ey / (nrm + 1e-6)
    A0 = Kt[:, 0:1] - Kt[:, 2:3] * (ix + d * ex)
    A1 = Kt[:, 1:2] - Kt[:, 2:3] * (iy + d * ey)
    b0, b1 = pz * d * ex, pz * d * ey
    idepth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1)
    idepth = (~bad).to(dtype) * idepth
    return idepth.reshape(n, 1, rows, cols)


def _project(K, T_right_in_left, left_idepthmap, dtype):
    """Back-project with depth = 1 / (idepth + 1e-6), move to the right frame, project (image_predictor.py:554-568)."""
    n, rows, cols = K.shape[0], left_idepthmap.shape[-2], left_idepthmap.shape[-1]
    x, y = _pixel_grid(rows, cols, dtype)
    K, Kinv, T_lr, KRK, KT = _mats(K, T_right_in_left, dtype)
    depth = 1.0 / (left_idepthmap.to(dtype).reshape(n, -1) + 1e-6)
    cam = [depth * (Kinv[:, i, 0:1] * x + Kinv[:, i, 1:2] * y + Kinv[:, i, 2:3]) for i in range(3)]
    right = [T_lr[:, i, 0:1] * cam[0] + T_lr[:, i, 1:2] * cam[1] + T_lr[:, i, 2:3] * cam[2] + T_lr[:, i, 3:4]
             for i in range(3)]
    w = [KT[:, i, 0:1] * cam[0] + KT[:, i, 1:2] * cam[1] + KT[:, i, 2:3] * cam[2] + KT[:, i, 3:4] for i in range(3)]
    u = (w[0] / (w[2] + 1e-7) + 0.5) * 2.0 / cols - 1.0
    v = (w[1] / (w[2] + 1e-7) + 0.5) * 2.0 / rows - 1.0
    return (x, y), K, KRK, right, u, v


def idepthmap_projector(K, T_right_in_left, left_idepthmap, dtype=torch.float32):
    """IDepthmapProjector.forward (:538-576) -> right_pixels (n, rows, cols, 2), right_idepths, mask (True = outside)."""
    n, rows, cols = K.shape[0], left_idepthmap.shape[-2], left_idepthmap.shape[-1]
    _, _, _, right, u, v = _project(K, T_right_in_left, left_idepthmap, dtype)
    pixels = torch.stack([u, v], dim=-1).reshape(n, rows, cols, 2)
    ridepth = (1.0 / (right[2] + 1e-6)).reshape(left_idepthmap.shape)
    mask = ((u.abs() > 1.0) | (v.abs() > 1.0)).reshape(n, 1, rows, cols)
    return pixels, ridepth, mask


def idepth_to_disparity(K, T_right_in_left, left_idepthmap, dtype=torch.float32):
    """IDepthToDisparity.forward (:228-273): distance between the projection and the pixel at infinite depth."""
    n, rows, cols = K.shape[0], left_idepthmap.shape[-2], left_idepthmap.shape[-1]
    (x, y), K, KRK, right, _, _ = _project(K, T_right_in_left, left_idepthmap, dtype)
    ix, iy, _ = _pix_inf(KRK, x, y)
    q = [K[:, i, 0:1] * right[0] + K[:, i, 1:2] * right[1] + K[:, i, 2:3] * right[2] for i in range(3)]
    dx, dy = q[0] / q[2] - ix, q[1] / q[2] - iy
    return torch.sqrt(dx * dx + dy * dy).reshape(n, 1, rows, cols)


def _sample(right_image, u, v, dtype):
    n, _, rows, cols = right_image.shape
    grid = torch.stack([u, v], dim=-1).reshape(n, rows, cols, 2)
    return F.grid_sample(right_image.to(dtype), grid, mode="bilinear", padding_mode="border", align_corners=False)


def idepth_image_predictor(K, T_right_in_left, left_idepthmap, right_image, dtype=torch.float32):
    """IDepthImagePredictor.forward (:357-398)."""
    n, rows, cols = K.shape[0], left_idepthmap.shape[-2], left_idepthmap.shape[-1]
    _, _, _, _, u, v = _project(K, T_right_in_left, left_idepthmap, dtype)
    mask = ((u.abs() > 1.0) | (v.abs() > 1.0)).reshape(n, 1, rows, cols)
    return _sample(right_image, u, v, dtype), mask


def image_predictor(K, T_right_in_left, left_disparity, right_image, dtype=torch.float32):
    """ImagePredictor.forward (:587-601): disparity -> idepth -> projection -> sample."""
    idepth = disparity_to_idepth(K, T_right_in_left, left_disparity, dtype)
    return idepth_image_predictor(K, T_right_in_left, idepth, right_image, dtype)


def rectified_image_predictor(K, T_right_in_left, left_disparity, right_image, dtype=torch.float32):
    """RectifiedImagePredictor.forward (:284-345): shift along x by sign(t_x) * disparity."""
    n, rows, cols = left_disparity.shape[0], left_disparity.shape[-2], left_disparity.shape[-1]
    x, y = _pixel_grid(rows, cols, dtype)
    sign = torch.sign(T_right_in_left[:, 0, 3]).to(dtype).reshape(n, 1)
    px = x - sign * left_disparity.to(dtype).reshape(n, -1)
    u = (px + 0.5) * 2.0 / cols - 1.0
    v = ((y + 0.5) * 2.0 / rows - 1.0).expand(n, -1)
    mask = ((u.abs() > 1.0) | (v.abs() > 1.0)).reshape(n, 1, rows, cols)
    return _sample(right_image, u, v, dtype), mask
