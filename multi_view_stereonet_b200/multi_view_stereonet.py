"""Drop-in for the reference's `multi_view_stereonet.multi_view_stereonet.MultiViewStereoNet`
(reference multi_view_stereonet/multi_view_stereonet.py:494-695).

Same constructor, same `forward(...)` signature and output dict, same state-dict
key names (so the reference's weights load unchanged); the work is done by the
hand-written sm_100a kernels behind the C ABI in include/b200mvs.h.  This module
only holds parameters, validates arguments and passes device pointers.
"""
import collections.abc
import ctypes
import warnings
from typing import Dict, List, Optional

import torch
import torch.nn as tnn

from . import _lib

ListTensor = List[torch.Tensor]


class LazyMaskPyramid(collections.abc.Sequence):
    """`left_idepthmap_mask_pyr` produced on demand (mask_mode = "lazy").

    The reference upsamples the (B, D, h, w) validity volume to every pyramid level inside forward
    (multi_view_stereonet.py:627-673) although nothing on the path reads it: 28 MB per image group at 512x640 / 64
    hypotheses, 168 MB at 1024x1280 / 128.  In lazy mode forward produces the level-4 volume only; indexing this
    sequence runs the MaskUpsampler chain (level 4 -> ... -> requested level, each level from the thresholded
    coarser one, exactly as the reference) the first time a level is asked for and caches it.  `packed(level)`
    returns the same volume as bits, (B, D, H, ceil(W / 8)) uint8 in numpy.packbits(axis=-1) layout."""

    def __init__(self, mask4, sizes):
        self._sizes = [tuple(s) for s in sizes]
        self._out_device = mask4.device
        self._dense = {4: mask4}        # uint8 0/1, (B, D, H_l, W_l), on the compute device once needed
        self._packed = {}

    def __len__(self):
        return len(self._sizes)

    def _compute_device(self):
        return self._out_device if self._out_device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())

    def _upsample(self, src, lvl, packed):
        lib = _lib.load()
        b, d, h, w = src.shape
        H, W = self._sizes[lvl]
        out = torch.empty((b, d, H, (W + 7) // 8 if packed else W), dtype=torch.uint8, device=src.device)
        with torch.cuda.device(src.device):
            stream = torch.cuda.current_stream(src.device).cuda_stream
            _lib.check(lib.b200mvs_upsample_mask(src.data_ptr(), b * d, h, w, H, W, int(packed), out.data_ptr(),
                                                 ctypes.c_void_p(stream)), "b200mvs_upsample_mask")
        return out

    def _dense_on_device(self, lvl):
        """uint8 volume of `lvl` on the compute device, building the chain from the finest level already there."""
        dev = self._compute_device()
        have = min(l for l in self._dense if l >= lvl)
        cur = self._dense[have]
        if cur.device != dev:
            cur = cur.to(dev)
        for l in range(have - 1, lvl - 1, -1):
            cur = self._upsample(cur, l, False)
            self._dense[l] = cur
        return cur

    def __getitem__(self, lvl):
        if isinstance(lvl, slice):
            return [self[i] for i in range(*lvl.indices(len(self)))]
        if lvl < 0:
            lvl += len(self)
        if not 0 <= lvl < len(self):
            raise IndexError(lvl)
        return self._dense_on_device(lvl).to(self._out_device).view(torch.bool)

    def packed(self, lvl):
        if lvl not in self._packed:
            if lvl in self._dense or lvl == len(self) - 1:
                src, same = self._dense_on_device(lvl), True
            else:
                src, same = self._dense_on_device(lvl + 1), False
            # from the coarser level the bits are written directly; the dense volume of `lvl` is never materialised
            self._packed[lvl] = self._upsample(src, lvl, True).to(self._out_device)
        return self._packed[lvl]


class _Conv(tnn.Module):
    """Parameter holder with the names of torch.nn.Conv2d / Conv3d."""

    def __init__(self, out_channels, in_channels, *kernel, bias=True):
        super().__init__()
        self.weight = tnn.Parameter(torch.empty(out_channels, in_channels, *kernel).normal_(0, 0.01))
        if bias:
            self.bias = tnn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)


class _Norm(tnn.Module):
    """Parameter holder with the names of torch.nn.GroupNorm(4, 32)."""

    def __init__(self, channels=32):
        super().__init__()
        self.weight = tnn.Parameter(torch.ones(channels))
        self.bias = tnn.Parameter(torch.zeros(channels))


class _ResBlock(tnn.Module):
    # utils/resnet.py:62-109 (conv1 + bn1)
    def __init__(self, bias):
        super().__init__()
        self.conv1 = _Conv(32, 32, 3, 3, bias=bias)
        self.bn1 = _Norm()


class FeatureNetwork(tnn.Module):
    # multi_view_stereonet.py:78-107
    def __init__(self, in_channels=3):
        super().__init__()
        self.channels = [in_channels, 32, 32, 32, 32]
        self.conv0 = _Conv(32, in_channels, 5, 5, bias=False)
        self.conv1 = _Conv(32, 32, 5, 5, bias=False)
        self.conv2 = _Conv(32, 32, 5, 5, bias=False)
        self.conv3 = _Conv(32, 32, 5, 5, bias=False)
        for i in range(6):
            setattr(self, f"res{i}", _ResBlock(bias=False))
        self.conv_final = _Conv(32, 32, 3, 3, bias=True)


class FeatureRefiner(tnn.Module):
    # multi_view_stereonet.py:398-422
    def __init__(self):
        super().__init__()
        self.conv0 = _Conv(32, 35, 3, 3)
        self.bn0 = _Norm()
        self.res0 = _ResBlock(bias=True)
        self.conv_final = _Conv(32, 32, 3, 3)


class IncrementalFastGeometryAwareFeatureNetwork(tnn.Module):
    # multi_view_stereonet.py:237-245; shares the left extractor (:507)
    def __init__(self, feature_extractor):
        super().__init__()
        self.feature_extractor = feature_extractor
        self.refiner = FeatureRefiner()


class CostVolumeFilter(tnn.Module):
    # multi_view_stereonet.py:302-339
    def __init__(self):
        super().__init__()
        for i in range(4):
            setattr(self, f"conv{i}", _Conv(32, 32, 3, 3, 3))
            setattr(self, f"bn{i}", _Norm())
        self.conv4 = _Conv(1, 32, 3, 3, 3)


class IDepthmapRefiner(tnn.Module):
    # multi_view_stereonet.py:442-466
    def __init__(self, image_channels):
        super().__init__()
        self.conv0 = _Conv(32, image_channels + 1, 3, 3)
        self.bn0 = _Norm()
        for i in range(6):
            setattr(self, f"res{i}", _ResBlock(bias=True))
        self.conv_final = _Conv(1, 32, 3, 3)


class MultiViewStereoNet(tnn.Module):
    """Multi-view stereo matching network; B200-native implementation of the
    reference module of the same name."""

    def __init__(self):
        super().__init__()
        self.min_idepth = 0.0
        self.num_levels = 5
        self.left_feature_extractor = FeatureNetwork(3)
        self.right_feature_extractor = IncrementalFastGeometryAwareFeatureNetwork(self.left_feature_extractor)
        self.volume_filter4 = CostVolumeFilter()
        self.refiner4 = IDepthmapRefiner(35)
        self.refiner3 = IDepthmapRefiner(35)
        self.refiner2 = IDepthmapRefiner(35)
        self.refiner1 = IDepthmapRefiner(35)
        self.refiner0 = IDepthmapRefiner(3)
        self._handle = None
        self._handle_key = None
        self._keep_stages = False
        self._host_out = None
        self._options = {}
        # "dense": the five mask volumes are computed inside forward, as the reference does (default, drop-in);
        # "lazy":  forward computes level 4 only and returns a LazyMaskPyramid;
        # "none":  the list holds level 4 only, None elsewhere.
        self.mask_mode = "dense"

    # -- native handle ---------------------------------------------------------------------
    # The native handle holds a packed copy of the weights; it must follow every change of the parameters, including
    # the in-place ones no module hook sees: `net.refiner0.load_state_dict(...)`, `torch.nn.init.*`, `p.copy_()`,
    # `optimizer.step()`.  Each of those bumps the parameter's autograd version counter, so the key -- the parameter
    # objects plus the sum of their versions -- is re-checked on EVERY forward over a cached parameter list (~20 us;
    # walking the module tree instead costs ~230 us).  Two edits are invisible to it and need `refresh_weights()`:
    # writes through `param.data` (PyTorch gives `.data` its own version counter) and re-assigning a submodule's
    # Parameter object.
    def _param_list(self):
        params = getattr(self, "_param_cache", None)
        if params is None:
            params = self._param_cache = tuple(self.parameters())
        return params

    def _weights_key(self, device_index):
        params = self._param_list()
        return (device_index, tuple(map(id, params)), sum(p._version for p in params))

    def refresh_weights(self):
        """Forces a re-upload of the weights on the next forward (only needed after `.data` writes or after
        replacing a submodule's Parameter object; tracked in-place updates are picked up automatically)."""
        self._param_cache = None
        self._handle_key = None

    def load_state_dict(self, *args, **kwargs):
        self._param_cache = None
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._param_cache = None
        return super()._apply(fn, *args, **kwargs)

    def register_parameter(self, name, param):
        self.__dict__["_param_cache"] = None
        return super().register_parameter(name, param)

    def _release(self):
        if self._handle is not None:
            _lib.load().b200mvs_destroy(self._handle)
            self._handle = None
            self._handle_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _native(self, device_index):
        key = self._weights_key(device_index)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        lib = _lib.load()
        self._release()
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in self.state_dict().items()}
        names = (ctypes.c_char_p * len(sd))(*[k.encode() for k in sd])
        data = _lib.ptr_array([v.data_ptr() for v in sd.values()])
        numels = (ctypes.c_int64 * len(sd))(*[v.numel() for v in sd.values()])
        handle = ctypes.c_void_p()
        _lib.check(lib.b200mvs_create(device_index, len(sd), names, data, numels, ctypes.byref(handle)),
                   "b200mvs_create")
        self._handle = handle
        self._handle_key = key
        if self._keep_stages:
            lib.b200mvs_set_debug(handle, 1)
        for name, value in self._options.items():
            _lib.check(lib.b200mvs_set_option(handle, name.encode(), value), "b200mvs_set_option")
        return handle

    def keep_stages(self, enable=True):
        """Test hook: preserve in-place-overwritten stage buffers for `get_stage`."""
        self._keep_stages = bool(enable)
        if self._handle is not None:
            _lib.load().b200mvs_set_debug(self._handle, 1 if enable else 0)

    def get_stage(self, name, dtype=torch.float32):
        """Test hook: a flat tensor holding stage `name` of the last forward."""
        lib = _lib.load()
        nbytes = ctypes.c_int64()
        _lib.check(lib.b200mvs_get_stage(self._handle, name.encode(), None, 0, ctypes.byref(nbytes), None),
                   "b200mvs_get_stage")
        dev = torch.device("cuda", self._handle_key[0])
        buf = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.b200mvs_get_stage(self._handle, name.encode(), buf.data_ptr(), nbytes.value,
                                         ctypes.byref(nbytes), stream), "b200mvs_get_stage")
        return buf.view(dtype)

    def set_option(self, name, value):
        """Library options (include/b200mvs.h: b200mvs_set_option), e.g. ("tensor_cores", 0)."""
        self._options[name] = int(value)
        if self._handle is not None:
            _lib.check(_lib.load().b200mvs_set_option(self._handle, name.encode(), int(value)), "b200mvs_set_option")

    def probe_select(self, kernel_class):
        """Brackets every launch of `kernel_class` with CUDA events (bench.py roofline leg)."""
        assert self._handle is not None, "run a forward first"
        _lib.check(_lib.load().b200mvs_probe_select(self._handle, kernel_class.encode()), "b200mvs_probe_select")

    def probe_read(self):
        """(summed device ms, launches) of the selected class since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(_lib.load().b200mvs_probe_read(self._handle, ctypes.byref(ms), ctypes.byref(n)),
                   "b200mvs_probe_read")
        return ms.value, n.value

    def last_stage_profile(self):
        """{stage: microseconds} of the last forward run with set_option("stage_profile", 1)."""
        if self._handle is None:
            return {}
        txt = _lib.load().b200mvs_last_stage_profile(self._handle).decode()
        return {k: float(v) for k, v in (kv.split("=") for kv in txt.split(";") if kv)}

    def last_launch_count(self):
        return int(_lib.load().b200mvs_last_launch_count(self._handle)) if self._handle is not None else 0

    # -- forward ---------------------------------------------------------------------------
    @staticmethod
    def _shape(left_image_pyr, T_right_in_lefts, num_idepth_samples, do_cost_volume_filter, do_refiners):
        s = _lib.Shape()
        s.batch, _, s.rows, s.cols = left_image_pyr[0].shape
        s.views = len(T_right_in_lefts)
        s.num_idepth_samples = int(num_idepth_samples)
        s.do_cost_volume_filter = int(bool(do_cost_volume_filter))
        for i in range(5):
            s.do_refiners[i] = int(bool(do_refiners[i]))
        return s

    def _check_args(self, left_image_pyr, K_pyr, T_right_in_lefts, right_image_pyrs, do_refiners):
        # the reference's asserts (multi_view_stereonet.py:548-549)
        assert len(K_pyr) == self.num_levels
        assert len(left_image_pyr) == self.num_levels
        assert len(do_refiners) == self.num_levels
        assert len(T_right_in_lefts) == len(right_image_pyrs) and len(T_right_in_lefts) >= 1
        b, c, h, w = left_image_pyr[0].shape
        assert c == 3
        for lvl in range(1, self.num_levels):
            h, w = (h + 1) // 2, (w + 1) // 2
            assert tuple(left_image_pyr[lvl].shape) == (b, 3, h, w), "pyramid level sizes must be ((h+1)//2, (w+1)//2)"
        for pyr in right_image_pyrs:
            assert len(pyr) == self.num_levels
            assert pyr[0].shape == left_image_pyr[0].shape and pyr[-1].shape == left_image_pyr[-1].shape
        for K in K_pyr:
            assert tuple(K.shape) == (b, 4, 4)
        for T in T_right_in_lefts:
            assert tuple(T.shape) == (b, 4, 4)

    def forward(self,
                left_image_pyr: ListTensor,
                K_pyr: ListTensor,
                T_right_in_lefts: ListTensor,
                right_image_pyrs: List[ListTensor],
                num_idepth_samples: int,
                do_cost_volume_filter: bool,
                do_refiners: List[bool]) -> Dict[str, List[Optional[torch.Tensor]]]:
        """Returns estimated left idepth maps (reference forward, multi_view_stereonet.py:538-695).

        CUDA tensors are processed in place on their device and stream.  CPU
        tensors are uploaded, processed on cuda:current and downloaded (the
        end-to-end entry `b200mvs_forward_host`)."""
        self._check_args(left_image_pyr, K_pyr, T_right_in_lefts, right_image_pyrs, do_refiners)
        if torch.is_grad_enabled():
            # the reference module is trainable; this one implements the inference path only (no backward kernels)
            if any(t.requires_grad for t in list(left_image_pyr) + [p[0] for p in right_image_pyrs]):
                raise RuntimeError("MultiViewStereoNet (B200) is inference-only: its outputs carry no grad_fn, so no "
                                   "gradient can reach inputs that require grad.  Call it under torch.no_grad().")
            if self.training and not getattr(self, "_warned_training", False):
                self._warned_training = True
                warnings.warn("MultiViewStereoNet (B200) is inference-only: called in training mode with autograd "
                              "enabled, but the outputs carry no grad_fn and the parameters will receive no "
                              "gradients.  Use .eval() and torch.no_grad() (test.py:191,196).", stacklevel=2)
        lib = _lib.load()
        dev = left_image_pyr[0].device
        shape = self._shape(left_image_pyr, T_right_in_lefts, num_idepth_samples, do_cost_volume_filter, do_refiners)
        on_host = dev.type != "cuda"
        if on_host:
            if not torch.cuda.is_available():
                raise RuntimeError("MultiViewStereoNet (B200): no CUDA device available and there is no CPU path")
            dev_index = torch.cuda.current_device()
        else:
            dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        handle = self._native(dev_index)

        prep = lambda t: t.detach().to(dtype=torch.float32).contiguous()
        left = [prep(t) for t in left_image_pyr]
        Ks = [prep(t) for t in K_pyr]
        Ts = [prep(t) for t in T_right_in_lefts]
        r0 = [prep(p[0]) for p in right_image_pyrs]
        r4 = [prep(p[-1]) for p in right_image_pyrs]
        for t in left + Ks + Ts + r0 + r4:
            assert t.device == dev, "all inputs must live on one device"

        b, d = shape.batch, shape.num_idepth_samples
        out_dev = torch.device("cpu") if on_host else dev
        sizes = [tuple(t.shape[-2:]) for t in left]
        want_raw, want_masks = True, True
        if on_host and self._host_out is not None:
            # bench.py's end-to-end leg: caller-provided (pinned) outputs, possibly a subset
            idepth = self._host_out["left_idepthmap_pyr"]
            raw = self._host_out.get("left_idepthmap_raw_pyr")
            mask = self._host_out.get("left_idepthmap_mask_pyr")
            want_raw, want_masks = raw is not None, mask is not None
        else:
            idepth = [torch.empty((b, 1) + s, dtype=torch.float32, device=out_dev) for s in sizes]
            raw = [torch.empty((b, 1) + s, dtype=torch.float32, device=out_dev) for s in sizes]
            assert self.mask_mode in ("dense", "lazy", "none"), self.mask_mode
            mask = [torch.empty((b, d) + s, dtype=torch.uint8, device=out_dev)
                    if (self.mask_mode == "dense" or lvl == 4) else None for lvl, s in enumerate(sizes)]
        null5 = _lib.ptr_array([None] * 5)
        args = [ctypes.byref(shape),
                _lib.ptr_array([t.data_ptr() for t in left]), _lib.ptr_array([t.data_ptr() for t in Ks]),
                _lib.ptr_array([t.data_ptr() for t in Ts]), _lib.ptr_array([t.data_ptr() for t in r0]),
                _lib.ptr_array([t.data_ptr() for t in r4]), _lib.ptr_array([t.data_ptr() for t in idepth]),
                _lib.ptr_array([t.data_ptr() for t in raw]) if want_raw else null5,
                _lib.ptr_array([t.data_ptr() if t is not None else None for t in mask]) if want_masks else null5]
        if on_host:
            h2d, d2h = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(lib.b200mvs_forward_host(handle, *args, ctypes.byref(h2d), ctypes.byref(d2h)),
                       "b200mvs_forward_host")
            self.last_h2d_bytes, self.last_d2h_bytes = h2d.value, d2h.value
        else:
            with torch.cuda.device(dev_index):
                stream = torch.cuda.current_stream(dev_index).cuda_stream
                _lib.check(lib.b200mvs_forward(handle, *args, ctypes.c_void_p(stream)), "b200mvs_forward")

        outputs: Dict[str, List[Optional[torch.Tensor]]] = {}
        outputs["left_idepthmap_pyr"] = idepth
        outputs["left_idepthmap_raw_pyr"] = raw if want_raw else [None] * 5
        if not want_masks:
            outputs["left_idepthmap_mask_pyr"] = [None] * 5
        elif mask[0] is None and self.mask_mode == "lazy":
            outputs["left_idepthmap_mask_pyr"] = LazyMaskPyramid(mask[4], sizes)
        else:
            outputs["left_idepthmap_mask_pyr"] = [None if m is None else (m.view(torch.bool) if m.dtype == torch.uint8 else m)
                                                  for m in mask]
        return outputs

    def set_host_outputs(self, out):
        """For CPU-tensor calls: write results into the given (pinned) tensors instead
        of allocating.  `out` maps "left_idepthmap_pyr" (required) and optionally
        "left_idepthmap_raw_pyr" / "left_idepthmap_mask_pyr" (uint8) to 5-lists; lists
        left out are computed on the device but not downloaded.  None restores the default."""
        self._host_out = out
