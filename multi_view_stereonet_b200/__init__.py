"""B200-native depth-inference hot path of MultiViewStereoNet."""
_LAYERS = ("HomographyImagePredictor", "ImagePredictor", "IDepthImagePredictor", "IDepthmapProjector",
           "DisparityToIDepth", "IDepthToDisparity", "RectifiedImagePredictor")
__all__ = ["MultiViewStereoNet"] + list(_LAYERS)


def __getattr__(name):
    # Lazy so that `import multi_view_stereonet_b200.synthetic` (used by the CPU
    # oracle tests) does not need the CUDA library to be built.
    if name == "MultiViewStereoNet":
        from .multi_view_stereonet import MultiViewStereoNet
        return MultiViewStereoNet
    if name in _LAYERS:
        from . import image_predictor
        return getattr(image_predictor, name)
    raise AttributeError(name)
