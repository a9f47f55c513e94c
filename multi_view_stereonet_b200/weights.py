"""Weight sources for the depth-inference hot path.

The reference ships its weights as TorchScript archives written by torch 1.5
(`pretrained/*/checkpoints/*/stereo_network.pt`, loaded at reference
`test.py:311` with `torch.jit.load`).  Under torch >= 2 that call fails on a
removed operator schema, so the 226 state-dict tensors are recovered by reading
the archive as a plain zip + pickle.  The state-dict key
names are the reference's (`multi_view_stereonet/multi_view_stereonet.py:506-532`)
so the same dict loads into the reference model and into ours.
"""
import collections
import io
import pickle
import zipfile

import numpy as np
import torch


class _Stub:
    """Stands in for any `__torch__.*` scripted class inside the archive."""

    def __init__(self, *args, **kwargs):
        self.state = {}

    def __setstate__(self, state):
        self.state = state


def _walk(obj, prefix, out):
    if isinstance(obj, torch.Tensor):
        out[prefix] = obj
    elif isinstance(obj, _Stub):
        _walk(obj.state, prefix, out)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            if isinstance(k, str):
                _walk(v, f"{prefix}.{k}" if prefix else k, out)
    elif isinstance(obj, (tuple, list)):
        for v in obj:
            if isinstance(v, (dict, _Stub)):
                _walk(v, prefix, out)


_ALLOWED_GLOBALS = {("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"),
                    ("torch._utils", "_rebuild_parameter"), ("torch", "device"), ("torch", "Size"),
                    ("torch.jit._pickle", "build_intlist"), ("torch.jit._pickle", "build_boollist"),
                    ("torch.jit._pickle", "build_doublelist"), ("torch.jit._pickle", "build_tensorlist")}


def load_torchscript_archive_weights(path):
    """Returns {state_dict key: float32 tensor} from a reference `.pt` archive."""
    zf = zipfile.ZipFile(path)
    root = zf.namelist()[0].split("/")[0]

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("__torch__"):
                return type(name, (_Stub,), {})
            if module == "collections" and name == "OrderedDict":
                return collections.OrderedDict
            # The archive needs nothing but tensor rebuild helpers and storage types; any other global (os.system,
            # builtins.eval, ...) in a crafted data.pkl would run code here, so it is refused.
            if (module, name) in _ALLOWED_GLOBALS or (module == "torch" and name.endswith("Storage")):
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f"weight archive references disallowed global {module}.{name}")

        def persistent_load(self, pid):
            # ('storage', storage_type, key, location, numel)
            _, storage_type, key, _, numel = pid
            raw = bytearray(zf.read(f"{root}/data/{key}"))
            dtype = {"FloatStorage": torch.float32, "LongStorage": torch.int64,
                     "IntStorage": torch.int32, "DoubleStorage": torch.float64,
                     "BoolStorage": torch.bool}[storage_type.__name__]
            t = torch.frombuffer(raw, dtype=dtype) if len(raw) else torch.empty(0, dtype=dtype)
            return torch.storage.TypedStorage(
                wrap_storage=t.untyped_storage(), dtype=dtype, _internal=True)

    obj = Unpickler(io.BytesIO(zf.read(f"{root}/data.pkl"))).load()
    found = {}
    _walk(obj, "", found)
    # Keep parameters only; drop anything that is not a float tensor
    # (e.g. scripted constants), and de-alias shared storages.
    state = collections.OrderedDict()
    for k, v in found.items():
        if v.dtype == torch.float32 and not k.endswith("num_batches_tracked"):
            state[k] = v.clone().contiguous()
    return state


def save_state_npz(state, path):
    np.savez(path, **{k: v.detach().cpu().numpy() for k, v in state.items()})


def load_state_npz(path):
    """Loads a state dict stored by `save_state_npz` (a torch-free fixture)."""
    with np.load(path) as z:
        return collections.OrderedDict((k, torch.from_numpy(z[k].copy())) for k in z.files)


def seeded_random_state(seed=0):
    """A deterministic non-trivial weight set with the reference's key names and
    shapes, for runs where no pretrained fixture is available (e.g. bench.py on a
    box without tests/golden).  Weight scale is chosen so that costs are not ~0
    (the reference's N(0, 0.01) init makes soft-argmin uniform)."""
    g = torch.Generator().manual_seed(seed)
    state = collections.OrderedDict()

    def conv(name, o, i, *k, bias=True):
        fan_in = i * int(np.prod(k))
        state[name + ".weight"] = torch.randn(o, i, *k, generator=g) * (1.0 / fan_in) ** 0.5
        if bias:
            state[name + ".bias"] = torch.randn(o, generator=g) * 0.05

    def gn(name):
        state[name + ".weight"] = 1.0 + 0.1 * torch.randn(32, generator=g)
        state[name + ".bias"] = 0.05 * torch.randn(32, generator=g)

    fe = "left_feature_extractor"
    conv(f"{fe}.conv0", 32, 3, 5, 5, bias=False)
    for i in (1, 2, 3):
        conv(f"{fe}.conv{i}", 32, 32, 5, 5, bias=False)
    for i in range(6):
        conv(f"{fe}.res{i}.conv1", 32, 32, 3, 3, bias=False)
        gn(f"{fe}.res{i}.bn1")
    conv(f"{fe}.conv_final", 32, 32, 3, 3)
    rf = "right_feature_extractor.refiner"
    conv(f"{rf}.conv0", 32, 35, 3, 3)
    gn(f"{rf}.bn0")
    conv(f"{rf}.res0.conv1", 32, 32, 3, 3)
    gn(f"{rf}.res0.bn1")
    conv(f"{rf}.conv_final", 32, 32, 3, 3)
    for i in range(4):
        conv(f"volume_filter4.conv{i}", 32, 32, 3, 3, 3)
        gn(f"volume_filter4.bn{i}")
    conv("volume_filter4.conv4", 1, 32, 3, 3, 3)
    for lvl in range(5):
        r = f"refiner{lvl}"
        conv(f"{r}.conv0", 32, 36 if lvl > 0 else 4, 3, 3)
        gn(f"{r}.bn0")
        for i in range(6):
            conv(f"{r}.res{i}.conv1", 32, 32, 3, 3)
            gn(f"{r}.res{i}.bn1")
        conv(f"{r}.conv_final", 1, 32, 3, 3)
    # The reference registers the shared FeatureNetwork twice
    # (multi_view_stereonet.py:506-507), so its state dict carries both names.
    for k in [k for k in state if k.startswith(fe + ".")]:
        state["right_feature_extractor.feature_extractor." + k[len(fe) + 1:]] = state[k]
    return state
