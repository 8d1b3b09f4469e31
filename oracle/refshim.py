"""TEST / BENCH INFRASTRUCTURE -- not product code.  Import the UNMODIFIED reference (alanqrwang/keymorph)
from wherever oracle/build_ref.py installed it (oracle/_ref/, git-ignored, travels to the GPU box) or, in
the build container, from /root/reference.

The reference's package __init__ eagerly imports nibabel / skimage / h5py / torchio / matplotlib
(keymorph/__init__.py:4-11); none of them is installed here and none is used by the registration hot
path, so those five top-level names resolve to MagicMock modules.  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
MOCKED = {"nibabel", "skimage", "h5py", "torchio", "matplotlib", "wandb", "torchvision", "SimpleITK", "itk"}


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = set(names)

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


def _missing(names):
    out = set()
    for n in names:
        try:
            if importlib.util.find_spec(n) is None:
                out.add(n)
        except (ImportError, ValueError):
            out.add(n)
    return out


def reference_root():
    """Directory that holds the reference's `keymorph` package, or None."""
    env = os.environ.get("KEYMORPH_REFERENCE")
    for cand in (env, REF_DIR, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "keymorph", "model.py")):
            return cand
    return None


def available():
    return reference_root() is not None


_installed = False


def import_reference():
    """-> the reference's `keymorph` package (imported once; absent third-party modules are mocked)."""
    global _installed
    root = reference_root()
    if root is None:
        raise ImportError("reference not found: run `python oracle/build_ref.py` in the build container "
                          "(installs /root/reference into oracle/_ref/)")
    if not _installed:
        import importlib.util  # noqa: F401
        sys.meta_path.insert(0, _MockFinder(_missing(MOCKED)))
        sys.path.insert(0, root)
        _installed = True
    import keymorph
    assert os.path.dirname(os.path.dirname(os.path.abspath(keymorph.__file__))) == os.path.abspath(root), \
        f"`keymorph` resolved to {keymorph.__file__}, not the reference at {root}"
    return keymorph


def import_reference_scripts():
    """-> the reference's `scripts` package (register.py / pairwise_register_eval.py / script_utils.py),
    copied beside the installed package by oracle/build_ref.py (setup.py excludes it from the wheel)."""
    import_reference()
    import scripts  # noqa: F401
    return scripts


def build_reference_model(num_keypoints, backbone="truncatedunet", seed=23, device="cpu", **kw):
    """KeyMorph(DataParallel(backbone)) exactly as scripts/register.py:212-301 `get_model` builds it
    (TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order='gcr', num_groups=8,
    num_levels=4, is_segmentation=False, conv_padding=1)), seeded with the scripts' default seed."""
    import torch
    km = import_reference()
    from keymorph.model import KeyMorph
    from keymorph.net import ConvNet
    from keymorph.unet3d.model import TruncatedUNet3D, UNet3D
    torch.manual_seed(seed)
    if backbone == "conv":
        net = ConvNet(3, 1, num_keypoints, "instance")
    else:
        cls = TruncatedUNet3D if backbone == "truncatedunet" else UNet3D
        args = (1, num_keypoints) + ((1,) if backbone == "truncatedunet" else ())
        net = cls(*args, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                  is_segmentation=False, conv_padding=1)
    del km
    dp = torch.nn.DataParallel(net)
    if torch.device(device).type == "cpu":
        # on a box WITH GPUs DataParallel refuses CPU parameters (it scatters to device_ids); an empty
        # device list is what its constructor sets on a CPU-only box: forward() = module.forward()
        dp.device_ids = []
    else:
        dp.device_ids = [torch.device(device).index or 0]
        dp.output_device = dp.device_ids[0]
    return KeyMorph(dp, num_keypoints, 3, **kw).eval().to(device)
