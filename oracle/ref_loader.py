"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (read-only mount at /root/reference).

In the build container the reference is the read-only mount at /root/reference; on the GPU box (no mount) it is
the unmodified copy that ``baseline/stage_reference.py`` staged under the git-ignored ``baseline/_ref/`` (it travels
with the gpurun snapshot).  Used by ``oracle/make_golden*.py`` to (a) pin the oracles against the reference modules and
(b) generate the fixtures under ``tests/golden/``; by the reference-wrapper GPU tests (the reference's OWN
``PredictorBasedGenerator`` / ``FlowGenerator`` driving the drop-in predictor); and by ``bench.py --impl reference``.
Nothing in the product package imports this.

The reference needs three pip packages that are absent here: ``timm`` (five
symbols: vmae.py:12-15, VideoMAE/utils.py:6-9), ``kornia`` and ``matplotlib`` (import-only on this
path).  We register in-process stubs in ``sys.modules`` before importing ``cwm``.
"""
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# the staged copy first: the GPU tests and bench.py then import the same files on both boxes and never touch the mount
_CANDIDATES = (os.path.join(os.path.dirname(_HERE), "baseline", "_ref"), "/root/reference")


def reference_root():
    """The first location that holds the reference's ``cwm`` package (None if neither does)."""
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "cwm", "models", "VideoMAE", "vmae.py")):
            return root
    return None


REFERENCE_ROOT = reference_root() or _CANDIDATES[1]


def available():
    return reference_root() is not None


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


def install_stubs():
    if "timm" not in sys.modules:
        timm = _stub("timm")
        models = _stub("timm.models")
        registry = _stub("timm.models.registry")
        layers = _stub("timm.models.layers")
        data = _stub("timm.data")
        constants = _stub("timm.data.constants")
        registry.register_model = lambda f: f

        def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
            return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        def drop_path(x, drop_prob=0., training=False):
            assert (not training) or (not drop_prob)
            return x

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        layers.trunc_normal_ = trunc_normal_
        layers.drop_path = drop_path
        layers.to_2tuple = to_2tuple
        constants.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
        constants.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
        timm.models, timm.data = models, data
        models.registry, models.layers = registry, layers
        data.constants = constants
    for name in ("kornia",):
        if name not in sys.modules:
            _stub(name)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib")
            plt = _stub("matplotlib.pyplot")
            mpl.pyplot = plt
            for sub in ("patches", "widgets", "cm", "colors"):
                setattr(mpl, sub, _stub("matplotlib." + sub))


def import_reference():
    """Returns the reference's ``cwm.models.VideoMAE.vmae`` and ``cwm.models.prediction`` modules."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import cwm.models.VideoMAE.vmae as ref_vmae
    import cwm.models.prediction as ref_prediction
    return ref_vmae, ref_prediction
