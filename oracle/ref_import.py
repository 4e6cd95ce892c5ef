"""TEST INFRASTRUCTURE ONLY — imports the unmodified reference renderer from /root/reference.

Only `tests/`, `oracle/make_golden.py` and `bench.py`'s CPU-baseline leg may use this module,
and only inside the build container: `/root/reference` does not exist on the GPU box, so
`available()` is False there and callers must fall back to the committed golden fixtures
(`tests/golden/`) and the restatements in `oracle/`.

The reference needs a handful of third-party modules that are not installed here
(SURVEY.md §8c); they are stubbed with just enough behaviour for the inference path:
  easydict   -> attribute-mirroring dict (utils/utils.py:7, conv.py:35,52 use vars(cfg))
  patoolib   -> import only (utils/utils.py:6)
  h5py       -> import only (evaluator.py:3, HSM_auto_dataset.py:5)
  piq        -> psnr/ssim names only (evaluator.py:7)
  albumentations -> Resize + Compose + KeypointParams (evaluator.py:18-26)
  imageio / matplotlib -> import only (utils/visualize.py)
"""
import os
import sys
import types

REF_ROOT = '/root/reference/Pose_Guided_Neural_Rendering'


def available():
    return os.path.isdir(REF_ROOT)


class _EasyDict(dict):
    """dict whose items are mirrored as instance attributes (so vars(cfg) works)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) else x for x in v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__


def _install_stubs():
    import numpy as np
    if not hasattr(np, 'float'):
        np.float = float  # HSM_auto_dataset.py:217 (removed in numpy >= 1.24)

    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        import easydict  # noqa: F401
    except ImportError:
        mod('easydict', EasyDict=_EasyDict)
    mod('patoolib')
    try:
        import h5py  # noqa: F401
    except ImportError:
        mod('h5py', File=None)
    try:
        import piq  # noqa: F401
    except ImportError:
        mod('piq', ssim=None, psnr=None)
    try:
        import imageio  # noqa: F401
    except ImportError:
        mod('imageio')
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = mod('matplotlib')
        mpl.use = lambda *a, **k: None
        mod('matplotlib.pyplot')
        mpl.pyplot = sys.modules['matplotlib.pyplot']
    try:
        import albumentations  # noqa: F401
    except ImportError:
        import cv2

        class Resize:
            def __init__(self, height, width, interpolation=cv2.INTER_CUBIC, always_apply=True):
                self.h, self.w, self.interp = height, width, interpolation

        class KeypointParams:
            def __init__(self, format='xy', remove_invisible=False):
                pass

        class Compose:
            def __init__(self, tlist, keypoint_params=None):
                self.tlist = tlist

            def __call__(self, image, keypoints):
                kp = [tuple(k) for k in keypoints]
                for t in self.tlist:
                    h0, w0 = image.shape[:2]
                    image = cv2.resize(image, (t.w, t.h), interpolation=t.interp)
                    # albumentations' Resize.apply_to_keypoint: scale_x = width / cols, then x * scale_x
                    kp = [(x * (t.w / w0), y * (t.h / h0)) for (x, y) in kp]
                return {'image': image, 'keypoints': kp}

        mod('albumentations', Resize=Resize, KeypointParams=KeypointParams, Compose=Compose,
            ShiftScaleRotate=None, Crop=None)


_loaded = {}


def load():
    """Returns a namespace with the reference's Generator, dataset class, Evaluator, get_config."""
    if _loaded:
        return _loaded['ns']
    if not available():
        raise RuntimeError('reference tree not present (%s)' % REF_ROOT)
    _install_stubs()
    # The reference uses top-level package names (models, utils, datasets); make sure ours do not
    # shadow them and that theirs do not leak into the product: import under a guarded sys.path.
    sys.path.insert(0, REF_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            from utils.utils import get_config, read_json_keypoint, tensor2images
            from models.generator import Generator
            from datasets.HSM_auto_dataset import HSMAutoDataset
            from models.evaluator import Evaluator
    finally:
        sys.path.remove(REF_ROOT)
    ns = types.SimpleNamespace(
        get_config=get_config, read_json_keypoint=read_json_keypoint, tensor2images=tensor2images,
        Generator=Generator, HSMAutoDataset=HSMAutoDataset, Evaluator=Evaluator,
        config_path=os.path.join(REF_ROOT, 'configs', 'HSM.yaml'))
    _loaded['ns'] = ns
    return ns


def make_dataset(height, width):
    """A reference HSMAutoDataset in test phase at the given model size (no h5 needed)."""
    ns = load()
    cfg = ns.get_config(ns.config_path)
    cfg.model_height, cfg.model_width = height, width
    cfg.load_height, cfg.load_width = height, width
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        ds = ns.HSMAutoDataset(cfg, cfg.h5_file, phase='test')
    return ds


def make_generator():
    ns = load()
    cfg = ns.get_config(ns.config_path)
    return ns.Generator(cfg.gen)
