"""rib — B200-native hot path of Render-In-Between's pose-guided renderer.

Public surface (mirrors the reference's names for this path):
    Generator(gen_cfg).forward(label, label_prev, img_fake, img_prev)   models/generator.py
    rasterize / warp / composite                                        evaluator.py, HSM_auto_dataset.py
    ClipRenderer                                                        evaluator.py:238-266 (AR loop)
    evaluate_from_folder                                                evaluator.py:165-269 (folder-driven entry)
    resize_cubic_u8 / frames_from_u8                                    evaluator.py:18-26, HSM_auto_dataset.py:73-75
    get_config / default_gen_cfg                                        utils/utils.py:77-79
    MotionTransformer / MotionInference / PositionEmbeddingSine1D       Human_Motion_Modelling/models/transformer.py,
                                                                        position_encoding.py, inference.py (joints upstream)
"""
from .config import AttrDict, default_gen_cfg, get_config  # noqa: F401
from .arch import Arch  # noqa: F401


def __getattr__(name):
    # CUDA-backed symbols load the shared library on first use, so that pure-host helpers
    # (config, arch, synth) stay importable without it; using them without the library raises.
    if name in ('Generator',):
        from .generator import Generator
        return Generator
    if name in ('rasterize', 'warp', 'composite', 'gaussian_taps', 'frames_from_u8', 'resize_cubic_u8'):
        from . import ops
        return getattr(ops, name)
    if name in ('evaluate_from_folder',):
        from .folder import evaluate_from_folder
        return evaluate_from_folder
    if name in ('ClipRenderer',):
        from .clip import ClipRenderer
        return ClipRenderer
    if name in ('MotionTransformer', 'MotionInference', 'PositionEmbeddingSine1D'):
        from . import motion
        return getattr(motion, name)
    if name == 'lib':
        from ._lib import lib
        return lib
    raise AttributeError(name)
