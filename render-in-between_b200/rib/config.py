"""Config handling that mirrors the reference's `get_config` (PGNR/utils/utils.py:77-79):
a YAML file becomes an attribute-style dict.  The reference's own `configs/HSM.yaml` is consumed
unchanged; `default_gen_cfg()` restates its `gen:` block (HSM.yaml:35-67) so that tests, the
benchmark and the smoke run work on a box where the reference tree is absent.
"""
import copy

import yaml


class AttrDict(dict):
    """dict with attribute access; nested dicts are converted on assignment.

    Like the EasyDict the reference uses, items are mirrored into ``__dict__`` so ``vars(cfg)``
    and ``getattr(cfg, key, default)`` behave the same way (PGNR/models/layers/conv.py:35,52).
    """

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        v = AttrDict._wrap(v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setattr__ = __setitem__

    def __deepcopy__(self, memo):
        return AttrDict(copy.deepcopy(dict(self), memo))


def get_config(path):
    with open(path, 'r') as f:
        return AttrDict(yaml.load(f, Loader=yaml.FullLoader))


_HSM_GEN = {
    'num_frames_G': 2,
    'input_image_nc': 3,
    'input_label_nc': 22,
    'num_filters': 16,
    'max_num_filters': 512,
    'num_layers': 6,
    'num_downsamples': 4,           # ignored by the reference (it reads num_downsamples_img, default 4)
    'kernel_size': 3,
    'activation_norm_type': 'spatially_adaptive',
    'activation_norm_params': {'activation_norm_type': 'instance', 'num_filters': 0, 'kernel_size': 1},
    'weight_norm_type': 'spectral',
    'do_checkpoint': True,
    'mask': {'generate_raw_output': False, 'num_filters': 32, 'max_num_filters': 512,
             'num_downsamples': 3, 'num_res_blocks': 4, 'kernel_size': 3,
             'activation_norm_type': 'instance', 'weight_norm_type': 'spectral'},
    'embed': {'use_embed': True, 'arch': 'encoder', 'num_filters': 64, 'max_num_filters': 512,
              'num_downsamples': 4, 'kernel_size': 3, 'weight_norm_type': 'spectral'},
}

# Rasteriser thresholds (HSM.yaml:184-187).
HSM_RASTER = {'gauss_sigma': 5, 'skeleton_thres': 0.001, 'foot_thres': 0.001}


def default_gen_cfg():
    """The `gen:` block of the reference's configs/HSM.yaml as an AttrDict."""
    return AttrDict(copy.deepcopy(_HSM_GEN))
