"""Drop-in replacement of the upstream motion model for inference:

    models.transformer.Transformer           Human_Motion_Modelling/models/transformer.py:16-132
    models.position_encoding.PositionEmbeddingSine_1D   Human_Motion_Modelling/models/position_encoding.py:9-56
    Model_inference.inference                Human_Motion_Modelling/inference.py:21-43

Same constructor arguments as `Transformer(...)` / `build_transformer(args)`, same state-dict keys and shapes (188
tensors for configs/config.yaml), same positional `forward(src, src_mask, src_pos, tgt, tgt_mask, tgt_pos, rate) ->
(joints, reco)`, so a checkpoint's `transformer` entry loads unchanged (trainer.resume).  The math runs in the fp32 kernels
of librib_b200.so (csrc/motion.cu) on the caller's stream: a renderer process can interpolate the joints on the GPU and feed
them to `rib.ClipRenderer` without the JSON round trip between the two stages.  Only the options the repository ships are
supported (pre_norm, leaky_relu, two_stage, no intermediate outputs); there is no PyTorch fallback.
"""
import ctypes as C
import math

import torch
from torch import nn

from ._lib import MotionConfig, Tensor, check, lib


class _Node(nn.Module):
    """Bare container used to reproduce the reference's module tree (and therefore its key names)."""


def state_spec(input_nc, d_model, dim_feedforward, num_encoder_layers, num_decoder_layers):
    """(key, shape) of every parameter of the reference Transformer, in its state-dict order."""
    d, ff, j = d_model, dim_feedforward, input_nc
    spec = [('input_embed.weight', (d, j)), ('input_embed.bias', (d,))]

    def attn(p):
        return [(p + '.in_proj_weight', (3 * d, d)), (p + '.in_proj_bias', (3 * d,)),
                (p + '.out_proj.weight', (d, d)), (p + '.out_proj.bias', (d,))]

    def rest(p, norms):
        s = [(p + '.linear1.weight', (ff, d)), (p + '.linear1.bias', (ff,)), (p + '.linear2.weight', (d, ff)), (p + '.linear2.bias', (d,))]
        for i in range(1, norms + 1):
            s += [(p + '.norm%d.weight' % i, (d,)), (p + '.norm%d.bias' % i, (d,))]
        return s

    for i in range(num_encoder_layers):
        p = 'encoder.layers.%d' % i
        spec += attn(p + '.self_attn') + rest(p, 2)
    spec += [('encoder.norm.weight', (d,)), ('encoder.norm.bias', (d,))]
    for i in range(num_decoder_layers):
        p = 'decoder.layers.%d' % i
        spec += attn(p + '.self_attn') + attn(p + '.multihead_attn') + rest(p, 3)
    spec += [('decoder.norm.weight', (d,)), ('decoder.norm.bias', (d,)), ('joints_embed.weight', (j, d)), ('joints_embed.bias', (j,))]
    return spec


class PositionEmbeddingSine1D(nn.Module):
    """position_encoding.py:9-56 (normalize=True, the 'v2' setting): mask [N, L] -> [L, N, 2 * num_pos_feats]; only the
    shape of the mask is read.  Same float32 operations as the reference, on the mask's device."""

    def __init__(self, num_pos_feats=64, temperature=10000, scale=None):
        super().__init__()
        self.num_pos_feats, self.temperature = num_pos_feats, temperature
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, mask):
        n, length = mask.shape
        position = torch.arange(0, length, dtype=torch.float32).unsqueeze(0).repeat(n, 1)
        position = position / (position[:, -1:] + 1e-6) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_pos_feats)
        pe = torch.zeros(n, length, self.num_pos_feats * 2)                    # (the reference builds it on the host too)
        pe[:, :, 0::2] = torch.sin(position[:, :, None] / dim_t)
        pe[:, :, 1::2] = torch.cos(position[:, :, None] / dim_t)
        return pe.permute(1, 0, 2).contiguous().to(mask.device)


class MotionTransformer(nn.Module):
    def __init__(self, input_nc, d_model=128, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=256,
                 dropout=0.1, activation='leaky_relu', normalize_before=True, return_intermediate_dec=False, two_stage=True):
        super().__init__()
        if activation != 'leaky_relu' or not normalize_before or return_intermediate_dec or not two_stage:
            raise ValueError('rib.MotionTransformer supports the shipped options only: activation=leaky_relu, pre_norm, '
                             'two_stage, no intermediate decoder outputs (configs/config.yaml:77-94)')
        if d_model != nhead * 16 or d_model % 64 or max(input_nc, d_model, dim_feedforward) > 256:
            raise ValueError('rib.MotionTransformer: head dimension must be 16, hidden_dim a multiple of 64, no dimension above 256')
        self.joints_dim, self.d_model, self.nhead = input_nc, d_model, nhead
        self.dim_feedforward = dim_feedforward
        self.num_encoder_layers, self.num_decoder_layers = num_encoder_layers, num_decoder_layers
        for key, shape in state_spec(input_nc, d_model, dim_feedforward, num_encoder_layers, num_decoder_layers):
            parts = key.split('.')
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            node.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape, dtype=torch.float32), requires_grad=False))
        self._handle = None
        self._ws = {}

    # -- device copy of the parameters ----------------------------------------------------------
    def _invalidate(self):
        if getattr(self, '_handle', None):
            lib.rib_motion_destroy(self._handle)
        self._handle = None
        self._ws = {}

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._invalidate()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def _ensure(self, device):
        if self._handle is not None:
            return
        sd = {k: v.detach().contiguous() for k, v in self.state_dict().items()}
        for k, v in sd.items():
            if not (v.is_cuda and v.dtype == torch.float32):
                raise RuntimeError('rib.MotionTransformer: parameter %s must be a CUDA float32 tensor (call .to("cuda") '
                                   'first); there is no CPU path' % k)
        cfg = MotionConfig(self.joints_dim, self.d_model, self.nhead, self.dim_feedforward, self.num_encoder_layers,
                           self.num_decoder_layers)
        arr = (Tensor * len(sd))()
        names = []
        for i, (k, v) in enumerate(sd.items()):
            names.append(k.encode())
            arr[i] = Tensor(names[-1], v.data_ptr(), v.numel())
        handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.rib_motion_create(C.byref(cfg), arr, len(sd), C.c_void_p(torch.cuda.current_stream().cuda_stream),
                                        C.byref(handle)), 'rib_motion_create')
        self._handle = handle

    # -- Transformer.forward ---------------------------------------------------------------------
    def forward(self, src, src_mask, src_pos, tgt, tgt_mask, tgt_pos, rate):
        """src [N, C, L] float32; src_mask / tgt_mask bool [N, L] (True = hidden key) or None; src_pos / tgt_pos
        [L, N, d_model]; tgt is accepted for signature compatibility and not read (two_stage).  Returns
        (joints [L, N, C], reco [L, N, C])."""
        if not (src.is_cuda and src.dtype == torch.float32 and src.dim() == 3 and src.shape[1] == self.joints_dim):
            raise ValueError('src must be a CUDA float32 tensor [N, %d, L]' % self.joints_dim)
        n, c, length = src.shape
        dev = src.device
        self._ensure(dev)
        src = src.contiguous()

        def pos_arg(t, name):
            if not (t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (length, n, self.d_model)):
                raise ValueError('%s must be a CUDA float32 tensor [L, N, %d]' % (name, self.d_model))
            return t.contiguous()

        def mask_arg(t, name):
            if t is None:
                return None
            if not (t.is_cuda and tuple(t.shape) == (n, length)):
                raise ValueError('%s must be a CUDA tensor [N, L]' % name)
            return t.to(torch.uint8).contiguous()

        sp, tp = pos_arg(src_pos, 'src_pos'), pos_arg(tgt_pos, 'tgt_pos')
        sm, tm = mask_arg(src_mask, 'src_mask'), mask_arg(tgt_mask, 'tgt_mask')
        key = (n, length, dev.index)
        ws = self._ws.get(key)
        if ws is None:
            ws = torch.empty(lib.rib_motion_workspace_bytes(self._handle, n, length), dtype=torch.uint8, device=dev)
            self._ws = {key: ws}
        joints = torch.empty(length, n, c, dtype=torch.float32, device=dev)
        reco = torch.empty(length, n, c, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.rib_motion_forward(self._handle, n, length, src.data_ptr(), sm.data_ptr() if sm is not None else None,
                                         sp.data_ptr(), tm.data_ptr() if tm is not None else None, tp.data_ptr(), int(rate),
                                         joints.data_ptr(), reco.data_ptr(), ws.data_ptr(), ws.numel(),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rib_motion_forward')
        return joints, reco


class MotionInference(nn.Module):
    """Model_inference (inference.py:10-43): `inference(data [C, L], interp, encoder_mask [L], decoder_mask [L], rate)` ->
    pred [1, C, L]."""

    def __init__(self, enc, transformer):
        super().__init__()
        self.pos_encode, self.transformer = enc, transformer

    def inference(self, data, interp, encoder_mask, decoder_mask, rate):
        dev = next(self.transformer.parameters()).device
        src = data.unsqueeze(0).to(dev)
        sm, tm = encoder_mask.unsqueeze(0).to(dev), decoder_mask.unsqueeze(0).to(dev)
        pos_src, pos_tar = self.pos_encode(sm), self.pos_encode(tm)
        pred, _ = self.transformer.forward(src, sm, pos_src, None, tm, pos_tar, rate)
        return pred.permute(1, 2, 0)
