"""CPU oracle (TEST INFRASTRUCTURE, never on the product path) of the upstream motion model: the DETR-style
encoder / decoder Transformer of Human_Motion_Modelling that interpolates 2-D joint sequences.

Restates, as explicit torch fp32 arithmetic on the CPU,
    Transformer.forward / encode / decode / interpolate_embedding   Human_Motion_Modelling/models/transformer.py:57-132
    TransformerEncoderLayer.forward_pre                              Human_Motion_Modelling/models/transformer.py:243-254
    TransformerDecoderLayer.forward_pre                              Human_Motion_Modelling/models/transformer.py:316-337
    TransformerEncoder / TransformerDecoder (final LayerNorm)        Human_Motion_Modelling/models/transformer.py:134-196
    PositionEmbeddingSine_1D.forward                                 Human_Motion_Modelling/models/position_encoding.py:26-56
    Model_inference.inference                                        Human_Motion_Modelling/inference.py:21-43
for the shipped configuration (configs/config.yaml:77-94: 38 joint coordinates, d_model 128, 8 heads, feed-forward 256,
6 + 6 layers, leaky_relu, pre_norm, two_stage, sine position encoding), eval mode (dropout = identity).
torch.nn.MultiheadAttention is written out: packed in-projection split in three, q scaled by sqrt(1 / head_dim) BEFORE
the dot products, boolean masks as -inf, softmax over keys, out-projection.

Pinned against the unmodified reference modules in tests/test_oracle_vs_reference.py (float32: max-abs 1e-5) and against
tests/golden/motion_case.npz (made by oracle/make_golden_motion.py from the reference) everywhere.
"""
import math

import torch
import torch.nn.functional as F

CFG = dict(input_joints=38, hidden_dim=128, nheads=8, dim_feedforward=256, enc_layers=6, dec_layers=6)
LRELU_SLOPE = 0.01   # F.leaky_relu default (transformer.py:370-371)
LN_EPS = 1e-5        # nn.LayerNorm default


def state_spec(cfg=CFG):
    """(key, shape) of every parameter of the reference Transformer, in state-dict order."""
    d, ff, j = cfg['hidden_dim'], cfg['dim_feedforward'], cfg['input_joints']
    spec = [('input_embed.weight', (d, j)), ('input_embed.bias', (d,))]

    def attn(p):
        return [(p + '.in_proj_weight', (3 * d, d)), (p + '.in_proj_bias', (3 * d,)),
                (p + '.out_proj.weight', (d, d)), (p + '.out_proj.bias', (d,))]

    def ffn_norms(p, n_norm):
        s = [(p + '.linear1.weight', (ff, d)), (p + '.linear1.bias', (ff,)),
             (p + '.linear2.weight', (d, ff)), (p + '.linear2.bias', (d,))]
        for i in range(1, n_norm + 1):
            s += [(p + '.norm%d.weight' % i, (d,)), (p + '.norm%d.bias' % i, (d,))]
        return s

    for i in range(cfg['enc_layers']):
        p = 'encoder.layers.%d' % i
        spec += attn(p + '.self_attn') + ffn_norms(p, 2)
    spec += [('encoder.norm.weight', (d,)), ('encoder.norm.bias', (d,))]
    for i in range(cfg['dec_layers']):
        p = 'decoder.layers.%d' % i
        spec += attn(p + '.self_attn') + attn(p + '.multihead_attn') + ffn_norms(p, 3)
    spec += [('decoder.norm.weight', (d,)), ('decoder.norm.bias', (d,)),
             ('joints_embed.weight', (j, d)), ('joints_embed.bias', (j,))]
    return spec


def synth_state_dict(seed=0, cfg=CFG):
    """Deterministic random-init weights of realistic scale (xavier-like matrices, non-trivial biases and LayerNorm affines)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in state_spec(cfg):
        if k.endswith('weight') and len(shape) == 2:
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            sd[k] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif '.norm' in k and k.endswith('weight'):
            sd[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(shape, generator=g)
    return sd


def position_encoding(n, length, hidden_dim=CFG['hidden_dim']):
    """PositionEmbeddingSine_1D(hidden_dim // 2, normalize=True)(mask[n, length]) -> [L, N, hidden_dim]."""
    npf = hidden_dim // 2
    position = torch.arange(0, length, dtype=torch.float32).unsqueeze(0).repeat(n, 1)
    position = position / (position[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = 10000 ** (2 * (dim_t // 2) / npf)
    pe = torch.zeros(n, length, npf * 2)
    pe[:, :, 0::2] = torch.sin(position[:, :, None] / dim_t)
    pe[:, :, 1::2] = torch.cos(position[:, :, None] / dim_t)
    return pe.permute(1, 0, 2)


def _mha(sd, p, query, key, value, nhead, attn_mask=None, key_padding_mask=None):
    """nn.MultiheadAttention.forward(query, key, value, attn_mask, key_padding_mask)[0]; tensors are [L, N, E]."""
    lq, n, e = query.shape
    lk = key.shape[0]
    dh = e // nhead
    w, b = sd[p + '.in_proj_weight'], sd[p + '.in_proj_bias']
    q = F.linear(query, w[:e], b[:e])
    k = F.linear(key, w[e:2 * e], b[e:2 * e])
    v = F.linear(value, w[2 * e:], b[2 * e:])
    q = q.reshape(lq, n * nhead, dh).transpose(0, 1) * math.sqrt(1.0 / dh)      # [N*h, Lq, dh]
    k = k.reshape(lk, n * nhead, dh).transpose(0, 1)
    v = v.reshape(lk, n * nhead, dh).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2))                                    # [N*h, Lq, Lk]
    neg = torch.zeros(n, 1, lq, lk)
    if attn_mask is not None:                                                   # bool [Lq, Lk], True = not allowed
        neg = neg.masked_fill(attn_mask[None, None], float('-inf'))
    if key_padding_mask is not None:                                            # bool [N, Lk], True = ignored key
        neg = neg.masked_fill(key_padding_mask[:, None, None, :], float('-inf'))
    scores = scores + neg.expand(n, nhead, lq, lk).reshape(n * nhead, lq, lk)
    attn = torch.softmax(scores, dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(lq, n, e)
    return F.linear(out, sd[p + '.out_proj.weight'], sd[p + '.out_proj.bias'])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], LN_EPS)


def _ffn(sd, p, x):
    return F.linear(F.leaky_relu(F.linear(x, sd[p + '.linear1.weight'], sd[p + '.linear1.bias']), LRELU_SLOPE),
                    sd[p + '.linear2.weight'], sd[p + '.linear2.bias'])


def interpolate_embedding(x, rate):
    """transformer.py:57-73: frame i becomes the linear blend of the surrounding frames that are multiples of `rate`."""
    seq_len = x.shape[0]
    idx = torch.arange(seq_len, dtype=torch.int64)
    chunk, remain = idx // rate, idx % rate
    prev = x[chunk * rate]
    nxt = torch.cat([x[(chunk[:-1] + 1) * rate], x[-1].unsqueeze(0)], dim=0)
    return (prev / rate * (rate - remain.view(-1, 1, 1))) + (nxt / rate * remain.view(-1, 1, 1))


def transformer_forward(sd, src, src_mask, src_pos, tgt_mask, tgt_pos, rate, cfg=CFG, taps=None):
    """Transformer.forward(src, src_mask, src_pos, tgt, tgt_mask, tgt_pos, rate) with two_stage=True (tgt itself is not
    read then).  src [N, C, L]; masks bool [N, L]; pos [L, N, E].  Returns (joints [L, N, C], reco [L, N, C])."""
    nhead = cfg['nheads']
    x_in = src.permute(2, 0, 1)                                   # [L, N, C]
    l = x_in.shape[0]
    x = F.linear(x_in, sd['input_embed.weight'], sd['input_embed.bias'])
    eye = torch.eye(l).bool()                                     # encode(): a frame does not attend to itself
    for i in range(cfg['enc_layers']):
        p = 'encoder.layers.%d' % i
        x2 = _ln(sd, p + '.norm1', x)
        qk = x2 + src_pos
        x = x + _mha(sd, p + '.self_attn', qk, qk, x2, nhead, attn_mask=eye, key_padding_mask=src_mask)
        x = x + _ffn(sd, p, _ln(sd, p + '.norm2', x))
        if taps is not None:
            taps['enc%d' % i] = x
    mem = _ln(sd, 'encoder.norm', x)
    reco = F.linear(mem, sd['joints_embed.weight'], sd['joints_embed.bias']) + x_in
    interp = interpolate_embedding(reco, rate)
    y = F.linear(interp, sd['input_embed.weight'], sd['input_embed.bias'])
    mem_k = mem + src_pos
    for i in range(cfg['dec_layers']):
        p = 'decoder.layers.%d' % i
        y2 = _ln(sd, p + '.norm1', y)
        qk = y2 + tgt_pos
        y = y + _mha(sd, p + '.self_attn', qk, qk, y2, nhead, key_padding_mask=tgt_mask)
        y2 = _ln(sd, p + '.norm2', y)
        y = y + _mha(sd, p + '.multihead_attn', y2 + tgt_pos, mem_k, mem, nhead, key_padding_mask=src_mask)
        y = y + _ffn(sd, p, _ln(sd, p + '.norm3', y))
        if taps is not None:
            taps['dec%d' % i] = y
    out = _ln(sd, 'decoder.norm', y)
    joints = F.linear(out, sd['joints_embed.weight'], sd['joints_embed.bias']) + interp
    if taps is not None:
        taps.update(mem=mem, interp=interp)
    return joints, reco


def inference(sd, data, encoder_mask, decoder_mask, rate, cfg=CFG):
    """Model_inference.inference (inference.py:21-43): data [C, L] float32, masks bool [L] -> pred [1, C, L]."""
    src = data.unsqueeze(0)
    sm, tm = encoder_mask.unsqueeze(0), decoder_mask.unsqueeze(0)
    pos = position_encoding(1, data.shape[1], cfg['hidden_dim'])
    pred, _ = transformer_forward(sd, src, sm, pos, tm, pos, rate, cfg)
    return pred.permute(1, 2, 0)


def synth_motion(length, rate, seed=0, joints=CFG['input_joints']):
    """A smooth normalised joint sequence [C, L], the encoder mask of AMASS_dataset.get_openpose_data (every frame that is
    not a multiple of `rate` is hidden and zeroed in the input) and the all-visible decoder mask."""
    g = torch.Generator().manual_seed(1000 + seed)
    t = torch.linspace(0, 1, length)[None]
    f = torch.rand(joints, 1, generator=g) * 6 + 1
    ph = torch.rand(joints, 1, generator=g) * 6.28
    full = torch.sin(f * t * 6.28 + ph) * (0.5 + torch.rand(joints, 1, generator=g)) + 0.2 * torch.randn(joints, 1, generator=g)
    enc_mask = (torch.arange(length) % rate) != 0
    data = full * (~enc_mask)[None].float()
    return data.float(), enc_mask, torch.zeros(length, dtype=torch.bool)
