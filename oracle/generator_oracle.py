"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement (plain torch functional ops) of the reference
generator forward, the warp stage and the composite.

Never imported by the product path; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs use it, as the checker or as the timed CPU baseline.

Parity pin: compared against the unmodified reference `Generator` (imported from /root/reference
in the build container) by tests/test_oracle_vs_reference.py, and against the committed outputs
of that reference in tests/golden/generator_*.npz (made by oracle/make_golden.py).
The warp stage has no reference call site: "parity unpinned" for A3 — the oracle is
torch.nn.functional.grid_sample itself.

Restates:
  Generator.forward           PGNR/models/generator.py:181-234
  LabelEmbedder.forward       PGNR/models/generator.py:360-387 (arch='encoder')
  MaskGenerator.forward       PGNR/models/generator.py:493-510
  Res2dBlock / Conv2dBlock    PGNR/models/layers/residual.py:125-151, conv.py:56-69
  SpatiallyAdaptiveNorm       PGNR/models/layers/activation_norm.py:211-234
  spectral_norm (eval mode)   PGNR/models/layers/weight_norm.py:84-85 -> W / (u . (W_mat v))
  composite                   PGNR/models/evaluator.py:256-258
  tensor2images               PGNR/utils/utils.py:122-147
"""
import torch
import torch.nn.functional as F


def sn_weight(sd, prefix):
    """Eval-mode spectral-normalised weight of `<prefix>.weight_{orig,u,v}`."""
    w = sd[prefix + '.weight_orig']
    sigma = torch.dot(sd[prefix + '.weight_u'], torch.mv(w.reshape(w.shape[0], -1), sd[prefix + '.weight_v']))
    return w / sigma


def _conv(sd, prefix, x, stride=1, sn=True):
    w = sn_weight(sd, prefix) if sn else sd[prefix + '.weight']
    return F.conv2d(x, w, sd[prefix + '.bias'], stride=stride, padding=w.shape[-1] // 2)


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def _spade(sd, prefix, x, cond):
    """prefix = '<block>.conv_block_k.layers.norm'"""
    p = prefix + '.mlps.0.0.layers.conv'
    gb = F.conv2d(cond, sd[p + '.weight'], sd[p + '.bias'])
    gamma, beta = gb.chunk(2, dim=1)
    return F.instance_norm(x, eps=1e-5) * (1 + gamma) + beta


def _spade_block(sd, name, x, cond, taps=None):
    p0, p1, ps = (name + '.conv_block_%s.layers' % k for k in ('0', '1', 's'))
    dx = _conv(sd, p0 + '.conv', _lrelu(_spade(sd, p0 + '.norm', x, cond)))
    dx = _conv(sd, p1 + '.conv', _lrelu(_spade(sd, p1 + '.norm', dx, cond)))
    if (ps + '.conv.weight_orig') in sd:
        xs = _conv(sd, ps + '.conv', _spade(sd, ps + '.norm', x, cond))
    else:
        xs = x
    return xs + dx


def _cna(sd, prefix, x, stride=1, act=True):
    """conv(SN) -> InstanceNorm(affine) -> LeakyReLU  (MaskGenerator blocks)."""
    y = _conv(sd, prefix + '.layers.conv', x, stride)
    y = F.instance_norm(y, weight=sd[prefix + '.layers.norm.weight'], bias=sd[prefix + '.layers.norm.bias'],
                        eps=1e-5)
    return _lrelu(y) if act else y


def embed(sd, arch, x, name='ref_embedding'):
    out = [_lrelu(_conv(sd, name + '.conv_first.layers.conv', x))]
    for i in range(arch.emb_down):
        out.append(_lrelu(_conv(sd, '%s.down_%d.layers.conv' % (name, i), out[-1], stride=2)))
    return out


def mask_net(sd, arch, label, img9, taps=None):
    f = 'flow_network_temp'
    a, b = label, img9
    for i in range(arch.mask_down + 1):
        a = _cna(sd, '%s.down_lbl.%d' % (f, i), a, stride=1 if i == 0 else 2)
        b = _cna(sd, '%s.down_img.%d' % (f, i), b, stride=1 if i == 0 else 2)
    x = torch.cat([a, b], dim=1)
    if taps is not None:
        taps['mask.cat'] = x
    for i in range(arch.mask_res):
        p = '%s.res_flow.%d' % (f, i)
        dx = _cna(sd, p + '.conv_block_0', x)
        dx = _cna(sd, p + '.conv_block_1', dx, act=False)
        xs = _cna(sd, p + '.conv_block_s', x, act=False) if i == 0 else x
        x = xs + dx
        if taps is not None:
            taps['mask.res.%d' % i] = x
    for n in range(arch.mask_down):
        x = F.interpolate(x, scale_factor=2)
        x = _cna(sd, '%s.up_flow.%d' % (f, 2 * n + 1), x)
        if taps is not None:
            taps['mask.up.%d' % n] = x
    return torch.sigmoid(_conv(sd, f + '.conv_mask.0.layers.conv', x, sn=False))


def generator_forward(sd, arch, label, img_fake, img_prev, taps=None):
    """(img_final, mask) exactly as Generator.forward(label, label_prev, img_fake, img_prev).

    `taps`, if a dict, receives intermediate activations (NCHW fp32) keyed by name for
    per-kernel parity checks.
    """
    conds = embed(sd, arch, torch.cat([img_fake, img_prev], dim=1))
    x = _conv(sd, 'down_first.layers.conv', label, sn=False)
    if taps is not None:
        for i, c in enumerate(conds):
            taps['cond_%d' % i] = c
        taps['down_first'] = x
    for i in range(arch.n_down + 1):
        x = _spade_block(sd, 'down_%d' % i, x, conds[min(arch.emb_down, i)])
        if taps is not None:
            taps['down_%d' % i] = x
        if i != arch.n_down:
            x = F.avg_pool2d(x, 3, 2, 1)
    for i in range(arch.n_res):
        x = _spade_block(sd, 'res_%d' % i, x, conds[min(arch.emb_down, arch.n_down + 1)])
        if taps is not None:
            taps['res_%d' % i] = x
    for i in range(arch.n_down, -1, -1):
        x = _spade_block(sd, 'up_%d' % i, x, conds[min(i, arch.emb_down)])
        if taps is not None:
            taps['up_%d' % i] = x
        if i != 0:
            x = F.interpolate(x, scale_factor=2)
    img_final = torch.tanh(_conv(sd, 'conv_img.layers.conv', _lrelu(x), sn=False))
    mask = mask_net(sd, arch, label, torch.cat([img_prev, img_fake, img_final], dim=1), taps=taps)
    return img_final, mask


def composite(pred_img, mask, dain_img):
    """evaluator.py:256-258."""
    m = mask.repeat(1, 3, 1, 1)
    return pred_img * m + dain_img * (1 - m)


def to_uint8(frames):
    """tensor2images (utils.py:137-142): (x*0.5+0.5).clip(0,1)*255 truncated; NCHW f32 -> NHWC u8.
    The reference does this arithmetic in float64 numpy (image_numpy * std + mean with f64 arrays)."""
    x = frames.double().permute(0, 2, 3, 1) * 0.5 + 0.5
    return (x.clamp(0, 1) * 255.0).to(torch.uint8)


def warp(src, flow):
    """A3: bilinear resample of `src` at (x + flow_x, y + flow_y), border padding, align_corners=True
    (the upstream imaginaire `resample` convention; no call site in the reference tree)."""
    b, _, h, w = src.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing='ij')
    gx = (xs[None] + flow[:, 0]) * (2.0 / max(w - 1, 1)) - 1.0
    gy = (ys[None] + flow[:, 1]) * (2.0 / max(h - 1, 1)) - 1.0
    grid = torch.stack([gx, gy], dim=-1)
    return F.grid_sample(src, grid, mode='bilinear', padding_mode='border', align_corners=True)


def psnr(pred, target):
    """compute_metrics convention (evaluator.py:149-163): de-normalise, clamp to [0,1], 10 log10(1/MSE)."""
    p = (pred * 0.5 + 0.5).clamp(0, 1).double()
    t = (target * 0.5 + 0.5).clamp(0, 1).double()
    mse = ((p - t) ** 2).mean()
    return float(10.0 * torch.log10(1.0 / mse)) if mse > 0 else float('inf')
