"""GPU parity tests of the on-GPU clip scheduler (AR loop of evaluator.py:238-266) against a CPU restatement
of that loop built from the oracle pieces, and of the fused label path (rasterise -> generator input)."""
import numpy as np
import pytest
import torch

from oracle import generator_oracle as go
from oracle import raster_oracle as ro
from rib.synth import synth_flow, synth_image, synth_joints

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def gen(dev, arch, synth_sd):
    from rib.config import default_gen_cfg
    from rib.generator import Generator
    g = Generator(default_gen_cfg())
    g.load_state_dict(synth_sd, strict=True)
    return g.to(dev).eval()


def oracle_clip(sd, arch, key, joints, dain, rate):
    """evaluate_from_folder's phase-2 loop (evaluator.py:238-266), one frame at a time on the CPU."""
    t = joints.shape[0]
    h, w = key.shape[-2:]
    fuse, masks = [], []
    for i in range(t):
        if i % rate == 0:
            fuse.append(key[i // rate][None])
            masks.append(torch.zeros(1, 1, h, w))
            continue
        lab = torch.from_numpy(ro.label([(a[0], a[1]) for a in joints[i]], [a[2] for a in joints[i]], h, w))[None]
        with torch.no_grad():
            img, m = go.generator_forward(sd, arch, lab, dain[i][None], fuse[-1])
            fuse.append(go.composite(img, m, dain[i][None]))
        masks.append(m)
    return torch.cat(fuse), torch.cat(masks)


@pytest.mark.parametrize('rate,nkey', [(2, 4), (4, 3)], ids=['2x', '4x'])
def test_clip_renderer_matches_oracle_loop(dev, gen, arch, synth_sd, rate, nkey):
    from rib.clip import ClipRenderer
    h, w = 64, 96
    t = (nkey - 1) * rate + 1
    key = synth_image(nkey, h, w, seed=21)
    joints = synth_joints(t, h, w, seed=22)
    dain = synth_image(t, h, w, seed=23)
    ref_fuse, ref_mask = oracle_clip(synth_sd, arch, key, joints, dain, rate)
    r = ClipRenderer(gen, sample_rate=rate)
    with torch.no_grad():
        out = r.render(key.to(dev), torch.from_numpy(joints).to(dev), backgrounds=dain.to(dev), want_u8=True,
                       want_mask=True, want_fuse=True)
    fuse = out['fuse'].cpu()
    assert torch.equal(fuse[0::rate], key)                       # key frames pass through untouched
    assert (out['mask'].cpu()[0::rate] == 0).all()
    p = go.psnr(fuse, ref_fuse)
    assert p >= 45.0, 'clip PSNR %.2f dB over the AR chain (rate %d)' % (p, rate)
    assert torch.equal(out['u8'].cpu(), go.to_uint8(fuse))       # uint8 frames == tensor2images(fuse)
    # u8-only mode (what the benchmark uses) returns the same frames
    with torch.no_grad():
        out2 = r.render(key.to(dev), torch.from_numpy(joints).to(dev), backgrounds=dain.to(dev), want_u8=True,
                        want_fuse=False)
    # instance-norm statistics are accumulated as fixed-point integers (order-independent atomics), so a second run is
    # bit-identical
    assert out2['fuse'] is None and torch.equal(out2['u8'], out['u8'])


@pytest.mark.parametrize('rate,nkey', [(2, 33), (4, 17)], ids=['c2_2x_K33', 'c3_4x_K17'])
def test_clip_renderer_512_bench_clips(dev, gen, arch, synth_sd, rate, nkey):
    """The clips bench.py renders (BASELINE configs[2]: 2x, 33 key frames; configs[3]: 4x, 17 key frames; 65 frames at
    512x512, backgrounds resampled from the preceding key frame) against the CPU restatement of evaluator.py:238-266 on
    a subset of key-frame intervals (the oracle costs ~1.5 s per frame at this size).  An interval only depends on its
    own key frame, so the oracle chain of interval k is exact without the rest of the clip."""
    from rib.clip import ClipRenderer
    h = w = 512
    t = (nkey - 1) * rate + 1
    key = synth_image(nkey, h, w, seed=61)
    joints = synth_joints(t, h, w, seed=62)
    flows = synth_flow(t, h, w, seed=63)
    r = ClipRenderer(gen, sample_rate=rate)
    with torch.no_grad():
        out = r.render(key.to(dev), torch.from_numpy(joints).to(dev), flows=flows.to(dev), want_u8=True, want_fuse=True)
    fuse, u8 = out['fuse'].cpu(), out['u8'].cpu()
    assert torch.equal(fuse[0::rate], key)
    assert torch.equal(u8, go.to_uint8(fuse))
    worst = 1e9
    for k in ([0, nkey // 2, nkey - 2] if rate == 2 else [0, nkey - 2]):
        prev = key[k][None]
        for s in range(1, rate):
            i = k * rate + s
            lab = torch.from_numpy(ro.label([(a[0], a[1]) for a in joints[i]], [a[2] for a in joints[i]], h, w))[None]
            dain = go.warp(key[k][None], flows[i][None])
            with torch.no_grad():
                img, m = go.generator_forward(synth_sd, arch, lab, dain, prev)
                prev = go.composite(img, m, dain)
            p = go.psnr(fuse[i][None], prev)
            worst = min(worst, p)
            assert p >= 45.0, 'frame %d (interval %d, AR step %d): %.2f dB' % (i, k, s, p)
            assert (fuse[i][None] - prev).abs().max().item() <= 0.10
    print('512x512 %dx clip: worst checked frame %.2f dB' % (rate, worst))


def test_clip_renderer_flow_backgrounds(dev, gen):
    """backgrounds resampled from the preceding key frame (stage A3) == explicit backgrounds."""
    import rib
    from rib.clip import ClipRenderer
    h, w, nkey, rate = 64, 96, 3, 2
    t = (nkey - 1) * rate + 1
    key = synth_image(nkey, h, w, seed=31).to(dev)
    joints = torch.from_numpy(synth_joints(t, h, w, seed=32)).to(dev)
    flows = synth_flow(t, h, w, seed=33).to(dev)
    bg = torch.zeros(t, 3, h, w, device=dev)
    bg[1::rate] = rib.warp(key[:-1], flows[1::rate])
    r = ClipRenderer(gen, sample_rate=rate)
    with torch.no_grad():
        a = r.render(key, joints, flows=flows)
        b = r.render(key, joints, backgrounds=bg)
    assert torch.equal(a['fuse'], b['fuse']) and torch.equal(a['u8'], b['u8'])


@pytest.mark.parametrize('rate', [2, 4])
def test_clip_renderer_u8_keys_and_generated_only_flows(dev, gen, rate):
    """The lean upload of bench.py / evaluate_from_folder: key frames as decoded uint8 images and flows (or
    backgrounds) for the generated frames only give the same clip as fp32 key frames and per-frame tensors."""
    import rib
    from rib.clip import ClipRenderer
    h, w, nkey = 64, 96, 3
    t = (nkey - 1) * rate + 1
    g = torch.Generator().manual_seed(70 + rate)
    key_u8 = torch.randint(0, 256, (nkey, h, w, 3), generator=g, dtype=torch.uint8).to(dev)
    key = rib.frames_from_u8(key_u8)
    joints = torch.from_numpy(synth_joints(t, h, w, seed=71)).to(dev)
    flows = synth_flow(t, h, w, seed=72).to(dev)
    gen_rows = [i for i in range(t) if i % rate]
    r = ClipRenderer(gen, sample_rate=rate)
    with torch.no_grad():
        a = r.render(key, joints, flows=flows)
        b = r.render(key_u8, joints, flows=flows[gen_rows].contiguous())
        bg = torch.zeros(t, 3, h, w, device=dev)
        for s in range(1, rate):
            bg[s::rate] = rib.warp(key[:-1], flows[s::rate])
        c = r.render(key_u8, joints, backgrounds=bg[gen_rows].contiguous(), want_fuse=False)
    assert torch.equal(a['fuse'], b['fuse']) and torch.equal(a['u8'], b['u8'])
    assert torch.equal(c['u8'], a['u8'])             # a third run through another entry: bit-identical


def test_bound_label_path_is_bit_identical(dev, gen):
    """rasterise straight into the generator's input buffer == rasterise fp32 label -> forward(label)."""
    import rib
    b, h, w = 3, 64, 96
    joints = torch.from_numpy(synth_joints(b, h, w, seed=41)).to(dev)
    fake, prev = synth_image(b, h, w, seed=42).to(dev), synth_image(b, h, w, seed=43).to(dev)
    with torch.no_grad():
        label = rib.rasterize(joints, h, w)
        img0, mask0 = gen(label, None, fake, prev)
        addr = gen.bind(b, h, w, dev)
        both = rib.rasterize(joints, h, w, planar_out=addr)
        img1, mask1 = gen.forward_bound(b, h, w, fake, prev)
    assert torch.equal(both, label)
    assert torch.equal(img0, img1) and torch.equal(mask0, mask1)


def test_strided_composite_and_warp(dev):
    import rib
    g = torch.Generator().manual_seed(5)
    b, h, w, r = 3, 32, 48, 2
    img = (torch.rand(b, 3, h, w, generator=g) * 2 - 1).to(dev)
    dain = (torch.rand(b, 3, h, w, generator=g) * 2 - 1).to(dev)
    mask = torch.rand(b, 1, h, w, generator=g).to(dev)
    dense, dense_u8 = rib.composite(img, mask, dain, want_u8=True)
    clip = torch.zeros(b * r, 3, h, w, device=dev)
    clip_u8 = torch.zeros(b * r, h, w, 3, dtype=torch.uint8, device=dev)
    rib.composite(img, mask, dain, out=clip[1::r], out_u8=clip_u8[1::r])
    assert torch.equal(clip[1::r], dense) and torch.equal(clip_u8[1::r], dense_u8)
    assert (clip[0::r] == 0).all() and (clip_u8[0::r] == 0).all()
    rib.composite(img, None, None, out=clip[0::r], out_u8=clip_u8[0::r])       # pass-through (key frames)
    assert torch.equal(clip[0::r], img) and torch.equal(clip_u8[0::r].cpu(), go.to_uint8(img.cpu()))
    flows = synth_flow(b * r, h, w, seed=6).to(dev)
    assert torch.equal(rib.warp(img, flows[1::r]), rib.warp(img, flows[1::r].contiguous()))


def test_render_clips_matches_clip_by_clip(dev, gen):
    """Two 4x clips rendered as ONE batch per AR step (ClipRenderer.render_clips) against the same clips rendered one by
    one.  (a) With one clip per call both entries launch the same plan (batch K-1) and must agree bit for bit: the two
    code paths sequence the same kernels.  (b) With two clips per batch the launch shape differs (batch 2 x (K-1)).  With
    the static tilings (RIB_AUTOTUNE=0) the two renderings are still bit-identical; with the plan-time auto-tuner on, the
    two shapes may get different tilings per layer, the instance-norm statistics are then reduced over other tile
    ranges, single 16-bit roundings flip, and at this test size the deepest normalised maps have 2 x 3 pixels, which
    amplifies a flipped rounding into a few grey levels over three dependent AR passes (profiles/r3d_clip_plan_noise.txt:
    tuner off 0 differences; tuner on <= 4 levels, 52.3-52.7 dB, which tilings win the timing differs from box to box).
    The gate is the north star's own tolerance for final frames: >= 45 dB between the two renderings."""
    from rib.clip import ClipRenderer
    h, w, nkey, rate, nclip = 64, 96, 3, 4, 2
    t = (nkey - 1) * rate + 1
    g = torch.Generator().manual_seed(91)
    key_u8 = torch.randint(0, 256, (nclip, nkey, h, w, 3), generator=g, dtype=torch.uint8).to(dev)
    joints = torch.stack([torch.from_numpy(synth_joints(t, h, w, seed=92 + c)) for c in range(nclip)]).to(dev)
    gen_rows = [i for i in range(t) if i % rate]
    flows = torch.stack([synth_flow(t, h, w, seed=95 + c)[gen_rows] for c in range(nclip)]).to(dev)
    r = ClipRenderer(gen, sample_rate=rate)
    with torch.no_grad():
        both = r.render_clips(key_u8, joints, flows=flows)
        single = torch.stack([r.render(key_u8[c], joints[c], flows=flows[c], want_u8=True, want_fuse=False)['u8']
                              for c in range(nclip)])
        one_by_one = torch.cat([r.render_clips(key_u8[c:c + 1], joints[c:c + 1], flows=flows[c:c + 1]) for c in range(nclip)])
    assert both.shape == single.shape == (nclip, t, h, w, 3)
    assert torch.equal(one_by_one, single)                              # (a) same plan: bit-identical
    assert torch.equal(both[:, 0::rate], single[:, 0::rate])            # key frames: identical
    d = (both.int() - single.int()).abs()
    gen_mask = torch.tensor([i % rate != 0 for i in range(t)], device=dev)
    mse = (d[:, gen_mask].float() ** 2).mean().item()
    psnr = 10.0 * np.log10(255.0 ** 2 / max(mse, 1e-12))
    line = ('render_clips vs clip by clip (64x96, 4x): max |d| %d levels, %.2f %% of the values differ, %.1f dB'
            % (d.max().item(), 100.0 * (d > 0).float().mean().item(), psnr))
    print(line)
    import os
    rep = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(rep):
        with open(os.path.join(rep, 'parity_report.txt'), 'a') as f:
            f.write(line + '\n')
    assert d.max().item() <= 16 and psnr >= 45.0, (d.max().item(), (d > 0).float().mean().item(), psnr)
