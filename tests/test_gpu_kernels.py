"""GPU parity tests of the individual kernels, through the C ABI, against the CPU oracle."""
import ctypes as C
import glob
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import generator_oracle as go
from oracle import raster_oracle as ro
from rib.layout import from_planar, to_parity_planar, to_planar
from rib.synth import synth_flow, synth_image, synth_joints

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


# ---------------------------------------------------------------- A1 rasteriser (bit-exact)
def test_rasterize_golden_bit_exact(dev, golden_dir):
    import rib
    for f in sorted(glob.glob(os.path.join(golden_dir, 'raster_[0-9].npz'))):
        z = np.load(f)
        h, w = int(z['height']), int(z['width'])
        j = torch.from_numpy(z['joints'])[None].to(dev)
        lab = rib.rasterize(j, h, w).cpu().numpy()[0]
        ref = z['label']
        bad = np.argwhere(lab != ref)
        assert bad.shape[0] == 0, '%s: %d mismatching elements, first %s got %r want %r' % (
            os.path.basename(f), bad.shape[0], bad[:1], lab[tuple(bad[0])] if len(bad) else None,
            ref[tuple(bad[0])] if len(bad) else None)


def test_rasterize_fullsize_hash_and_oracle(dev, golden_dir):
    import rib
    spec = json.load(open(os.path.join(golden_dir, 'raster_fullsize_sha256.json')))
    for s in spec:
        h, w = s['height'], s['width']
        j = synth_joints(s['n_frames'], h, w, seed=s['seed'])
        lab = rib.rasterize(torch.from_numpy(j).to(dev), h, w).cpu().numpy()
        for t in range(s['n_frames']):
            got = hashlib.sha256(np.ascontiguousarray(lab[t]).tobytes()).hexdigest()
            if got != s['sha256'][t]:   # locate the damage with the oracle before failing
                ref = ro.label([(a[0], a[1]) for a in j[t]], [a[2] for a in j[t]], h, w)
                bad = np.argwhere(lab[t] != ref)
                pytest.fail('%dx%d frame %d: hash mismatch, %d elements differ from the oracle, channels %s' % (
                    h, w, t, bad.shape[0], sorted(set(bad[:, 0].tolist()))))


def test_rasterize_batch_edge_cases(dev):
    import rib
    h, w = 64, 80
    j = np.zeros((3, 19, 3))
    j[0, :, 2] = 0.0                                   # nothing valid -> skeleton all -1, heat-maps 0
    j[1] = synth_joints(1, h, w, seed=11)[0]
    j[1, :, 2] = 1.0
    j[2] = j[1]
    j[2, :, 0] += 0.25                                 # same pose shifted by a sub-pixel amount
    lab = rib.rasterize(torch.from_numpy(j).to(dev), h, w).cpu().numpy()
    assert (lab[0, :3] == -1.0).all() and (lab[0, 3:] == 0.0).all()
    for t in (1, 2):
        ref = ro.label([(a[0], a[1]) for a in j[t]], [a[2] for a in j[t]], h, w)
        assert np.array_equal(lab[t], ref)


# ---------------------------------------------------------------- A4 composite
def test_composite_matches_oracle(dev):
    import rib
    g = torch.Generator().manual_seed(0)
    b, h, w = 3, 64, 96
    img = torch.rand(b, 3, h, w, generator=g) * 2 - 1
    dain = torch.rand(b, 3, h, w, generator=g) * 2.4 - 1.2
    mask = torch.rand(b, 1, h, w, generator=g)
    ref = go.composite(img, mask, dain)
    out, u8 = rib.composite(img.to(dev), mask.to(dev), dain.to(dev), want_u8=True)
    assert (out.cpu() - ref).abs().max().item() <= 1e-6       # tolerance stated in SURVEY.md §8c
    assert torch.equal(out.cpu(), ref)                         # and in fact bit-exact (same roundings)
    assert torch.equal(u8.cpu(), go.to_uint8(ref))


# ---------------------------------------------------------------- A3 warp
@pytest.mark.parametrize('shape', [(64, 96), (20, 28), (9, 7)], ids=['px128', 'four', 'scalar'])
def test_warp_matches_grid_sample(dev, shape):
    """The three forms of the kernel: one pixel per lane (H * W % 128 == 0), four pixels per thread (W % 4 == 0), scalar."""
    import rib
    b, (h, w) = 2, shape
    src = synth_image(b, max(h, 16), max(w, 16), seed=5)[:, :, :h, :w].contiguous()
    flow = synth_flow(b, max(h, 64), max(w, 64), seed=5, max_px=8.0)[:, :, :h, :w].contiguous()
    flow[0, :, :4, :4] = 50.0          # far out of range -> border clamp
    flow[1, :, -4:, -4:] = -50.0
    ref = go.warp(src, flow)
    out = rib.warp(src.to(dev), flow.to(dev)).cpu()
    assert (out - ref).abs().max().item() <= 1e-5
    zero = rib.warp(src.to(dev), torch.zeros_like(flow).to(dev)).cpu()
    assert (zero - src).abs().max().item() <= 1e-5            # identity flow (grid round-trip in fp32)


# ---------------------------------------------------------------- implicit-GEMM convolution
def _act_dtype():
    from rib._lib import lib
    return torch.float16 if lib.rib_act_is_fp16() else torch.bfloat16


def _decode_stats(t):
    """Statistics slots are fixed-point integers (sum * 2^32, sum of squares * 2^24; csrc/common.cuh stat_add)."""
    q = t.view(torch.int64).double()
    return torch.stack([q[..., 0] / 2.0 ** 32, q[..., 1] / 2.0 ** 24], dim=-1)


def _encode_stats(t):
    q = torch.stack([torch.round(t[..., 0] * 2.0 ** 32), torch.round(t[..., 1] * 2.0 ** 24)], dim=-1)
    return q.to(torch.int64).view(torch.float64)


def _run_conv(dev, x_nchw, w, bias, k, stride, act, want_stats, simt):
    from rib._lib import check, lib
    dt = _act_dtype()
    b, cin, h, wd = x_nchw.shape
    cout = w.shape[0]
    # a stride-2 convolution reads its input from the parity-planar layout its producer writes
    x = (to_parity_planar(x_nchw, dt) if stride == 2 else to_planar(x_nchw, dt)).to(dev)
    out = torch.zeros(b, cout // 8, h // stride, wd // stride, 8, dtype=dt, device=dev)
    stats = torch.zeros(b, cout, 2, dtype=torch.float64, device=dev) if want_stats else None
    scratch = torch.empty(lib.rib_conv_test_scratch_bytes(cin, cout, k) + 1024, dtype=torch.uint8, device=dev)
    wd_, bd_ = w.contiguous().to(dev), bias.contiguous().to(dev)
    lib.rib_debug_set_simt(1 if simt else 0)
    try:
        check(lib.rib_conv_test(x.data_ptr(), wd_.data_ptr(), bd_.data_ptr(), out.data_ptr(),
                                stats.data_ptr() if want_stats else None, b, h, wd, cin, cout, k, stride, act,
                                scratch.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rib_conv_test')
        torch.cuda.synchronize()
    finally:
        lib.rib_debug_set_simt(0)
    return from_planar(out.float().cpu()), (_decode_stats(stats.cpu()) if want_stats else None)


CONV_CASES = [
    # (B, Cin, Cout, H, W, k, stride)
    (1, 16, 16, 16, 8, 3, 1),      # one tile, one stage of 16 channels (weights SW32), BN=16
    (1, 16, 16, 16, 16, 3, 1),     # two tiles side by side (halo columns come from the neighbour)
    (2, 32, 32, 16, 32, 3, 1),     # 32-channel stage (SW64)
    (1, 64, 64, 32, 32, 3, 1),     # 64-channel stage, 32-channel K groups
    (1, 128, 256, 16, 16, 3, 1),   # streamed weights, several stages, two N tiles
    (1, 256, 256, 32, 32, 3, 1),   # streamed weights with two M sub-tiles per super-tile (MT=2)
    (2, 64, 128, 32, 32, 3, 2),    # stride 2 through the four parity tiles
    (1, 512, 64, 16, 16, 1, 1),    # 1x1 (SPADE-shaped K), resident weights
    (2, 512, 128, 64, 16, 1, 1),   # 1x1, streamed weights, MT=2
    (1, 32, 16, 20, 30, 3, 1),     # ragged: H, W not multiples of the tile (HSM.yaml's 320x480 / 16)
    (1, 16, 32, 64, 96, 3, 2),     # stride 2, 16-channel stage
    (3, 16, 64, 128, 128, 3, 1),   # more tiles than persistent CTAs: tile loop, TMEM double buffering, stats per image
]


@pytest.mark.parametrize('simt', [True, False], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'B%d_%dto%d_%dx%d_k%ds%d' % c)
def test_conv_gemm_matches_conv2d(dev, case, simt):
    b, cin, cout, h, w, k, stride = case
    dt = _act_dtype()
    g = torch.Generator().manual_seed(cin * 1000 + cout + h)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    # reference on the same 16-bit-rounded operands, fp32 math (the kernel accumulates in fp32)
    xr, wr = x.to(dt).float(), wt.to(dt).float()
    ref = F.conv2d(xr, wr, bias, stride=stride, padding=k // 2)
    out, stats = _run_conv(dev, x, wt, bias, k, stride, act=0, want_stats=True, simt=simt)
    err = (out - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-3        # one 16-bit rounding of the output + accumulation order
    assert bool((err <= tol).all()), 'max err %.4g at %s (ref %.4g)' % (
        err.max().item(), np.unravel_index(err.argmax().item(), err.shape), ref.flatten()[err.argmax()].item())
    # instance-norm statistics from the epilogue (computed on the un-rounded fp32 accumulators)
    s_ref = torch.stack([ref.double().sum(dim=(2, 3)), (ref.double() ** 2).sum(dim=(2, 3))], dim=2)
    assert torch.allclose(stats, s_ref, rtol=2e-3, atol=2e-2), (stats - s_ref).abs().max().item()


PAIR_CASES = [
    # (B, Cin, Cout, H, W, stride, policy): 3x3 layers with 128-column tiles (H, W = input size); policy 4 = streamed
    # weights, 5 = each CTA keeps its half of the weight rows resident
    (2, 256, 256, 32, 32, 1, 4),     # MT=2, two N tiles, 8 super-tiles
    (3, 128, 128, 24, 40, 1, 4),     # MT=1 (H < 32), ragged tile rows / columns, 30 super-tiles
    (4, 512, 256, 64, 64, 1, 4),     # the mask net's res_flow shape at a smaller batch: more pair-tiles than CTA pairs
    (2, 256, 512, 64, 64, 2, 4),     # stride 2 from a parity-planar input (the embedder's emb_3 shape), four N tiles
    (2, 64, 128, 64, 64, 2, 5),      # emb_1's shape: resident half-tiles of the weights, deep halo ring
    (3, 64, 128, 24, 40, 1, 5),      # stride 1, ragged
]


@pytest.mark.parametrize('case', PAIR_CASES, ids=lambda c: 'B%d_%dto%d_%dx%d_s%d_p%d' % c)
def test_conv_gemm_cta_pairs_match_conv2d(dev, case, monkeypatch):
    """The CTA-pair form (cluster of two, tcgen05.mma.cta_group::2 with M = 256, each CTA staging half of the weight
    rows; policy 4 of conv_gemm_configure) against conv2d, and bit-identical to the single-CTA kernel."""
    b, cin, cout, h, w, stride, policy = case
    dt = _act_dtype()
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x.to(dt).float(), wt.to(dt).float(), bias, stride=stride, padding=1)
    single, s_single = _run_conv(dev, x, wt, bias, 3, stride, act=0, want_stats=True, simt=False)
    monkeypatch.setenv('RIB_TEST_POLICY', str(policy))
    out, stats = _run_conv(dev, x, wt, bias, 3, stride, act=0, want_stats=True, simt=False)
    monkeypatch.delenv('RIB_TEST_POLICY')
    err = (out - ref).abs()
    assert bool((err <= 2.0 ** -7 * ref.abs() + 2e-3).all()), 'max err %.4g' % err.max().item()
    s_ref = torch.stack([ref.double().sum(dim=(2, 3)), (ref.double() ** 2).sum(dim=(2, 3))], dim=2)
    assert torch.allclose(stats, s_ref, rtol=2e-3, atol=2e-2), (stats - s_ref).abs().max().item()
    # same K order and the same fp32 accumulation per output element: the pair computes the same bits
    assert torch.equal(out, single)


@pytest.mark.parametrize('case', [(2, 64, 128, 32, 32, 0), (1, 32, 64, 24, 40, 0), (3, 64, 128, 128, 64, 0), (1, 128, 128, 16, 16, 0),
                                  (2, 64, 128, 64, 64, 5), (1, 64, 128, 40, 24, 5)],
                         ids=lambda c: 'B%d_%dto%d_%dx%d_p%d' % c)
def test_conv_gemm_stride2_from_normal_layout(dev, case, monkeypatch):
    """Stride-2 conv whose input is a NORMAL planar map (emb_1 reads cond_0 this way) through the strided TMA parity
    views; with RIB_GATHER=1 in the environment of the process the kernel's extra warps gather the four parity tiles
    with cp.async instead (ConvGemmParams::a_gather, opt-in: measured slower)."""
    from rib._lib import check, lib
    b, cin, cout, h, w, policy = case
    if policy:     # 5: CTA pairs with resident half-tiles of the weights (5-D strided TMA views in their pair form)
        monkeypatch.setenv('RIB_TEST_POLICY', str(policy))
    dt = _act_dtype()
    g = torch.Generator().manual_seed(cin * 3 + cout + h)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x.to(dt).float(), wt.to(dt).float(), bias, stride=2, padding=1), 0.2)
    monkeypatch.setenv('RIB_TEST_S2_NORMAL', '1')
    out = torch.zeros(b, cout // 8, h // 2, w // 2, 8, dtype=dt, device=dev)
    scratch = torch.empty(lib.rib_conv_test_scratch_bytes(cin, cout, 3) + 1024, dtype=torch.uint8, device=dev)
    xd, wd_, bd_ = to_planar(x, dt).to(dev), wt.contiguous().to(dev), bias.contiguous().to(dev)
    check(lib.rib_conv_test(xd.data_ptr(), wd_.data_ptr(), bd_.data_ptr(), out.data_ptr(), None, b, h, w, cin, cout, 3, 2, 1,
                            scratch.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rib_conv_test')
    torch.cuda.synchronize()
    got = from_planar(out.float().cpu())
    err = (got - ref).abs()
    assert bool((err <= 2.0 ** -7 * ref.abs() + 2e-3).all()), 'max err %.4g' % err.max().item()


def test_conv_gemm_lrelu_epilogue(dev):
    b, cin, cout, h, w = 1, 32, 32, 16, 16
    g = torch.Generator().manual_seed(1)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.zeros(cout)
    dt = _act_dtype()
    ref = F.leaky_relu(F.conv2d(x.to(dt).float(), wt.to(dt).float(), bias, padding=1), 0.2)
    out, _ = _run_conv(dev, x, wt, bias, 3, 1, act=1, want_stats=False, simt=False)
    assert bool(((out - ref).abs() <= 2.0 ** -7 * ref.abs() + 2e-3).all())


# ---------------------------------------------------------------- in-kernel fusions of the mask network
def _run_conv_ex(dev, x_planar, w, bias, b, h, wd, cin, cout, stride, subpix, xf, out_shape, want_stats, simt):
    from rib._lib import check, lib
    dt = _act_dtype()
    out = torch.zeros(*out_shape, dtype=dt, device=dev)
    stats = torch.zeros(b, cout, 2, dtype=torch.float64, device=dev) if want_stats else None
    scratch = torch.empty(lib.rib_conv_test_scratch_bytes(cin, cout, 3) + 1024, dtype=torch.uint8, device=dev)
    wd_, bd_ = w.contiguous().to(dev), bias.contiguous().to(dev)
    xs, xw, xb = (t.contiguous().to(dev) for t in xf) if xf is not None else (None, None, None)
    lib.rib_debug_set_simt(1 if simt else 0)
    try:
        check(lib.rib_conv_test_ex(x_planar.data_ptr(), wd_.data_ptr(), bd_.data_ptr(), out.data_ptr(),
                                   stats.data_ptr() if want_stats else None, b, h, wd, cin, cout, 3, stride, 0,
                                   1 if subpix else 0, xs.data_ptr() if xf is not None else None,
                                   xw.data_ptr() if xf is not None else None, xb.data_ptr() if xf is not None else None,
                                   1, scratch.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              'rib_conv_test_ex')
        torch.cuda.synchronize()
    finally:
        lib.rib_debug_set_simt(0)
    return out.float().cpu(), (_decode_stats(stats.cpu()) if want_stats else None)


@pytest.mark.parametrize('ppc', [1, 4], ids=['ppc1', 'ppc4'])
@pytest.mark.parametrize('simt', [True, False], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('case', [(1, 16, 16, 16, 16), (2, 64, 32, 32, 32), (1, 128, 64, 16, 24), (1, 256, 128, 16, 16)],
                         ids=lambda c: 'B%d_%dto%d_%dx%d' % c)
def test_subpixel_conv_matches_upsample_conv(dev, case, simt, ppc, monkeypatch):
    """conv3x3(nearest_x2(x)) == the four 2x2 parity convs on the low-resolution map (generator.py:478-481); ppc4: the
    form that computes up to four output parities per CTA from one halo load."""
    monkeypatch.setenv('RIB_TEST_PPC', str(ppc))
    b, cin, cout, h, w = case
    dt = _act_dtype()
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(F.interpolate(x.to(dt).float(), scale_factor=2), wt, bias, padding=1)
    out_pp, stats = _run_conv_ex(dev, to_planar(x, dt).to(dev), wt, bias, b, h, w, cin, cout, 1, True, None,
                                 (b, cout // 8, 2, 2, h, w, 8), True, simt)
    # parity-planar [b][plane][py][px][y][x][e] -> NCHW at twice the size
    out = out_pp.permute(0, 1, 6, 4, 2, 5, 3).reshape(b, cout, 2 * h, 2 * w)
    # (a) against the same folded, 16-bit-rounded 2x2 weights (what the kernel multiplies with): tight
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}        # source taps landing on low-res neighbour a for parity p
    xp = F.pad(x.to(dt).float(), (1, 1, 1, 1))
    ref2 = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            w2 = torch.stack([torch.stack([wt[:, :, rows[py][a]][:, :, :, rows[px][bb]].sum(dim=(2, 3)) for bb in range(2)], dim=-1)
                              for a in range(2)], dim=-2)                                  # [cout, cin, 2, 2]
            o = F.conv2d(xp, w2.to(dt).float(), bias)                                       # [b, cout, h+1, w+1]
            ref2[:, :, py::2, px::2] = o[:, :, py:py + h, px:px + w]
    err2 = (out - ref2).abs()
    assert bool((err2 <= 2.0 ** -7 * ref2.abs() + 2e-3).all()), 'max err %.4g vs folded weights' % err2.max().item()
    # (b) against conv3x3(nearest_x2(x)) with the un-rounded weights: the fold itself is exact, only the single
    # 16-bit rounding of the summed taps (instead of one per tap) differs
    err = (out - ref).abs()
    assert bool((err <= 2.0 ** -6 * ref.abs() + 2e-2).all()), 'max err %.4g (ref %.4g)' % (
        err.max().item(), ref.flatten()[err.argmax()].item())
    assert float((err ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()) < 6e-3
    s_ref = torch.stack([ref.double().sum(dim=(2, 3)), (ref.double() ** 2).sum(dim=(2, 3))], dim=2)
    assert torch.allclose(stats, s_ref, rtol=1e-2, atol=0.2), (stats - s_ref).abs().max().item()


@pytest.mark.parametrize('simt', [True, False], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('case', [(2, 32, 64, 32, 32, 2), (1, 64, 128, 32, 48, 2), (2, 128, 128, 16, 16, 2), (1, 64, 64, 20, 28, 1),
                                  (3, 256, 128, 16, 16, 1)], ids=lambda c: 'B%d_%dto%d_%dx%d_s%d' % c)
def test_transform_conv_matches_norm_act_conv(dev, case, simt):
    """conv(lrelu(instance_norm_affine(x))) with the N-A applied to the halo tiles in shared memory (conv.py:56-69,
    order C-N-A of the producer) == the separate normalisation pass followed by the conv; zero padding stays zero."""
    b, cin, cout, h, w, stride = case
    dt = _act_dtype()
    g = torch.Generator().manual_seed(cin * 7 + cout + h + stride)
    x = (torch.randn(b, cin, h, w, generator=g) * 1.7 + 0.4).to(dt).float()
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    aw, ab = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.2
    xd = x.double()
    stats = torch.stack([xd.sum(dim=(2, 3)), (xd ** 2).sum(dim=(2, 3))], dim=2)           # [b, cin, 2]
    mean = stats[..., 0] / (h * w)
    var = (stats[..., 1] / (h * w) - mean ** 2).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (aw.double() * rstd).float()[:, :, None, None]
    shift = (ab.double() - mean * aw.double() * rstd).float()[:, :, None, None]
    xn = F.leaky_relu(x * scale + shift, 0.2).to(dt).float()
    ref = F.conv2d(xn, wt.to(dt).float(), bias, stride=stride, padding=1)
    xin = (to_parity_planar(x, dt) if stride == 2 else to_planar(x, dt)).to(dev)
    out, _ = _run_conv_ex(dev, xin, wt, bias, b, h, w, cin, cout, stride, False, (_encode_stats(stats), aw, ab),
                          (b, cout // 8, h // stride, w // stride, 8), False, simt)
    out = from_planar(out)
    err = (out - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 4e-3
    assert bool((err <= tol).all()), 'max err %.4g (ref %.4g)' % (err.max().item(), ref.flatten()[err.argmax()].item())


def test_frames_from_u8_matches_to_tensor_norm(dev):
    """uint8 HWC -> fp32 CHW == transforms.ToTensor + Normalize(0.5, 0.5) (HSM_auto_dataset.py:73-75), bit-exact,
    dense and strided; and the clip renderer gives the same frames for uint8 and fp32 key frames."""
    import rib
    g = torch.Generator().manual_seed(9)
    b, h, w = 5, 32, 48
    u8 = torch.randint(0, 256, (b, h, w, 3), generator=g, dtype=torch.uint8)
    u8[0, 0, :, 0] = torch.arange(48, dtype=torch.uint8)
    u8[0, 1, :, 1] = torch.arange(208, 256, dtype=torch.uint8)
    ref = (u8.permute(0, 3, 1, 2).float().div(255) - 0.5) / 0.5
    out = rib.frames_from_u8(u8.to(dev))
    assert torch.equal(out.cpu(), ref)
    clip = torch.zeros(2 * b, 3, h, w, device=dev)
    rib.frames_from_u8(u8.to(dev)[::2], out=clip[1::2][:3])
    assert torch.equal(clip[1::2][:3].cpu(), ref[::2]) and (clip[0::2] == 0).all()


def test_warp_accepts_half_precision_flows(dev):
    """rib.warp with a float16 flow == rib.warp with that flow converted to float32 (exact conversion on load),
    dense and strided."""
    import rib
    from rib.synth import synth_flow, synth_image
    b, h, w = 4, 64, 96
    src = synth_image(b, h, w, seed=3).to(dev)
    flow16 = synth_flow(2 * b, h, w, seed=4).to(dev).half()
    assert torch.equal(rib.warp(src, flow16[:b].contiguous()), rib.warp(src, flow16[:b].float()))
    assert torch.equal(rib.warp(src, flow16[1::2]), rib.warp(src, flow16[1::2].float().contiguous()))


# ---------------------------------------------------------------- AvgPool2d(3, 2, 1) between the encoder blocks
@pytest.mark.parametrize('case', [(2, 16, 16, 16), (1, 8, 64, 96), (3, 32, 12, 6), (1, 64, 128, 128), (2, 8, 2, 2),
                                  (1, 16, 256, 512)], ids=lambda c: 'B%d_C%d_%dx%d' % c)
def test_avgpool_matches_avg_pool2d(dev, case):
    """generator.py:203-208 (`AvgPool2d(3, stride=2, padding=1)`, zero padding counted): the strip kernel against
    F.avg_pool2d on the same 16-bit inputs; outputs within one 16-bit rounding (the summation order differs), statistics
    of the un-rounded pooled values against float64."""
    from rib._lib import check, lib
    b, c, h, w = case
    dt = _act_dtype()
    g = torch.Generator().manual_seed(c * 7 + h + w)
    x = (torch.randn(b, c, h, w, generator=g) * 2.0 + 0.3).to(dt)
    ref = F.avg_pool2d(x.double(), 3, 2, 1)
    xp = to_planar(x.float(), dt).to(dev)
    out = torch.full((b, c // 8, h // 2, w // 2, 8), float('nan'), dtype=dt, device=dev)
    stats = torch.zeros(b, c, 2, dtype=torch.float64, device=dev)
    for want_stats in (True, False):
        check(lib.rib_avgpool_test(xp.data_ptr(), out.data_ptr(), stats.data_ptr() if want_stats else None, b, h, w, c,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rib_avgpool_test')
        torch.cuda.synchronize()
        got = from_planar(out.float().cpu()).double()
        ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10   # one unit in the last place, relative
        assert torch.isfinite(got).all()
        assert ((got - ref).abs() <= ulp * ref.abs() + 1e-6).all(), float((got - ref).abs().max())
    st = _decode_stats(stats.cpu())
    assert torch.allclose(st[..., 0], ref.sum(dim=(2, 3)), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[..., 1], (ref * ref).sum(dim=(2, 3)), rtol=1e-5, atol=1e-3)
    # a frame's pooled map and statistics do not depend on the batch it is part of (bit for bit)
    out1 = torch.empty((1,) + tuple(out.shape[1:]), dtype=dt, device=dev)
    stats1 = torch.zeros(1, c, 2, dtype=torch.float64, device=dev)
    last = xp[b - 1:b].contiguous()
    check(lib.rib_avgpool_test(last.data_ptr(), out1.data_ptr(), stats1.data_ptr(), 1, h, w, c,
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rib_avgpool_test')
    torch.cuda.synchronize()
    assert torch.equal(out1[0].view(torch.int16), out[b - 1].view(torch.int16))
    assert torch.equal(stats1[0].view(torch.int64), stats[b - 1].view(torch.int64))
