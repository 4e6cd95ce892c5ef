"""On-GPU clip scheduler: the autoregressive loop of `Evaluator.evaluate_from_folder`
(PGNR/models/evaluator.py:238-266) without the per-frame host round trips.

Reference semantics kept: every `sample_rate`-th frame is a key frame that is copied through and
resets the chain (:240-244); a generated frame i uses label_i, the background `dain_i` and
`prev_img = fuse[i-1]` (:249-255) and is blended as fuse = pred*mask + dain*(1-mask) (:256-258).
Because key frames are known up front, AR step s (= i mod sample_rate) of all key-frame intervals
is one batch: 2x interpolation is a single batched generator call, r-x needs r-1 calls.
"""
import torch

from . import ops


class ClipRenderer:
    def __init__(self, generator, sample_rate=2):
        if sample_rate < 2 or sample_rate & (sample_rate - 1):
            raise ValueError('sample_rate must be a power of two >= 2 (evaluator.py:190)')
        self.gen = generator
        self.rate = sample_rate

    @staticmethod
    def seq_len(num_keyframes, sample_rate):
        return (num_keyframes - 1) * sample_rate + 1            # evaluator.py:191

    def render(self, key_frames, joints, backgrounds=None, flows=None, want_u8=True, want_mask=False):
        """key_frames [K,3,H,W] f32, joints [T,19,3] f64, and either backgrounds [T,3,H,W] f32 (the
        pre-computed DAIN frames of the reference) or flows [T,2,H,W] f32 from which the background of
        frame i is resampled out of the preceding key frame (stage A3).  All CUDA tensors.
        Returns dict(fuse [T,3,H,W] f32, u8 [T,H,W,3] uint8 | None, mask [T,1,H,W] | None)."""
        r = self.rate
        k, _, h, w = key_frames.shape
        t = self.seq_len(k, r)
        if joints.shape[0] != t:
            raise ValueError('need %d joint sets for %d key frames at %dx' % (t, k, r))
        if (backgrounds is None) == (flows is None):
            raise ValueError('pass exactly one of backgrounds / flows')
        dev = key_frames.device
        fuse = torch.empty(t, 3, h, w, dtype=torch.float32, device=dev)
        u8 = torch.empty(t, h, w, 3, dtype=torch.uint8, device=dev) if want_u8 else None
        mask_out = torch.zeros(t, 1, h, w, dtype=torch.float32, device=dev) if want_mask else None
        fuse[0::r] = key_frames                                   # key frames pass through (:240-244)
        label = ops.rasterize(joints, h, w)                       # all frames at once (label-only work)
        for s in range(1, r):
            idx = torch.arange(s, t, r, device=dev)               # frames i = j*r + s of every interval j
            lab = label[s::r]
            if backgrounds is not None:
                dain = backgrounds[s::r].contiguous()
            else:
                dain = ops.warp(key_frames[:-1].contiguous(), flows[s::r].contiguous())
            prev = fuse[s - 1::r][:k - 1].contiguous()            # fuse[i-1] (key frame when s == 1)
            pred, m = self.gen(lab.contiguous(), None, dain, prev)
            fused = ops.composite(pred, m, dain)
            fuse[idx] = fused
            if want_mask:
                mask_out[idx] = m
        if want_u8:
            # tensor2images of every output frame (utils.py:122-147): composite with mask == 1 is a pure convert
            ones = torch.ones(t, 1, h, w, dtype=torch.float32, device=dev)
            _, u8 = ops.composite(fuse, ones, fuse, want_u8=True)
        return {'fuse': fuse, 'u8': u8, 'mask': mask_out}
