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

    def render(self, key_frames, joints, backgrounds=None, flows=None, want_u8=True, want_mask=False, want_fuse=True):
        """key_frames [K,3,H,W] f32 in [-1,1] or [K,H,W,3] uint8 (decoded images; normalised on the GPU exactly like
        dataset.to_tensor_norm), joints [T,19,3] f64, and either backgrounds (the pre-computed DAIN frames of the
        reference: [T,3,H,W] f32, or [T-K,3,H,W] f32 holding only the generated frames in frame order) or flows
        ([T,2,H,W] or [T-K,2,H,W] f32) from which the background of frame i is resampled out of the preceding key
        frame (stage A3).  All CUDA tensors.
        Returns dict(fuse [T,3,H,W] f32 | None, u8 [T,H,W,3] uint8 | None, mask [T,1,H,W] | None).

        Only the labels of generated frames are rasterised: a key frame's label is consumed by the reference
        solely as `label_prev`, which Generator.forward never reads (generator.py:181-234).  Each AR step
        rasterises straight into the generator's input buffer and blends straight into frames s, s+r, ... of
        the clip, so no frame is gathered, scattered or converted twice."""
        r = self.rate
        key_u8 = key_frames if key_frames.dtype == torch.uint8 else None     # decoded images: converted below
        if key_u8 is not None:
            k, h, w = key_u8.shape[0], key_u8.shape[1], key_u8.shape[2]
        else:
            k, _, h, w = key_frames.shape
        t = self.seq_len(k, r)
        if joints.shape[0] != t:
            raise ValueError('need %d joint sets for %d key frames at %dx' % (t, k, r))
        if (backgrounds is None) == (flows is None):
            raise ValueError('pass exactly one of backgrounds / flows')
        if not (want_u8 or want_fuse):
            raise ValueError('nothing to return')
        per_frame = backgrounds if backgrounds is not None else flows
        if per_frame.shape[0] not in (t, t - k):
            raise ValueError('backgrounds / flows must cover all %d frames or the %d generated ones' % (t, t - k))
        gen_only = per_frame.shape[0] == t - k          # rows = generated frames only: frame k*r+s -> row k*(r-1)+s-1

        def of_step(x, s):
            return x[s - 1::r - 1] if gen_only else x[s::r]

        dev = key_u8.device if key_u8 is not None else key_frames.device
        fuse = torch.empty(t, 3, h, w, dtype=torch.float32, device=dev) if want_fuse else None
        u8 = torch.empty(t, h, w, 3, dtype=torch.uint8, device=dev) if want_u8 else None
        mask_out = torch.zeros(t, 1, h, w, dtype=torch.float32, device=dev) if want_mask else None
        # key frames pass through (:240-244).  Their output frame is tensor2images(to_tensor_norm(v)), which truncates
        # and is NOT v for 63 of the 256 levels: uint8 key frames are normalised and re-quantised in one pass.
        if key_u8 is not None:
            key_frames = ops.frames_from_u8(key_u8, out_u8=u8[0::r] if want_u8 else None)
            if want_fuse:
                fuse[0::r] = key_frames
        else:
            key_frames = key_frames.contiguous()
            ops.composite(key_frames, None, None, out=fuse[0::r] if want_fuse else None,
                          out_u8=u8[0::r] if want_u8 else None)
        b = k - 1
        prev = key_frames[:b]                                      # fuse[i-1] of step 1 is the key frame
        for s in range(1, r):
            lab_addr = self.gen.bind(b, h, w, dev)
            ops.rasterize(joints[s::r].contiguous(), h, w, planar_out=lab_addr, want_label=False)
            if backgrounds is not None:
                dain = of_step(backgrounds, s).contiguous()
            else:
                dain = ops.warp(key_frames[:b], of_step(flows, s))
            pred, m = self.gen.forward_bound(b, h, w, dain, prev)
            need_f32 = want_fuse or s + 1 < r                      # the next AR step reads this step's frames
            step_out = None
            if want_fuse:
                step_out = fuse[s::r]
            elif need_f32:
                step_out = torch.empty(b, 3, h, w, dtype=torch.float32, device=dev)
            res = ops.composite(pred, m, dain, out=step_out, out_u8=u8[s::r] if want_u8 else None)
            prev = (res[0] if isinstance(res, tuple) else res) if need_f32 else None
            if want_mask:
                mask_out[s::r] = m
        return {'fuse': fuse, 'u8': u8, 'mask': mask_out}

    def render_clips(self, key_frames, joints, backgrounds=None, flows=None):
        """Several clips of the same shape in one pass: key_frames [C,K,H,W,3] uint8 (or [C,K,3,H,W] f32), joints
        [C,T,19,3] f64, and backgrounds [C,T-K,3,H,W] f32 or flows [C,T-K,2,H,W] f32 for the generated frames.  Clips are
        independent, so AR step s of ALL intervals of ALL clips is one generator batch of C * (K - 1) frames: two 4x clips
        of 17 key frames run at the batch-32 efficiency of one 2x clip of 33.  Returns uint8 [C,T,H,W,3] (the frames
        `render` gives clip by clip; the statistics of a batch are reduced over other tile ranges, so single values may
        differ by one 16-bit rounding)."""
        r = self.rate
        if key_frames.dim() != 5:
            raise ValueError('key_frames must be [C, K, H, W, 3] uint8 or [C, K, 3, H, W] float32')
        if (backgrounds is None) == (flows is None):
            raise ValueError('pass exactly one of backgrounds / flows')
        nclip, k = key_frames.shape[0], key_frames.shape[1]
        dev = key_frames.device
        t = self.seq_len(k, r)
        b = k - 1
        if key_frames.dtype == torch.uint8:
            h, w = key_frames.shape[2], key_frames.shape[3]
        else:
            h, w = key_frames.shape[3], key_frames.shape[4]
        per_frame = backgrounds if backgrounds is not None else flows
        if tuple(joints.shape[:2]) != (nclip, t) or tuple(per_frame.shape[:2]) != (nclip, t - k):
            raise ValueError('render_clips: need joints [C, %d, 19, 3] and %d generated-frame rows per clip' % (t, t - k))
        u8 = torch.empty(nclip, t, h, w, 3, dtype=torch.uint8, device=dev)
        keys = torch.empty(nclip, k, 3, h, w, dtype=torch.float32, device=dev)
        for c in range(nclip):                                   # key frames pass through (evaluator.py:240-244)
            if key_frames.dtype == torch.uint8:
                ops.frames_from_u8(key_frames[c], out=keys[c], out_u8=u8[c, 0::r])
            else:
                keys[c].copy_(key_frames[c])
                ops.composite(keys[c], None, None, out_u8=u8[c, 0::r])
        label_bytes = 32 * h * w * 2                             # one frame of the generator's planar 16-bit label input
        prev = keys[:, :b].reshape(nclip * b, 3, h, w)           # fuse[i-1] of step 1 is the key frame (a copy: clips are not adjacent)
        for s in range(1, r):
            lab_addr = self.gen.bind(nclip * b, h, w, dev)
            dain = torch.empty(nclip * b, 3, h, w, dtype=torch.float32, device=dev)
            for c in range(nclip):
                ops.rasterize(joints[c, s::r].contiguous(), h, w, planar_out=lab_addr + c * b * label_bytes, want_label=False)
                if backgrounds is not None:
                    dain[c * b:(c + 1) * b].copy_(backgrounds[c, s - 1::r - 1])
                else:
                    ops.warp(keys[c, :b], flows[c, s - 1::r - 1], out=dain[c * b:(c + 1) * b])
            pred, m = self.gen.forward_bound(nclip * b, h, w, dain, prev)
            nxt = torch.empty(nclip * b, 3, h, w, dtype=torch.float32, device=dev) if s + 1 < r else None
            for c in range(nclip):
                sl = slice(c * b, (c + 1) * b)
                ops.composite(pred[sl], m[sl], dain[sl], out=nxt[sl] if nxt is not None else None, out_u8=u8[c, s::r])
            prev = nxt
        return u8
