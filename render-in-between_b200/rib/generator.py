"""Drop-in replacement of the reference `models.generator.Generator`
(PGNR/models/generator.py:35-302) for inference.

Same constructor (`Generator(gen_cfg)`), same positional forward
(`forward(label, label_prev, img_fake, img_prev) -> (img_final, mask)`), same state-dict keys and
shapes (372 tensors for configs/HSM.yaml, including the unused `label_embedding.*` / `conv_mask.*`
and the spectral-norm triplets), so `utils.load_state_dict(net_G, path)` works unchanged.  The math
runs in the sm_100a kernels of librib_b200.so; there is no PyTorch fallback.
"""
import ctypes as C

import torch
from torch import nn

from ._lib import GenConfig, Tensor, check, lib
from .arch import Arch


class _Node(nn.Module):
    """Bare container used to reproduce the reference's module tree (and therefore its key names)."""


class Generator(nn.Module):
    def __init__(self, gen_cfg):
        super().__init__()
        self.gen_cfg = gen_cfg
        self.arch = Arch(gen_cfg)
        for key, shape, kind in self.arch.state_spec():
            parts = key.split('.')
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            t = torch.zeros(shape, dtype=torch.float32)
            if kind in ('sn_u', 'sn_v'):
                node.register_buffer(parts[-1], t)
            else:
                node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        self._handle = None
        self._workspaces = {}
        self._keepalive = None

    # -- packed-weight lifecycle --------------------------------------------------------------
    def _invalidate(self):
        if getattr(self, '_handle', None):
            lib.rib_generator_destroy(self._handle)
        self._handle = None
        self._workspaces = {}
        self._keepalive = None

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

    def _ensure_packed(self, device):
        if self._handle is not None:
            return
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        for k, v in sd.items():
            if not (v.is_cuda and v.dtype == torch.float32):
                raise RuntimeError('rib.Generator: parameter %s must be a CUDA float32 tensor '
                                   '(call .to("cuda") first); there is no CPU path' % k)
        sd = {k: v.contiguous() for k, v in sd.items()}
        a = self.arch
        cfg = GenConfig(a.label_nc, a.img_nc, a.nf, a.maxf, a.n_down, a.n_res, a.emb_nf, a.emb_max, a.emb_down,
                        a.mask_nf, a.mask_max, a.mask_down, a.mask_res)
        arr = (Tensor * len(sd))()
        names = []
        for i, (k, v) in enumerate(sd.items()):
            names.append(k.encode())
            arr[i].name = names[-1]
            arr[i].data = v.data_ptr()
            arr[i].numel = v.numel()
        handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.rib_generator_create(C.byref(cfg), arr, len(sd),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(handle)),
                  'rib_generator_create')
        self._handle = handle
        self._keepalive = sd  # the C side keeps pointers to the instance-norm affine tensors

    def _workspace(self, b, h, w, device):
        key = (b, h, w)
        ws = self._workspaces.get(key)
        if ws is None:
            need = lib.rib_generator_workspace_bytes(self._handle, b, h, w)
            if need < 0:
                check(int(need), 'rib_generator_workspace_bytes')
            buf = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            off = (-buf.data_ptr()) % 1024
            ws = (buf, buf.data_ptr() + off, need)
            self._workspaces = {key: ws}   # one live workspace: a plan is bound to its workspace
        return ws

    # -- reference API ------------------------------------------------------------------------
    def forward(self, label, label_prev, img_fake, img_prev):
        """label [B,22,H,W], label_prev (ignored, as in the reference), img_fake / img_prev [B,3,H,W]
        -> (img_final [B,3,H,W], mask [B,1,H,W])   (PGNR/models/generator.py:181-234)"""
        if not (label.is_cuda and label.dtype == torch.float32):
            raise RuntimeError('rib.Generator: label must be a CUDA float32 tensor')
        label = label.contiguous()
        b, c, h, w = label.shape
        if c != self.arch.label_nc:
            raise ValueError('rib.Generator: input shape mismatch')
        return self._run(label, b, h, w, img_fake, img_prev)

    def bind(self, b, h, w, device):
        """Builds the launch plan for a (b, h, w) batch and returns the device address of the generator's label
        input (16-bit [b][4][h][w][8]) inside its workspace: `rib.rasterize(..., planar_out=addr)` writes the
        label there, and `forward_bound` then runs without the fp32 label round trip."""
        self._ensure_packed(device)
        _, ws_ptr, ws_bytes = self._workspace(b, h, w, device)
        out = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.rib_generator_bind(self._handle, b, h, w, C.c_void_p(ws_ptr), ws_bytes, C.byref(out),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rib_generator_bind')
        return out.value

    def forward_bound(self, b, h, w, img_fake, img_prev):
        """forward() for a label that was rasterised straight into the buffer returned by bind()."""
        return self._run(None, b, h, w, img_fake, img_prev)

    def _run(self, label, b, h, w, img_fake, img_prev):
        if self.training:
            raise RuntimeError('rib.Generator implements the eval-mode forward only; call .eval() '
                               '(Evaluator does, PGNR/models/evaluator.py:170)')
        for name, t in (('img_fake', img_fake), ('img_prev', img_prev)):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError('rib.Generator: %s must be a CUDA float32 tensor' % name)
        img_fake, img_prev = img_fake.contiguous(), img_prev.contiguous()
        if tuple(img_fake.shape) != (b, 3, h, w) or tuple(img_prev.shape) != (b, 3, h, w):
            raise ValueError('rib.Generator: input shape mismatch')
        dev = img_fake.device
        self._ensure_packed(dev)
        out_img = torch.empty(b, 3, h, w, dtype=torch.float32, device=dev)
        out_mask = torch.empty(b, 1, h, w, dtype=torch.float32, device=dev)
        _, ws_ptr, ws_bytes = self._workspace(b, h, w, dev)
        with torch.cuda.device(dev):
            check(lib.rib_generator_forward(self._handle, b, h, w, label.data_ptr() if label is not None else None,
                                            img_fake.data_ptr(), img_prev.data_ptr(), out_img.data_ptr(),
                                            out_mask.data_ptr(), C.c_void_p(ws_ptr), ws_bytes,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  'rib_generator_forward')
        return out_img, out_mask

    # -- test / profiling hooks ---------------------------------------------------------------
    def plan_text(self):
        """One line per planned kernel launch of the last forward (layer, tiling, FLOPs)."""
        buf = C.create_string_buffer(1 << 16)
        check(lib.rib_generator_plan_text(self._handle, buf, len(buf)), 'rib_generator_plan_text')
        return buf.value.decode()

    def plan_dry_run(self, b, h, w):
        """(workspace bytes, plan text) of the launch plan for a (b, h, w) batch, computed WITHOUT a GPU: every layer's
        tiling is validated for the shape (tile geometry, shared-memory and TMEM budgets, imported tuning table)."""
        a = self.arch
        cfg = GenConfig(a.label_nc, a.img_nc, a.nf, a.maxf, a.n_down, a.n_res, a.emb_nf, a.emb_max, a.emb_down,
                        a.mask_nf, a.mask_max, a.mask_down, a.mask_res)
        need = C.c_longlong()
        buf = C.create_string_buffer(1 << 16)
        check(lib.rib_plan_dry_run(C.byref(cfg), b, h, w, C.byref(need), buf, len(buf)), 'rib_plan_dry_run')
        return need.value, buf.value.decode()

    @staticmethod
    def tune_log():
        """The auto-tuner's candidates / timings / choices for every launch shape tuned in this process."""
        buf = C.create_string_buffer(1 << 20)
        check(lib.rib_tune_log(buf, len(buf)), 'rib_tune_log')
        return buf.value.decode()

    def debug_tensor(self, name):
        """Intermediate activation of the last forward as an fp32 NCHW tensor (tests only)."""
        ptr, b, h, w, c, ld = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.rib_generator_debug_tensor(self._handle, name.encode(), C.byref(ptr), C.byref(b), C.byref(h),
                                             C.byref(w), C.byref(c), C.byref(ld)), 'rib_generator_debug_tensor')
        buf, _, _ = next(iter(self._workspaces.values()))
        off = ptr.value - buf.data_ptr()
        dt = torch.float16 if lib.rib_act_is_fp16() else torch.bfloat16
        b, h, w, c, ctot = b.value, h.value, w.value, c.value, ld.value
        # chunk-planar [B][ctot/8][H][W][8]; the view may be a channel slice (whole planes) of a wider buffer
        n = (b - 1) * ctot * h * w + c * h * w
        flat = buf[off:off + 2 * n].view(dt)
        t = torch.as_strided(flat, (b, c // 8, h, w, 8), (ctot * h * w, h * w * 8, w * 8, 8, 1))
        return t.permute(0, 1, 4, 2, 3).reshape(b, c, h, w).float().contiguous()
