"""Architecture bookkeeping for the pose-guided generator.

Derives, from the reference's `gen:` config block, the layer list of
`Generator.__init__` / `init_temporal_network` / `LabelEmbedder` / `MaskGenerator`
(PGNR/models/generator.py:43-178, :315-348, :423-491) and the exact state-dict key set + shapes
(372 tensors for HSM.yaml) so that a reference checkpoint loads with strict=True.
"""


class Arch:
    def __init__(self, gen_cfg):
        g = gen_cfg
        self.label_nc = g.input_label_nc
        self.img_nc = g.input_image_nc
        self.nf = getattr(g, 'num_filters', 32)
        self.maxf = getattr(g, 'max_num_filters', 1024)
        self.num_layers = getattr(g, 'num_layers', 7)
        self.n_down = getattr(g, 'num_downsamples_img', 4)       # generator.py:50
        self.ksize = getattr(g, 'kernel_size', 3)
        self.n_res = int(-(-(self.num_layers - self.n_down) // 2) * 2)   # generator.py:134-135
        e = g.embed
        self.emb_nf = getattr(e, 'num_filters', 32)
        self.emb_max = getattr(e, 'max_num_filters', 1024)
        self.emb_down = getattr(e, 'num_downsamples', 5)
        m = g.mask
        self.mask_nf = getattr(m, 'num_filters', 32)
        self.mask_max = getattr(m, 'max_num_filters', 1024)
        self.mask_down = getattr(m, 'num_downsamples', 5)
        self.mask_res = getattr(m, 'num_res_blocks', 6)
        self._check(g)

    def _check(self, g):
        # The CUDA path implements the HSM.yaml family only; fail loudly on anything else.
        e, m, a = g.embed, g.mask, g.activation_norm_params
        ok = (self.ksize == 3 and getattr(e, 'kernel_size', 3) == 3 and getattr(m, 'kernel_size', 3) == 3
              and getattr(g, 'weight_norm_type', 'spectral') == 'spectral'
              and getattr(e, 'weight_norm_type', 'spectral') == 'spectral'
              and getattr(m, 'weight_norm_type', 'spectral') == 'spectral'
              and g.activation_norm_type == 'spatially_adaptive'
              and getattr(a, 'activation_norm_type', '') == 'instance'
              and getattr(a, 'num_filters', 0) == 0 and getattr(a, 'kernel_size', 1) == 1
              and getattr(m, 'activation_norm_type', '') == 'instance'
              and getattr(e, 'arch', 'encoderdecoder') == 'encoder'
              and getattr(e, 'use_embed', True)
              and self.emb_down <= self.n_down
              and self.nf % 16 == 0 and self.emb_nf % 16 == 0 and self.mask_nf % 16 == 0)
        if not ok:
            raise NotImplementedError('rib: generator config outside the HSM.yaml family')

    # generator.py:25-33
    def nfilt(self, i):
        return min(self.maxf, self.nf * (2 ** i))

    def mask_nfilt(self, i):
        return min(self.mask_max, self.mask_nf * (2 ** i))

    def emb_ch(self, i):
        return min(self.emb_max, self.emb_nf * (2 ** i))

    # generator.py:270-287
    def cond_dim(self, i):
        return self.emb_ch(min(i, self.emb_down))

    def res_blocks(self):
        """[(name, cin, hid, cout, cond_level, learned_shortcut)] in forward order (generator.py:201-224)."""
        out = []
        for i in range(self.n_down + 1):
            ci, co = self.nfilt(i), self.nfilt(i + 1)
            out.append(('down_%d' % i, ci, min(ci, co), co, min(self.emb_down, i), ci != co))
        c = self.nfilt(self.n_down + 1)
        for i in range(self.n_res):
            out.append(('res_%d' % i, c, c, c, min(self.emb_down, self.n_down + 1), False))
        for i in range(self.n_down, -1, -1):
            ci, co = self.nfilt(i + 1), self.nfilt(i)
            out.append(('up_%d' % i, ci, min(ci, co), co, min(i, self.emb_down), ci != co))
        return out

    def state_spec(self):
        """[(key, shape, kind)] in the reference's state_dict order.

        kind: 'w' plain conv weight, 'b' bias, 'sn_w' weight_orig, 'sn_u', 'sn_v', 'in_w', 'in_b'.
        """
        spec = []

        def sn_conv(prefix, cout, cin, k):
            spec.append((prefix + '.bias', (cout,), 'b'))
            spec.append((prefix + '.weight_orig', (cout, cin, k, k), 'sn_w'))
            spec.append((prefix + '.weight_u', (cout,), 'sn_u'))
            spec.append((prefix + '.weight_v', (cin * k * k,), 'sn_v'))

        def plain_conv(prefix, cout, cin, k):
            spec.append((prefix + '.weight', (cout, cin, k, k), 'w'))
            spec.append((prefix + '.bias', (cout,), 'b'))

        def embedder(name, cin):
            sn_conv(name + '.conv_first.layers.conv', self.emb_ch(0), cin, 3)
            for i in range(self.emb_down):
                sn_conv('%s.down_%d.layers.conv' % (name, i), self.emb_ch(i + 1), self.emb_ch(i), 3)

        def spade_block(name, cin, hid, cout, cond, shortcut):
            for blk, (bi, bo, k) in (('conv_block_0', (cin, hid, 3)), ('conv_block_1', (hid, cout, 3)),
                                     ('conv_block_s', (cin, cout, 1))):
                if blk == 'conv_block_s' and not shortcut:
                    continue
                p = '%s.%s.layers' % (name, blk)
                plain_conv(p + '.norm.mlps.0.0.layers.conv', 2 * bi, cond, 1)
                sn_conv(p + '.conv', bo, bi, k)

        def cna(prefix, cout, cin, k):
            sn_conv(prefix + '.layers.conv', cout, cin, k)
            spec.append((prefix + '.layers.norm.weight', (cout,), 'in_w'))
            spec.append((prefix + '.layers.norm.bias', (cout,), 'in_b'))

        blocks = {b[0]: b for b in self.res_blocks()}
        embedder('ref_embedding', 2 * self.img_nc)
        embedder('label_embedding', self.label_nc)
        for i in range(self.n_down, -1, -1):
            n, ci, hid, co, lvl, sc = blocks['up_%d' % i]
            spade_block(n, ci, hid, co, self.emb_ch(lvl), sc)
        plain_conv('conv_img.layers.conv', self.img_nc, self.nf, 3)
        plain_conv('conv_mask.layers.conv', 1, self.nf, 3)
        plain_conv('down_first.layers.conv', self.nf, self.label_nc, 3)
        for i in range(self.n_down + 1):
            n, ci, hid, co, lvl, sc = blocks['down_%d' % i]
            spade_block(n, ci, hid, co, self.emb_ch(lvl), sc)
        for i in range(self.n_res):
            n, ci, hid, co, lvl, sc = blocks['res_%d' % i]
            spade_block(n, ci, hid, co, self.emb_ch(lvl), sc)
        # MaskGenerator (generator.py:423-491), stored as `flow_network_temp` (:175)
        f = 'flow_network_temp'
        for branch, cin in (('down_lbl', self.label_nc), ('down_img', 3 * self.img_nc)):
            cna('%s.%s.0' % (f, branch), self.mask_nf, cin, 3)
            for i in range(self.mask_down):
                cna('%s.%s.%d' % (f, branch, i + 1), self.mask_nfilt(i + 1), self.mask_nfilt(i), 3)
        ch = self.mask_nfilt(self.mask_down)
        for i in range(self.mask_res):
            ci = 2 * ch if i == 0 else ch
            cna('%s.res_flow.%d.conv_block_0' % (f, i), ch, ci, 3)
            cna('%s.res_flow.%d.conv_block_1' % (f, i), ch, ch, 3)
            if i == 0:
                cna('%s.res_flow.%d.conv_block_s' % (f, i), ch, ci, 1)
        for n, i in enumerate(reversed(range(self.mask_down))):
            cna('%s.up_flow.%d' % (f, 2 * n + 1), self.mask_nfilt(i), self.mask_nfilt(i + 1), 3)
        plain_conv('%s.conv_mask.0.layers.conv' % f, 1, self.mask_nf, 3)
        return spec
