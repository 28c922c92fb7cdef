"""The consumer of the MuPS tensor on the B200 tensor cores.

``TensorCoreExperts`` runs the forward pass of the Mixture-of-Experts normal estimator (models/experts_n_est.py:78-108,
155-314; restated in ``experts_net.ExpertsNormalEstimator``) with every convolution and fully connected layer on the
hand-written tcgen05 / TMEM / TMA implicit-GEMM kernel of ``csrc/moe_conv.cu`` (``mups_conv3d_bn_relu``): bf16 operands,
fp32 accumulation in tensor memory, bias + batch norm folded into a per-channel scale / shift and fused with the ReLU in
the epilogue, every branch of an inception module writing straight into its channel slice of the module's output (the
concat is free), activations kept NDHWC bf16 on the device between layers.  MuPS goes in as the fp32 tensor the
statistics kernel wrote and never leaves the GPU.  The TF 'SAME' average pools and the 2/2 max pools are one small bf16
NDHWC kernel (``mups_pool3d``); torch is left with the softmax over seven gate outputs.

Precision: bf16 products, fp32 sums.  The normals stay within a few tenths of a degree of the fp32 network
(tests/test_gpu.py::test_tensor_core_consumer_against_fp32_network states the measured deviation); that is two orders of
magnitude below the 5 / 10 degree thresholds of the reference's own metrics (utils/evaluate.py: PGP5, PGP10).

``TensorCoreExperts(model, precision="bf16x3")`` is the fp32-grade mode for callers who want the reference's fp32 numbers from
the tensor cores: every value is carried as two bf16 numbers (hi, lo), a tensor of w logical channels is stored as the
triplet [hi | lo | hi], the weights are expanded to [w_hi | w_hi | w_lo], and the SAME convolution kernel over the three
times longer channel axis accumulates a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32 (csrc/moe_split.cu).  Three times the
tensor-core work, 16 significant bits per operand instead of 8.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .experts_net import ExpertsNormalEstimator  # noqa: F401


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _pad8(c):
    return (int(c) + 7) // 8 * 8


def _pad16(c):
    return (int(c) + 15) // 16 * 16


class PackedConv(object):
    """One conv3d / fully connected layer of the reference with its batch norm (evaluated with the moving statistics,
    eps 1e-3) and bias folded: weights [k^3][Cout_pad][Cin_pad] bf16 (tap order dz, dy, dx; TF cross-correlation), scale /
    shift [Cout_pad] fp32."""

    def __init__(self, weight, bias, bn, relu, k, device, in_map=None, cin_pad=None, x3_segs=None):
        """weight: [Cout, Cin, k, k, k] (conv) or [Cout, Cin] (linear); in_map: for each real input channel its index in
        the padded input layout (None: identity).  x3_segs (bf16x3 mode): the widths of the input layout's segments, each of
        which is stored as a triplet [hi | lo | hi]; the weights' inner axis becomes [w_hi | w_hi | w_lo] per segment."""
        w = weight.detach().float().cpu()
        if w.ndim == 2:
            w = w[:, :, None, None, None]
        cout, cin = int(w.shape[0]), int(w.shape[1])
        self.k, self.relu, self.cout = int(k), bool(relu), cout
        self.cout_pad = _pad16(cout)
        self.cin_pad = _pad8(cin) if cin_pad is None else int(cin_pad)
        wp = torch.zeros((k ** 3, self.cout_pad, self.cin_pad), dtype=torch.float32)
        idx = torch.arange(cin) if in_map is None else torch.as_tensor(in_map, dtype=torch.long)
        wp[:, :cout, idx] = w.permute(2, 3, 4, 0, 1).reshape(k ** 3, cout, cin)
        b = bias.detach().float().cpu() if bias is not None else torch.zeros(cout)
        if bn is not None:
            s = bn.weight.detach().float().cpu() / torch.sqrt(bn.running_var.detach().float().cpu() + bn.eps)
            t = (b - bn.running_mean.detach().float().cpu()) * s + bn.bias.detach().float().cpu()
        else:
            s, t = torch.ones(cout), b
        scale, shift = torch.zeros(self.cout_pad), torch.zeros(self.cout_pad)
        scale[:cout], shift[:cout] = s, t
        if x3_segs is not None:
            if sum(x3_segs) != self.cin_pad:
                raise ValueError("bf16x3 segments %r do not cover the %d input channels" % (list(x3_segs), self.cin_pad))
            w_hi = wp.to(torch.bfloat16).float()
            w_lo = wp - w_hi                                  # exact in fp32; rounded to bf16 below
            parts, s0 = [], 0
            for wd in x3_segs:
                parts += [w_hi[..., s0:s0 + wd], w_hi[..., s0:s0 + wd], w_lo[..., s0:s0 + wd]]
                s0 += wd
            wp = torch.cat(parts, dim=-1)
            self.cin_pad *= 3
        self.w = wp.to(device=device, dtype=torch.bfloat16).contiguous()
        self.scale = scale.to(device).contiguous()
        self.shift = shift.to(device).contiguous()


def conv3d_bn_relu(x, cin_off, cin, layer, out=None, cout_off=0, out_f32=None):
    """x: NDHWC bf16 [B, D, D, D, Ct] (or [B, Ct] for a fully connected layer); reads channels [cin_off, cin_off + cin).
    Writes layer.cout_pad channels at ``cout_off`` of ``out`` (NDHWC bf16) and / or the fp32 tensor ``out_f32`` [rows, Cout_pad]."""
    B = int(x.shape[0])
    D = int(x.shape[1]) if x.ndim == 5 else 1
    ct = int(x.shape[-1])
    if out is None and out_f32 is None:
        out = torch.empty(tuple(x.shape[:-1]) + (layer.cout_pad,), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mups_conv3d_bn_relu(
            _ptr(x), B, D, ct, int(cin_off), int(cin), _ptr(layer.w), layer.cin_pad, layer.cout_pad, layer.k, _ptr(layer.scale),
            _ptr(layer.shift), 1 if layer.relu else 0, _ptr(out), int(out.shape[-1]) if out is not None else 0, int(cout_off),
            _ptr(out_f32), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "mups_conv3d_bn_relu")
    return out if out is not None else out_f32


def conv1_split(x, cin_off, cin, layer, split, relu_upto, out1, off1, out2, off2):
    """``layer``: two 1^3 convolutions stacked along the output channels (rows [0, split) / [split, cout_pad)), one read of x;
    the first goes to channels [off1, ...) of out1, the second to [off2, ...) of out2 (mups_conv1_split_bn_relu)."""
    B = int(x.shape[0])
    D = int(x.shape[1]) if x.ndim == 5 else 1
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mups_conv1_split_bn_relu(
            _ptr(x), B, D, int(x.shape[-1]), int(cin_off), int(cin), _ptr(layer.w), layer.cin_pad, layer.cout_pad, _ptr(layer.scale),
            _ptr(layer.shift), int(relu_upto), _ptr(out1), int(out1.shape[-1]), int(off1), int(split), _ptr(out2), int(out2.shape[-1]),
            int(off2), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "mups_conv1_split_bn_relu")


def avgpool_bn_relu(x, c_off, c, k, scale, shift, relu, out, y_off):
    """Average pool (TF 'SAME', window k) of channels [c_off, c_off + c) of x, then act(scale * . + shift) into channels
    [y_off, y_off + c) of out (mups_avgpool3d_bn_relu)."""
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mups_avgpool3d_bn_relu(
            _ptr(x), int(x.shape[0]), int(x.shape[1]), int(x.shape[-1]), int(c_off), int(c), int(k), _ptr(scale), _ptr(shift),
            1 if relu else 0, _ptr(out), int(out.shape[-1]), int(y_off),
            ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "mups_avgpool3d_bn_relu")


def pool3d(x, c_off, c, k, is_max):
    """tf_util.avg_pool3d (window k, stride 1, 'SAME': mean over the valid cells) or max_pool3d (2, stride 2) on channels
    [c_off, c_off + c) of the NDHWC bf16 tensor x (mups_pool3d) -> contiguous NDHWC bf16."""
    B, D = int(x.shape[0]), int(x.shape[1])
    Do = D // 2 if is_max else D
    y = torch.empty((B, Do, Do, Do, int(c)), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mups_pool3d(_ptr(x), B, D, int(x.shape[-1]), int(c_off), int(c), int(k), 1 if is_max else 0, _ptr(y),
                                           ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "mups_pool3d")
    return y


def conv3d_f32(x, cin_off, cin, layer):
    """The same convolution with its fp32 output only: [rows, Cout_pad] (scale / shift / ReLU applied)."""
    rows = x.numel() // int(x.shape[-1])
    out = torch.empty((rows, layer.cout_pad), dtype=torch.float32, device=x.device)
    conv3d_bn_relu(x, cin_off, cin, layer, None, 0, out)
    return out


def conv3d_x3(x, cin_off, cin, layer, out, cout_off, split=None):
    """The convolution with its output written as triplets by the epilogue (mups_conv3d_bn_relu_x3): output channels [0, split)
    -> the triplet of part width ``split`` at channels [cout_off, ...) of ``out``, channels [split, cout_pad) -> a second triplet
    right behind it (default: one triplet of part width cout_pad).  cin_off, cin, cout_off: PHYSICAL channels."""
    B = int(x.shape[0])
    D = int(x.shape[1]) if x.ndim == 5 else 1
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mups_conv3d_bn_relu_x3(
            _ptr(x), B, D, int(x.shape[-1]), int(cin_off), int(cin), _ptr(layer.w), layer.cin_pad, layer.cout_pad, layer.k, _ptr(layer.scale),
            _ptr(layer.shift), 1 if layer.relu else 0, _ptr(out), int(out.shape[-1]), int(cout_off),
            int(layer.cout_pad if split is None else split), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
            "mups_conv3d_bn_relu_x3")


def split_x3(src, src_off, w_src, dst, dst_off, w_dst):
    """fp32 src [rows, n] columns [src_off, src_off + w_src) -> the triplet [hi | lo | hi] at channels [dst_off, dst_off + 3 w_dst)
    of the bf16 tensor dst [..., Ct] (mups_split_bf16x3)."""
    rows = dst.numel() // int(dst.shape[-1])
    with torch.cuda.device(dst.device):
        _lib.check(_lib.load().mups_split_bf16x3(_ptr(src), rows, int(src.shape[-1]), int(src_off), int(w_src), _ptr(dst), int(dst.shape[-1]),
                                                 int(dst_off), int(w_dst), ctypes.c_void_p(torch.cuda.current_stream(dst.device).cuda_stream)),
                   "mups_split_bf16x3")


def pool3d_x3(x, c_off, w, k, is_max, y, y_off):
    """Average / max pool (as pool3d) of the triplet at channels [c_off, c_off + 3 w) of x into the triplet at y_off of y."""
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mups_pool3d_bf16x3(_ptr(x), int(x.shape[0]), int(x.shape[1]), int(x.shape[-1]), int(c_off), int(w), int(k),
                                                  1 if is_max else 0, _ptr(y), int(y.shape[-1]), int(y_off),
                                                  ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "mups_pool3d_bf16x3")


def avgpool_f32_x3(src, B, D, c, k, scale, shift, relu, out, y_off):
    """fp32 src [B * D^3, c] (a convolution's raw output) -> TF 'SAME' average pool (window k) -> act(scale * . + shift) -> the
    triplet at channels [y_off, y_off + 3 c) of out (mups_avgpool3d_f32_bn_relu_x3)."""
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().mups_avgpool3d_f32_bn_relu_x3(_ptr(src), int(B), int(D), int(c), int(k), _ptr(scale), _ptr(shift),
                                                             1 if relu else 0, _ptr(out), int(out.shape[-1]), int(y_off),
                                                             ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)),
                   "mups_avgpool3d_f32_bn_relu_x3")


def _pool_segments_x3(x, c_off, segs, k, is_max):
    """Pool the logical channels [c_off, c_off + sum(segs)) of a triplet tensor segment by segment -> a new triplet tensor with
    the same segments."""
    B, D = int(x.shape[0]), int(x.shape[1])
    Do = D // 2 if is_max else D
    y = torch.empty((B, Do, Do, Do, 3 * sum(segs)), dtype=torch.bfloat16, device=x.device)
    s0 = 0
    for wd in segs:
        pool3d_x3(x, 3 * (c_off + s0), wd, k, is_max, y, 3 * s0)
        s0 += wd
    return y


def _fuse_branches(a, b):
    """The two convolutions of an inception module that read the same tensor (kernel edges k_a < k_b, 'SAME') as ONE
    convolution with b's kernel: a's taps sit in the centre of the larger window (offset = difference of the paddings before),
    zeros elsewhere; output channels [a | b] -- which is how they lie in the module's output anyway.  Worth it when the
    fused tile is at most 128 channels wide: a 64-channel tile runs the tensor cores at two thirds of their rate (operand
    reads from shared memory per FLOP), a 128-channel one at the full rate, and a launch is saved."""
    off = (b.k - 1) // 2 - (a.k - 1) // 2
    assert a.cin_pad == b.cin_pad and a.k < b.k and off >= 0 and off + a.k <= b.k and a.relu == b.relu
    f = PackedConv.__new__(PackedConv)
    f.k, f.relu, f.cin_pad = b.k, b.relu, b.cin_pad
    f.cout = f.cout_pad = a.cout_pad + b.cout_pad
    w = torch.zeros((b.k ** 3, f.cout_pad, f.cin_pad), dtype=torch.bfloat16, device=b.w.device)
    w[:, a.cout_pad:, :] = b.w
    for dz in range(a.k):
        for dy in range(a.k):
            for dx in range(a.k):
                w[((dz + off) * b.k + dy + off) * b.k + dx + off, :a.cout_pad, :] = a.w[(dz * a.k + dy) * a.k + dx]
    f.w = w.contiguous()
    f.scale, f.shift = torch.cat([a.scale, b.scale]).contiguous(), torch.cat([a.shift, b.shift]).contiguous()
    return f


FUSE_POOL_BRANCH = True      # A/B switch (profiles/bench_moe.py): the pool branch's convolution computed with `one`, pooled afterwards


def _stack_one_and_pool(one, pool, pool_after):
    """`one` and the pool branch's 1^3 convolution as one layer (output channels [one | pool]).  pool_after: the average pool
    runs AFTER this convolution (they commute: both linear, the bias is a constant the valid-cell mean keeps), so the pool half
    leaves here raw (scale 1, shift 0, no ReLU) and its folded bias + batch norm + ReLU are applied by the pooling kernel."""
    f = PackedConv.__new__(PackedConv)
    f.k, f.relu, f.cin_pad = 1, True, one.cin_pad
    f.cout = f.cout_pad = one.cout_pad + pool.cout_pad
    f.w = torch.cat([one.w, pool.w], dim=1).contiguous()
    if pool_after:
        ident = torch.zeros_like(pool.scale)
        ident[:pool.cout] = 1.0
        f.scale, f.shift = torch.cat([one.scale, ident]).contiguous(), torch.cat([one.shift, torch.zeros_like(pool.shift)]).contiguous()
    else:
        f.scale, f.shift = torch.cat([one.scale, pool.scale]).contiguous(), torch.cat([one.shift, pool.shift]).contiguous()
    return f


class _PackedInception(object):
    def __init__(self, m, device, in_map=None, cin_pad=None):
        self.k0 = m.k0
        mk = lambda c, im=None, cp=None: PackedConv(c.conv.weight, c.conv.bias, c.bn, True, c.conv.kernel_size[0], device, im, cp)
        self.one, self.pool = mk(m.one, in_map, cin_pad), mk(m.pool, in_map, cin_pad)
        self.a, self.b = mk(m.a), mk(m.b)
        self.nf = self.one.cout_pad
        self.ab = _fuse_branches(self.a, self.b) if (self.a.cout_pad + self.b.cout_pad <= 128 and self.a.k < self.b.k) else None
        self.one_pool = _stack_one_and_pool(self.one, self.pool, self.k0 > 1)
        self.c_out = 2 * self.nf + self.a.cout_pad + self.b.cout_pad
        # real-channel positions inside this module's (padded) output, for the next layer's weight packing
        real = lambda l, off: [off + i for i in range(l.cout)]
        self.out_map = (real(self.one, 0) + real(self.a, self.nf) + real(self.b, self.nf + self.a.cout_pad)
                        + real(self.pool, self.nf + self.a.cout_pad + self.b.cout_pad))

    def __call__(self, x, cin_off, cin):
        out = torch.empty(tuple(x.shape[:-1]) + (self.c_out,), dtype=torch.bfloat16, device=x.device)
        off = self.nf + self.a.cout_pad + self.b.cout_pad
        fused = FUSE_POOL_BRANCH and self.nf % 16 == 0
        if fused and self.k0 == 1:               # a 1-wide average pool is the identity: both convolutions, one read of x
            conv1_split(x, cin_off, cin, self.one_pool, self.nf, 2 * self.nf, out, 0, out, off)
        elif fused:
            pre = torch.empty(tuple(x.shape[:-1]) + (self.nf,), dtype=torch.bfloat16, device=x.device)
            conv1_split(x, cin_off, cin, self.one_pool, self.nf, self.nf, out, 0, pre, 0)
        else:
            conv3d_bn_relu(x, cin_off, cin, self.one, out, 0)
        if self.ab is not None:
            conv3d_bn_relu(out, 0, self.nf, self.ab, out, self.nf)
        else:
            conv3d_bn_relu(out, 0, self.nf, self.a, out, self.nf)
            conv3d_bn_relu(out, 0, self.nf, self.b, out, self.nf + self.a.cout_pad)
        if fused and self.k0 > 1:
            avgpool_bn_relu(pre, 0, self.nf, self.k0, self.pool.scale, self.pool.shift, True, out, off)
        elif not fused and self.k0 == 1:        # a 1-wide average pool is the identity
            conv3d_bn_relu(x, cin_off, cin, self.pool, out, off)
        elif not fused:
            conv3d_bn_relu(pool3d(x, cin_off, cin, self.k0, False), 0, cin, self.pool, out, off)
        return out


class _PackedInceptionX3(object):
    """The inception module in bf16x3 mode: the module's output holds the triplets of its four branches
    [one | a | b | pool] one after the other, written by the convolutions' own epilogues (conv3d_x3)."""

    def __init__(self, m, device, in_map, cin_pad, in_segs):
        self.k0, self.in_segs = m.k0, list(in_segs)
        mk = lambda c, im, cp, sg: PackedConv(c.conv.weight, c.conv.bias, c.bn, True, c.conv.kernel_size[0], device, im, cp, sg)
        self.one, self.pool = mk(m.one, in_map, cin_pad, in_segs), mk(m.pool, in_map, cin_pad, in_segs)
        self.nf = self.one.cout_pad
        self.a, self.b = mk(m.a, None, self.nf, [self.nf]), mk(m.b, None, self.nf, [self.nf])
        self.ab = _fuse_branches(self.a, self.b) if (self.a.cout_pad + self.b.cout_pad <= 128 and self.a.k < self.b.k) else None
        if self.k0 > 1:
            # the pool branch's 1^3 convolution commutes with the average pool: it runs FIRST, raw (no bias, no batch norm, no
            # ReLU), on n_filters instead of C_in channels; the pool applies the folded bias + batch norm + ReLU (as in the bf16 mode)
            raw = PackedConv.__new__(PackedConv)
            raw.__dict__.update(self.pool.__dict__)
            raw.relu = False
            raw.scale, raw.shift = torch.zeros_like(self.pool.scale), torch.zeros_like(self.pool.shift)
            raw.scale[:self.pool.cout] = 1.0
            self.pool_raw = raw
        self.c_out = 2 * self.nf + self.a.cout_pad + self.b.cout_pad
        self.out_segs = [self.nf, self.a.cout_pad, self.b.cout_pad, self.nf]
        real = lambda l, off: [off + i for i in range(l.cout)]
        self.out_map = (real(self.one, 0) + real(self.a, self.nf) + real(self.b, self.nf + self.a.cout_pad)
                        + real(self.pool, self.nf + self.a.cout_pad + self.b.cout_pad))

    def __call__(self, x, cin_off, cin):
        """x: triplet tensor; cin_off, cin: LOGICAL channels (whole segments)."""
        out = torch.empty(tuple(x.shape[:-1]) + (3 * self.c_out,), dtype=torch.bfloat16, device=x.device)
        nf, ap, bp = self.nf, self.a.cout_pad, self.b.cout_pad
        conv3d_x3(x, 3 * cin_off, 3 * cin, self.one, out, 0)
        if self.ab is not None:                  # output channels [a | b] -> the two triplets [a | b] of the module's output
            conv3d_x3(out, 0, 3 * nf, self.ab, out, 3 * nf, ap)
        else:
            conv3d_x3(out, 0, 3 * nf, self.a, out, 3 * nf)
            conv3d_x3(out, 0, 3 * nf, self.b, out, 3 * (nf + ap))
        if self.k0 == 1:                         # a 1-wide average pool is the identity
            conv3d_x3(x, 3 * cin_off, 3 * cin, self.pool, out, 3 * (nf + ap + bp))
        else:
            t = conv3d_f32(x, 3 * cin_off, 3 * cin, self.pool_raw)
            avgpool_f32_x3(t, x.shape[0], x.shape[1], nf, self.k0, self.pool.scale, self.pool.shift, True, out, 3 * (nf + ap + bp))
        return out


class _PackedConvNet(object):
    def __init__(self, net, device, in_map, cin_pad, x3_segs=None):
        """x3_segs: None (bf16 activations) or the segment widths of the input slice (bf16x3 mode, triplet tensors)."""
        self.steps, self.x3 = [], x3_segs is not None
        it = iter(net.mods)
        cur_map, cur_pad, cur_segs = in_map, cin_pad, (list(x3_segs) if self.x3 else None)
        for step in net.plan:
            if step is None:
                inc = (_PackedInceptionX3(next(it), device, cur_map, cur_pad, cur_segs) if self.x3
                       else _PackedInception(next(it), device, cur_map, cur_pad))
                self.steps.append(inc)
                cur_map, cur_pad = inc.out_map, inc.c_out
                cur_segs = inc.out_segs if self.x3 else None
            else:
                self.steps.append((step, cur_segs))
        self.out_map, self.out_pad, self.out_segs = cur_map, cur_pad, cur_segs

    def __call__(self, x, cin_off, cin):
        """cin_off, cin: channels of x the first module reads (LOGICAL channels in bf16x3 mode)."""
        for step in self.steps:
            if isinstance(step, (_PackedInception, _PackedInceptionX3)):
                x = step(x, cin_off, cin)
                cin_off, cin = 0, step.c_out
            else:
                (step, segs) = step
                if (step[1], step[2]) != (2, 2) or x.shape[1] % 2:
                    raise ValueError("only the 2-wide, stride-2 max pools of the reference's 8^3 networks are packed")
                x = _pool_segments_x3(x, 0, segs, 2, True) if self.x3 else pool3d(x, 0, int(x.shape[-1]), 2, True)
        if x.shape[1] != 1:
            raise ValueError("only the 8^3 networks of the reference (which end at a 1^3 volume) are packed for the tensor cores")
        return x.reshape(x.shape[0], -1)       # TF flattens channels-last [B, d, h, w, C]


class TensorCoreExperts(object):
    """``ExpertsNormalEstimator`` packed for the tcgen05 kernel.  ``predict(mups)`` takes the fp32 MuPS tensor
    [B, res, res, res, 20 S] on the GPU and returns (normal of the most probable expert [B, 3], expert [B],
    probabilities [B, n_experts]) like ``ExpertsNormalEstimator.predict``."""

    def __init__(self, model, device=None, precision="bf16"):
        if not torch.cuda.is_available():
            raise RuntimeError("TensorCoreExperts needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3', not %r" % (precision,))
        self.precision, self.x3 = precision, precision == "bf16x3"
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        model = model.eval()
        self.expert_dict = model.expert_dict
        first = model.gate_conv.mods[0].one.conv
        self.n_rads = int(first.in_channels) // 20
        S = self.n_rads
        scale_map = lambda scales: [32 * i + c for i in range(len(scales)) for c in range(20)]      # 20 real channels per 32-wide scale slot
        segs = lambda scales: [32] * len(scales) if self.x3 else None                               # bf16x3: one triplet per scale slot
        self.gate = _PackedConvNet(model.gate_conv, self.device, scale_map(range(S)), 32 * S, segs(range(S)))
        self.gate_fc = self._pack_fc(model.gate_fc, self.gate)
        self.experts = []
        for i, (conv, fc) in enumerate(zip(model.expert_conv, model.expert_fc)):
            scales = self.expert_dict[i]
            net = _PackedConvNet(conv, self.device, scale_map(scales), 32 * len(scales), segs(scales))
            self.experts.append((int(np.min(scales)) * 32, 32 * len(scales), net, self._pack_fc(fc, net)))
        self.n_experts = len(self.experts)

    def _pack_fc(self, seq, net):
        layers, in_map, pad, sg = [], net.out_map, net.out_pad, net.out_segs
        for fc in seq:
            l = PackedConv(fc.lin.weight, fc.lin.bias, fc.bn, fc.relu, 1, self.device, in_map, pad, sg)
            layers.append(l)
            in_map, pad = None, l.cout_pad
            sg = [l.cout_pad] if self.x3 else None
        return layers

    def _run_fc(self, layers, x):
        for l in layers[:-1]:
            if self.x3:
                y = torch.empty((x.shape[0], 3 * l.cout_pad), dtype=torch.bfloat16, device=x.device)
                conv3d_x3(x, 0, int(x.shape[-1]), l, y, 0)
                x = y
            else:
                x = conv3d_bn_relu(x, 0, int(x.shape[-1]), l)
        last = layers[-1]
        out = torch.empty((x.shape[0], last.cout_pad), dtype=torch.float32, device=x.device)
        conv3d_bn_relu(x, 0, int(x.shape[-1]), last, None, 0, out)
        return out[:, :last.cout]

    def pack_input(self, mups):
        """fp32 MuPS [B, res, res, res, 20 S] -> bf16 [B, res, res, res, 32 S] (every scale padded to 32 channels); in bf16x3
        mode [B, res, res, res, 96 S]: one triplet [hi | lo | hi] of 32-wide parts per scale."""
        B, res = int(mups.shape[0]), int(mups.shape[1])
        S = self.n_rads
        if self.x3:
            x = torch.empty((B, res, res, res, 96 * S), dtype=torch.bfloat16, device=mups.device)
            src = mups.reshape(B * res ** 3, 20 * S)
            for s in range(S):
                split_x3(src, 20 * s, 20, x, 96 * s, 32)
            return x
        x = torch.empty((B, res, res, res, 32 * S), dtype=torch.bfloat16, device=mups.device)
        with torch.cuda.device(mups.device):
            _lib.check(_lib.load().mups_moe_pack_input(_ptr(mups), B * res ** 3, S, _ptr(x),
                                                       ctypes.c_void_p(torch.cuda.current_stream(mups.device).cuda_stream)),
                       "mups_moe_pack_input")
        return x

    @torch.no_grad()
    def forward(self, mups):
        """(experts_prob [n_experts, B], n_est [n_experts, B, 3]) like ExpertsNormalEstimator.forward."""
        mups = mups.to(self.device, torch.float32).contiguous()
        x = self.pack_input(mups)
        logits = self._run_fc(self.gate_fc, self.gate(x, 0, 32 * self.n_rads))
        prob = F.softmax(logits, dim=1).transpose(0, 1)
        normals = [self._run_fc(fc, net(x, off, width)) for off, width, net, fc in self.experts]
        return prob, torch.stack(normals)

    @torch.no_grad()
    def predict(self, mups):
        prob, n_est = self.forward(mups)
        expert = prob.argmax(dim=0)
        return n_est[expert, torch.arange(n_est.shape[1], device=n_est.device)], expert, prob.transpose(0, 1)
