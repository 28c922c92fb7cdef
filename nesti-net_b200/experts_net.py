"""Mixture-of-Experts normal estimator of Nesti-Net as a PyTorch module (forward only, random-init).

SURVEY.md section 8(f) rank 1 ("next" row): the consumer of the MuPS tensor, restated from
models/experts_n_est.py:78-108,155-314 with the TF1 layer semantics of utils/tf_util.py
(conv3d :254-311, fully_connected :314-351, max/avg_pool3d :406-455, batch norm :458-495).  It
exists for the fourth correctness gate of BASELINE.json ("downstream MoE normals within 1e-4 angular
RMS": the same randomly initialised network evaluated on oracle MuPS and on GPU MuPS) and as the
head of an inference driver; it is library code (torch.nn / cuDNN), not a hand-written kernel.

TF semantics reproduced: 'SAME' padding is asymmetric for even kernels (extra cell after),
avg_pool3d 'SAME' divides by the number of valid cells, batch norm epsilon 1e-3 (evaluated with
the moving statistics), the gate ends in ReLU -> softmax (:174-177), an expert sees the channel
slice MuPS[..., 20*min(scales) : 20*min(scales) + 20*len(scales)] (:99-103) and its first inception
module is np.round(128 / divider) wide (:254) -- under the reference's Python 2.7 that `/` is an INTEGER division
(42 for a three-scale expert, not 43).

The architecture is pinned by the reference's own text: tests/golden/moe_tf_emulated.npz holds the outputs of
get_model's network statements, scale_manager_net, conv_net_8g / _3g, normal_est_net and inception_module executed (with
the tf_util layers) on a numpy emulation of the primitive TF ops; ``load_tf_variables`` restores the same variables by
their TensorFlow names (tests/test_host.py::test_experts_net_against_reference_text_on_emulated_tf).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _same_pad(k):
    """TF 'SAME', stride 1: total k-1, the smaller half first."""
    total = k - 1
    return total // 2, total - total // 2


class Conv3dSame(nn.Module):
    """tf_util.conv3d with bn=True: conv ('SAME') + bias + batch norm (eps 1e-3) + ReLU, channels-first."""

    def __init__(self, c_in, c_out, k):
        super().__init__()
        self.pad = _same_pad(k)
        self.conv = nn.Conv3d(c_in, c_out, k, bias=True)
        self.bn = nn.BatchNorm3d(c_out, eps=1e-3)
        nn.init.xavier_uniform_(self.conv.weight)
        nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        a, b = self.pad
        if a or b:
            x = F.pad(x, (a, b, a, b, a, b))
        return F.relu(self.bn(self.conv(x)))


def avg_pool_same(x, k):
    """tf.nn.avg_pool3d(padding='SAME', stride 1): mean over the valid cells of each window."""
    if k == 1:
        return x
    a, b = _same_pad(k)
    xs = F.avg_pool3d(F.pad(x, (a, b, a, b, a, b)), k, stride=1) * float(k ** 3)
    ones = torch.ones((1, 1) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    cnt = F.avg_pool3d(F.pad(ones, (a, b, a, b, a, b)), k, stride=1) * float(k ** 3)
    return xs / cnt


def max_pool_same(x, k, stride):
    """tf.nn.max_pool3d(padding='SAME'): pads with -inf, the smaller half first."""
    n = x.shape[-1]
    out = -(-n // stride)
    total = max((out - 1) * stride + k - n, 0)
    a, b = total // 2, total - total // 2
    if total:
        x = F.pad(x, (a, b, a, b, a, b), value=float("-inf"))
    return F.max_pool3d(x, k, stride=stride)


class Inception3d(nn.Module):
    """models/experts_n_est.py:291-310: 1x1x1 | k0 conv of it | k1 conv of it | avg-pool(k0) -> 1x1x1, concatenated."""

    def __init__(self, c_in, n_filters, kernel_sizes):
        super().__init__()
        n_filters = int(n_filters)
        self.k0 = kernel_sizes[0]
        self.one = Conv3dSame(c_in, n_filters, 1)
        self.a = Conv3dSame(n_filters, n_filters // 2, kernel_sizes[0])
        self.b = Conv3dSame(n_filters, n_filters // 2, kernel_sizes[1])
        self.pool = Conv3dSame(c_in, n_filters, 1)
        self.c_out = 2 * n_filters + 2 * (n_filters // 2)

    def forward(self, x):
        one = self.one(x)
        return torch.cat([one, self.a(one), self.b(one), self.pool(avg_pool_same(x, self.k0))], dim=1)


class FC(nn.Module):
    """tf_util.fully_connected: linear + (batch norm) + activation."""

    def __init__(self, c_in, c_out, bn=True, relu=True):
        super().__init__()
        self.lin = nn.Linear(c_in, c_out)
        self.bn = nn.BatchNorm1d(c_out, eps=1e-3) if bn else None
        self.relu = relu
        nn.init.xavier_uniform_(self.lin.weight)
        nn.init.zeros_(self.lin.bias)

    def forward(self, x):
        x = self.lin(x)
        if self.bn is not None:
            x = self.bn(x)
        return F.relu(x) if self.relu else x


class _ConvNet(nn.Module):
    """conv_net_8g (:181-215), its lighter expert version (:253-275) and conv_net_3g (:217-241)."""

    def __init__(self, c_in, res, first_width, expert):
        super().__init__()
        layers = []

        def inc(c, n, ks):
            m = Inception3d(c, n, ks)
            layers.append(m)
            return m.c_out
        c = c_in
        if res == 8 and not expert:
            c = inc(c, 128, [3, 5]); c = inc(c, 256, [3, 5]); c = inc(c, 256, [3, 5]); layers.append(("max", 2, 2))
            c = inc(c, 512, [2, 4]); c = inc(c, 512, [2, 4]); layers.append(("max", 2, 2))
            c = inc(c, 512, [1, 2]); layers.append(("max", 2, 2))
            side = 1
        elif res == 8:
            c = inc(c, first_width, [3, 5]); c = inc(c, 256, [3, 5]); layers.append(("max", 2, 2))
            c = inc(c, 256, [2, 4]); layers.append(("max", 2, 2))
            c = inc(c, 512, [2, 4]); layers.append(("max", 2, 2))
            side = 1
        elif res == 3:
            c = inc(c, 128, [2, 3]); c = inc(c, 256, [2, 3]); c = inc(c, 256, [1, 2]); c = inc(c, 512, [1, 2])
            layers.append(("max", 3, 2))
            side = 2
        else:
            raise ValueError('Incompatible number of Gaussians - currently 3 and 8 are supported. '
                             'For other values you should tweak the architecture')
        self.mods = nn.ModuleList([m for m in layers if isinstance(m, nn.Module)])
        self.plan = [m if not isinstance(m, nn.Module) else None for m in layers]
        self.out_features = c * side ** 3

    def forward(self, x):
        it = iter(self.mods)
        for step in self.plan:
            x = next(it)(x) if step is None else max_pool_same(x, step[1], step[2])
        # TF flattens channels-last [B, d, h, w, C]
        return x.permute(0, 2, 3, 4, 1).reshape(x.shape[0], -1)


class ExpertsNormalEstimator(nn.Module):
    """get_model of models/experts_n_est.py minus the MuPS front end: MuPS [B,res,res,res,20*S] ->
    (experts_prob [n_experts, B], n_est [n_experts, B, 3])."""

    def __init__(self, n_rads, n_gaussians, n_experts=7, expert_dict=None):
        super().__init__()
        res = int(np.round(np.power(n_gaussians, 1.0 / 3.0)))
        if expert_dict is None:                      # :82-95 default assignment
            assignment = []
            for i in range(n_rads):
                assignment += [[i]] * (n_experts // n_rads)
            assignment += [list(range(n_rads))] * (n_experts % n_rads)
            expert_dict = {i: assignment[i] for i in range(n_experts)}
        elif n_experts != len(expert_dict):
            raise ValueError('Incompatible expert assignment values in variable expert_dict ')
        self.expert_dict = {int(k): list(v) for k, v in expert_dict.items()}
        self.gate_conv = _ConvNet(20 * n_rads, res, 128, expert=False)
        self.gate_fc = nn.Sequential(FC(self.gate_conv.out_features, 1024), FC(1024, 256), FC(256, 128),
                                     FC(128, n_experts, bn=False, relu=True))
        self.expert_conv = nn.ModuleList()
        self.expert_fc = nn.ModuleList()
        for i in range(n_experts):
            scales = self.expert_dict[i]
            conv = _ConvNet(20 * len(scales), res, int(np.round(128 // len(scales))), expert=True)   # py2 `/` (:254)
            self.expert_conv.append(conv)
            self.expert_fc.append(nn.Sequential(FC(conv.out_features, 512), FC(512, 128), FC(128, 64),
                                                FC(64, 3, bn=False, relu=False)))

    def forward(self, mups):
        x = mups.permute(0, 4, 1, 2, 3).contiguous()             # channels-first for torch
        prob = F.softmax(self.gate_fc(self.gate_conv(x)), dim=1).transpose(0, 1)      # [n_experts, B] (:176-178)
        normals = []
        for i, (conv, fc) in enumerate(zip(self.expert_conv, self.expert_fc)):
            scales = self.expert_dict[i]
            start = int(np.min(scales)) * 20                      # :99-100
            normals.append(fc(conv(x[:, start:start + 20 * len(scales)])))
        return prob, torch.stack(normals)

    def tf_variables(self):
        """[(TensorFlow variable name, parameter / buffer, layout)] in the reference graph's naming
        (tf_util.conv3d / fully_connected scopes under the `scope=` strings of models/experts_n_est.py:169-176,186-212,
        254-285,296-307): '<scope>/weights', '/biases', '/bn/beta', '/bn/gamma', '/bn/moving_mean', '/bn/moving_variance'.
        layout 'conv': TF [kd, kh, kw, Cin, Cout] -> torch [Cout, Cin, kd, kh, kw]; 'fc': TF [in, out] -> torch [out, in]."""
        out = []

        def bn(scope, m):
            if m is not None:
                out.extend([(scope + "/bn/beta", m.bias, None), (scope + "/bn/gamma", m.weight, None),
                            (scope + "/bn/moving_mean", m.running_mean, None),
                            (scope + "/bn/moving_variance", m.running_var, None)])

        def convnet(net, scope_str):
            it = iter(net.mods)
            for layer, step in enumerate(net.plan, start=1):          # pools take a layer number too (:186-212)
                if step is not None:
                    continue
                inc = next(it)
                for suffix, c in (("_conv1", inc.one), ("_conv2", inc.a), ("_conv3", inc.b), ("_conv4", inc.pool)):
                    scope = "inception%d%s%s" % (layer, scope_str, suffix)
                    out.extend([(scope + "/weights", c.conv.weight, "conv"), (scope + "/biases", c.conv.bias, None)])
                    bn(scope, c.bn)

        def fcs(seq, scope_str):
            for n, fc in enumerate(seq, start=1):
                scope = "fc%d%s" % (n, scope_str)
                out.extend([(scope + "/weights", fc.lin.weight, "fc"), (scope + "/biases", fc.lin.bias, None)])
                bn(scope, fc.bn)

        convnet(self.gate_conv, "gating_conv")
        fcs(self.gate_fc, "noise")
        res3 = any(step is not None and step[1] == 3 for step in self.gate_conv.plan)
        for i, (conv, fc) in enumerate(zip(self.expert_conv, self.expert_fc)):
            convnet(conv, "Expert_%d" % i + ("_expert_conv" if res3 else ""))     # :277
            fcs(fc, "Expert_%d" % i)
        return out

    @torch.no_grad()
    def load_tf_variables(self, get):
        """Restore the network from the reference's variables: ``get(name)`` returns the array of TensorFlow variable
        ``name`` (see ``tf_variables``; e.g. a dict exported from a checkpoint of train_n_est_w_experts.py, or
        ``canonical_tf_names`` of one).  Raises KeyError / ValueError on a missing variable or a shape mismatch."""
        for name, dst, layout in self.tf_variables():
            a = torch.as_tensor(np.asarray(get(name)), dtype=dst.dtype)
            if layout == "conv":
                a = a.permute(4, 3, 0, 1, 2)
            elif layout == "fc":
                a = a.t()
            if tuple(a.shape) != tuple(dst.shape):
                raise ValueError("variable %s has shape %s, the network expects %s" % (name, tuple(a.shape), tuple(dst.shape)))
            dst.copy_(a)
        return self

    @torch.no_grad()
    def predict(self, mups):
        """What test_n_est_w_experts.py:148-152 keeps: the normal of the most probable expert."""
        prob, n_est = self.forward(mups)
        expert = prob.argmax(dim=0)
        return n_est[expert, torch.arange(n_est.shape[1], device=n_est.device)], expert, prob.transpose(0, 1)


def canonical_tf_names(variables):
    """{checkpoint name: array} -> the same arrays under the names ``tf_variables`` uses.  TensorFlow 1.x stores the shadow
    variables of tf.train.ExponentialMovingAverage (tf_util.py:474-490) as
    '<scope>/bn/<scope>/bn/moments/Squeeze/ExponentialMovingAverage' (mean) and '.../Squeeze_1/ExponentialMovingAverage'
    (variance); everything else keeps its name."""
    out = {}
    for name, a in variables.items():
        name = name[:-2] if name.endswith(":0") else name
        if name.endswith("/ExponentialMovingAverage"):
            head = name.split("/bn/")[0]
            leaf = "moving_variance" if "Squeeze_1" in name else "moving_mean"
            name = head + "/bn/" + leaf
        out[name] = a
    return out


def angular_rms_deg(n_a, n_b):
    """Unoriented RMS angle in degrees between two sets of normals (the metric of utils/evaluate.py:140-147),
    evaluated in float64 as atan2(|a x b|, |a . b|) so that angles far below acos's fp32 resolution survive."""
    a = n_a.double()
    b = n_b.double()
    cross = torch.linalg.cross(a, b, dim=-1).norm(dim=-1)
    dot = (a * b).sum(-1).abs()
    ang = torch.rad2deg(torch.atan2(cross, dot))
    return float(torch.sqrt((ang ** 2).mean()))
