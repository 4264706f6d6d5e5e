"""Oracle: the three sparse-conv networks of the reference, functional style over a state dict
that uses the reference's parameter names.

TEST INFRASTRUCTURE — see `oracle/__init__.py`.

Follows /root/reference/lidiff/models/minkunet.py:
  blocks            :13-80     MinkGlobalEnc :83-141
  MinkUNetDiff      :144-497   (time embedding :390-401, NN match :403-418, forward :420-497)
  MinkUNet (refine) :500-619
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import me_cpu as me

CS = [32, 32, 64, 128, 256, 256, 128, 96, 96]          # minkunet.py:88,150,507
EMBED_DIM = CS[-1]


class Net:
    """Evaluates one network from `sd` (tensor dict with reference key names, no prefix)."""

    def __init__(self, sd: dict, dtype=torch.float32, calibrate: bool = False, rng: torch.Generator | None = None):
        self.sd = sd
        self.dtype = dtype
        self.calibrate = calibrate          # overwrite BN running stats with (perturbed) batch stats
        self.rng = rng
        self.trace = {}

    # ---- leaf ops ----------------------------------------------------------------------------
    def _bn(self, prefix: str) -> dict:
        return {k: self.sd[f"{prefix}.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}

    def conv_bn(self, x, pconv, pbn, ks, stride=1, transposed=False, relu=True):
        y = me.conv(x, self.sd[f"{pconv}.kernel"], ks, stride, transposed)
        if self.calibrate:
            mean = y.F.mean(0)
            var = y.F.var(0, unbiased=False)
            jm = torch.randn(mean.shape, generator=self.rng) * 0.1
            jv = torch.rand(mean.shape, generator=self.rng) * 0.6 + 0.7
            self.sd[f"{pbn}.bn.running_mean"] = (mean + jm * torch.sqrt(var + 1e-5)).float()
            self.sd[f"{pbn}.bn.running_var"] = (var * jv + 1e-4).float()
        F = me.batchnorm_eval(y.F, self._bn(pbn))
        if relu:
            F = torch.relu(F)
        return y.replace(F)

    def mlp(self, x, p):
        """nn.Sequential(Linear, LeakyReLU(0.1), Linear[, ...])  keys p.0 / p.2"""
        dt = x.dtype
        h = torch.nn.functional.linear(x, self.sd[f"{p}.0.weight"].to(dt), self.sd[f"{p}.0.bias"].to(dt))
        h = torch.nn.functional.leaky_relu(h, 0.1)
        return torch.nn.functional.linear(h, self.sd[f"{p}.2.weight"].to(dt), self.sd[f"{p}.2.bias"].to(dt))

    # ---- blocks (minkunet.py:13-80) ------------------------------------------------------------
    def basic(self, x, p, ks, stride):
        return self.conv_bn(x, f"{p}.net.0", f"{p}.net.1", ks, stride)

    def deconv(self, x, p):
        return self.conv_bn(x, f"{p}.net.0", f"{p}.net.1", 2, 2, transposed=True)

    def residual(self, x, p):
        h = self.conv_bn(x, f"{p}.net.0", f"{p}.net.1", 3)
        h = self.conv_bn(h, f"{p}.net.3", f"{p}.net.4", 3, relu=False)
        if f"{p}.downsample.0.kernel" in self.sd:
            s = self.conv_bn(x, f"{p}.downsample.0", f"{p}.downsample.1", 1, relu=False)
        else:
            s = x
        return h.replace(torch.relu(h.F + s.F))

    def stem(self, x):
        h = self.conv_bn(x, "stem.0", "stem.1", 3)
        return self.conv_bn(h, "stem.3", "stem.4", 3)

    def stage(self, x, p):
        h = self.basic(x, f"{p}.0", 2, 2)
        h = self.residual(h, f"{p}.1")
        return self.residual(h, f"{p}.2")

    def up(self, x, skip, p):
        h = self.deconv(x, f"{p}.0")
        h = me.cat(h, skip)
        h = self.residual(h, f"{p}.1.0")
        return self.residual(h, f"{p}.1.1")

    # ---- MinkGlobalEnc.forward (minkunet.py:134-141) ------------------------------------------
    def global_enc(self, field: me.TensorField) -> me.SparseTensor:
        x = field.sparse()
        x = x.replace(x.F.to(self.dtype))
        x0 = self.stem(x)
        x1 = self.stage(x0, "stage1")
        x2 = self.stage(x1, "stage2")
        x3 = self.stage(x2, "stage3")
        x4 = self.stage(x3, "stage4")
        self.trace.update(enc_x0=x0, enc_x1=x1, enc_x2=x2, enc_x3=x3, enc_x4=x4)
        return x4

    # ---- MinkUNetDiff (minkunet.py:390-497) ----------------------------------------------------
    def timestep_embedding(self, t: torch.Tensor) -> torch.Tensor:
        half = EMBED_DIM // 2
        e = np.log(10000) / (half - 1)
        e = torch.from_numpy(np.exp(np.arange(0, half) * -e)).float()
        e = t[:, None] * e[None, :]
        return torch.cat([torch.sin(e), torch.cos(e)], dim=1)

    def gate(self, x, part, temb, latent, temp, latemp, swap=False, tag=""):
        idx = me.match_part_to_full(x.C, part.C)
        p = self.mlp(part.F[idx].to(self.dtype), latent)
        t = self.mlp(temb.to(self.dtype), temp)
        counts = torch.unique(x.C[:, 0], return_counts=True)[1]
        t = torch.repeat_interleave(t, counts, dim=0)
        w = self.mlp(torch.cat((t, p) if swap else (p, t), -1), latemp)
        self.trace[f"idx{tag}"] = idx
        self.trace[f"w{tag}"] = w
        return x * w

    def unet_diff(self, field, x_sparse, part, t) -> torch.Tensor:
        temb = self.timestep_embedding(t)
        xs = x_sparse.replace(x_sparse.F.to(self.dtype))
        g = lambda x, a, b, c, swap=False, tag="": self.gate(x, part, temb, a, b, c, swap, tag)
        x0 = self.stem(xs)
        x1 = self.stage(g(x0, "latent_stage1", "stage1_temp", "latemp_stage1", tag="0"), "stage1")
        x2 = self.stage(g(x1, "latent_stage2", "stage2_temp", "latemp_stage2", tag="1"), "stage2")
        x3 = self.stage(g(x2, "latent_stage3", "stage3_temp", "latemp_stage3", tag="2"), "stage3")
        x4 = self.stage(g(x3, "latent_stage4", "stage4_temp", "latemp_stage4", tag="3"), "stage4")
        y1 = self.up(g(x4, "latent_up1", "up1_temp", "latemp_up1", swap=True, tag="4"), x3, "up1")   # :461 (t4,p4)
        y2 = self.up(g(y1, "latent_up2", "up2_temp", "latemp_up2", tag="5"), x2, "up2")
        y3 = self.up(g(y2, "latent_up3", "up3_temp", "latemp_up3", tag="6"), x1, "up3")
        y4 = self.up(g(y3, "latent_up4", "up4_temp", "latemp_up4", tag="7"), x0, "up4")
        self.trace.update(x0=x0, x1=x1, x2=x2, x3=x3, x4=x4, y1=y1, y2=y2, y3=y3, y4=y4)
        return self.mlp(y4.slice(field), "last")

    # ---- MinkUNet refine net (minkunet.py:596-619) ---------------------------------------------
    def unet_refine(self, field) -> torch.Tensor:
        x = field.sparse()
        x = x.replace(x.F.to(self.dtype))
        x0 = self.stem(x)
        x1 = self.stage(x0, "stage1")
        x2 = self.stage(x1, "stage2")
        x3 = self.stage(x2, "stage3")
        x4 = self.stage(x3, "stage4")
        y1 = self.up(x4, x3, "up1")
        y2 = self.up(y1, x2, "up2")
        y3 = self.up(y2, x1, "up3")
        y4 = self.up(y3, x0, "up4")
        self.trace.update(x0=x0, x1=x1, x2=x2, x3=x3, x4=x4, y1=y1, y2=y2, y3=y3, y4=y4)
        return torch.tanh(self.mlp(y4.slice(field), "last"))


# ------------------------------------------------------------------------------------------------
# seeded random parameters with the reference's names and shapes (App. A.7); kernels ME-style
# U(-s,s), s=1/sqrt(Cin*K) (transpose: Cout*K) (App. A.4); BN affine randomised so BN != identity.
# ------------------------------------------------------------------------------------------------
def _conv_entry(sd, g, pconv, pbn, K, cin, cout, transposed=False):
    s = 1.0 / math.sqrt((cout if transposed else cin) * K)
    shape = (cin, cout) if K == 1 else (K, cin, cout)
    sd[f"{pconv}.kernel"] = (torch.rand(shape, generator=g) * 2 - 1) * s
    sd[f"{pbn}.bn.weight"] = torch.rand(cout, generator=g) + 0.5
    sd[f"{pbn}.bn.bias"] = torch.randn(cout, generator=g) * 0.1
    sd[f"{pbn}.bn.running_mean"] = torch.randn(cout, generator=g) * 0.05
    sd[f"{pbn}.bn.running_var"] = torch.rand(cout, generator=g) * 0.5 + 0.75
    sd[f"{pbn}.bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _linear_entry(sd, g, p, cin, cout):
    b = 1.0 / math.sqrt(cin)
    sd[f"{p}.weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * b
    sd[f"{p}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * b


def _residual_entries(sd, g, p, cin, cout):
    _conv_entry(sd, g, f"{p}.net.0", f"{p}.net.1", 27, cin, cout)
    _conv_entry(sd, g, f"{p}.net.3", f"{p}.net.4", 27, cout, cout)
    if cin != cout:
        _conv_entry(sd, g, f"{p}.downsample.0", f"{p}.downsample.1", 1, cin, cout)


def random_state_dict(kind: str, seed: int = 0, in_channels: int = 3, out_channels: int = 3) -> dict:
    """kind in {"enc" (MinkGlobalEnc), "diff" (MinkUNetDiff), "refine" (MinkUNet)}."""
    g = torch.Generator().manual_seed(seed)
    sd, cs = {}, CS
    _conv_entry(sd, g, "stem.0", "stem.1", 27, in_channels, cs[0])
    _conv_entry(sd, g, "stem.3", "stem.4", 27, cs[0], cs[0])
    for n in range(1, 5):
        cin, cout = cs[n - 1], cs[n]
        _conv_entry(sd, g, f"stage{n}.0.net.0", f"stage{n}.0.net.1", 8, cin, cin)
        _residual_entries(sd, g, f"stage{n}.1", cin, cout)
        _residual_entries(sd, g, f"stage{n}.2", cout, cout)
    if kind == "enc":
        return sd
    for n in range(1, 5):
        cin, cout, cskip = cs[3 + n], cs[4 + n], cs[4 - n]
        _conv_entry(sd, g, f"up{n}.0.net.0", f"up{n}.0.net.1", 8, cin, cout, transposed=True)
        _residual_entries(sd, g, f"up{n}.1.0", cout + cskip, cout)
        _residual_entries(sd, g, f"up{n}.1.1", cout, cout)
    _linear_entry(sd, g, "last.0", cs[8], 20)
    _linear_entry(sd, g, "last.2", 20, out_channels)
    if kind == "refine":
        return sd
    assert kind == "diff"
    gate_out = [cs[0], cs[1], cs[2], cs[3], cs[4], cs[5], cs[6], cs[7]]
    hidden = [cs[4], cs[4], cs[4], cs[4], cs[4], cs[5], cs[6], cs[7]]      # minkunet.py:171-175 ... :355-359
    names = ["stage1", "stage2", "stage3", "stage4", "up1", "up2", "up3", "up4"]
    for nm, h, co in zip(names, hidden, gate_out):
        _linear_entry(sd, g, f"latent_{nm}.0", cs[4], cs[4])
        _linear_entry(sd, g, f"latent_{nm}.2", cs[4], cs[4])
        _linear_entry(sd, g, f"latemp_{nm}.0", cs[4] + cs[4], h)
        _linear_entry(sd, g, f"latemp_{nm}.2", h, co)
        _linear_entry(sd, g, f"{nm}_temp.0", EMBED_DIM, EMBED_DIM)
        _linear_entry(sd, g, f"{nm}_temp.2", EMBED_DIM, cs[4])
    return sd
