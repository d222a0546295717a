"""Parameter containers with the reference's module tree and state_dict keys
(models/common.py:9-68).  In the fused path these modules only OWN the weights
(aliased onto the flat CUDA parameter buffer); their `forward` methods exist for
composition with custom user modules and run as ordinary device ops."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CategoricalMLP(nn.Module):
    """in -> hidden -> softmax probabilities (keys: in_to_h.0.*, h_to_out.0.*)."""

    def __init__(self, in_dim, out_dim, h_dim):
        super().__init__()
        self.in_to_h = nn.Sequential(nn.Linear(in_dim, h_dim), nn.ReLU())
        self.h_to_out = nn.Sequential(nn.Linear(h_dim, out_dim), nn.Softmax(dim=1))

    def forward(self, x):
        return (self.h_to_out(self.in_to_h(x)),)


class GaussianMLP(nn.Module):
    """in -> hidden -> (mean, softplus std + min_std)
    (keys: in_to_h.0.*, h_to_mean.*, h_to_std.0.*)."""

    def __init__(self, in_dim, out_dim, h_dim, min_std=1e-3):
        super().__init__()
        self.min_std = min_std
        self.in_to_h = nn.Sequential(nn.Linear(in_dim, h_dim), nn.ReLU())
        self.h_to_mean = nn.Linear(h_dim, out_dim)
        self.h_to_std = nn.Sequential(nn.Linear(h_dim, out_dim), nn.Softplus())

    def forward(self, x):
        h = self.in_to_h(x)
        return self.h_to_mean(h), self.h_to_std(h) + self.min_std


class GaussianGTF(nn.Module):
    """Gated transition function (keys: z_to_gate.{0,2}.*, z_lin.*,
    z_nonlin.{0,2}.*, z_to_std.0.*)."""

    def __init__(self, z_dim, h_dim, min_std=0):
        super().__init__()
        self.min_std = min_std
        self.z_to_gate = nn.Sequential(nn.Linear(z_dim, h_dim), nn.ReLU(),
                                       nn.Linear(h_dim, z_dim), nn.Sigmoid())
        self.z_lin = nn.Linear(z_dim, z_dim)
        self.z_nonlin = nn.Sequential(nn.Linear(z_dim, h_dim), nn.ReLU(),
                                      nn.Linear(h_dim, z_dim))
        self.z_to_std = nn.Sequential(nn.Linear(z_dim, z_dim), nn.Softplus())

    def forward(self, z):
        gate = self.z_to_gate(z)
        nonlin = self.z_nonlin(z)
        mean = torch.lerp(self.z_lin(z), nonlin, gate)
        return mean, self.z_to_std(nonlin) + self.min_std


# ---------------------------------------------------------------------------------------
# Image modules of the Weizmann model (models/common.py:70-175).  They are injected by the
# caller as custom `encoders=` / `decoders=` (weizmann.py:64-76, which looks them up as
# `models.common.ImageEncoder / ImageDecoder`), stay ordinary torch / cuDNN modules and hand
# their (mean, std) / pixel probabilities to the fused temporal core (out of
# scope for the kernels; "next" item §8f-3).  Module names follow the reference so that
# state_dict keys — including the doubly registered `conv` / `net.0` entry — match.
# ---------------------------------------------------------------------------------------
def _conv_block(module, layer, n_out, last):
    """layer [-> BatchNorm2d -> ReLU]; the bare layer when it is the last of a stack."""
    module.net = layer if last else nn.Sequential(layer, nn.BatchNorm2d(n_out), nn.ReLU())
    nn.init.xavier_uniform_(layer.weight)


class Conv(nn.Module):
    """Strided 3x3 convolution block (keys: conv.*, net.0.* / net.*, net.1.*)."""

    def __init__(self, n_channels, n_kernels, kernel_size=3, stride=2, padding=1, last=False):
        super().__init__()
        self.conv = nn.Conv2d(n_channels, n_kernels, kernel_size, stride, padding)
        _conv_block(self, self.conv, n_kernels, last)

    def forward(self, x):
        return self.net(x)


class Deconv(nn.Module):
    """Strided 4x4 transposed convolution block (keys: deconv.*, net.*)."""

    def __init__(self, n_channels, n_kernels, kernel_size=4, stride=2, padding=1, last=False):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(n_channels, n_kernels, kernel_size, stride, padding)
        _conv_block(self, self.deconv, n_kernels, last)

    def forward(self, x):
        return self.net(x)


def _stack_widths(n_kernels, n_layers):
    """Channel widths of the hidden feature maps, narrowest first: n_kernels / 2^(L-1) .. n_kernels."""
    return [n_kernels // 2 ** (n_layers - 1 - i) for i in range(n_layers)]


class ImageEncoder(nn.Module):
    """img -> 3 stride-2 conv blocks -> (z mean, softplus z std); `gauss_out=False` returns the
    feature maps (keys: conv_stack.<i>.*, feat_to_z_mean.*, feat_to_z_std.0.*)."""

    def __init__(self, z_dim, gauss_out=True, img_size=64, n_channels=3, n_kernels=64, n_layers=3):
        super().__init__()
        self.feat_size = img_size // 2 ** n_layers
        self.feat_dim = self.feat_size ** 2 * n_kernels
        widths = [n_channels] + _stack_widths(n_kernels, n_layers)
        self.conv_stack = nn.Sequential(*[Conv(widths[i], widths[i + 1], last=(i == n_layers - 1))
                                          for i in range(n_layers)])
        self.gauss_out = gauss_out
        if gauss_out:
            self.feat_to_z_mean = nn.Linear(self.feat_dim, z_dim)
            self.feat_to_z_std = nn.Sequential(nn.Linear(self.feat_dim, z_dim), nn.Softplus())
            nn.init.xavier_uniform_(self.feat_to_z_mean.weight)
            nn.init.xavier_uniform_(self.feat_to_z_std[0].weight)

    def forward(self, x):
        feats = self.conv_stack(x)
        if not self.gauss_out:
            return feats
        flat = feats.reshape(-1, self.feat_dim)
        return self.feat_to_z_mean(flat), self.feat_to_z_std(flat)


class ImageDecoder(nn.Module):
    """z -> Linear + ReLU -> 3 stride-2 transposed conv blocks -> sigmoid pixel probabilities
    (keys: z_to_feat.0.*, deconv_stack.<i>.*)."""

    def __init__(self, z_dim, img_size=64, n_channels=3, n_kernels=64, n_layers=3):
        super().__init__()
        self.feat_size = img_size // 2 ** n_layers
        self.feat_dim = self.feat_size ** 2 * n_kernels
        self.feat_shape = (n_kernels, self.feat_size, self.feat_size)
        self.z_to_feat = nn.Sequential(nn.Linear(z_dim, self.feat_dim), nn.ReLU())
        widths = _stack_widths(n_kernels, n_layers)[::-1] + [n_channels]
        self.deconv_stack = nn.Sequential(*([Deconv(widths[i], widths[i + 1], last=(i == n_layers - 1))
                                             for i in range(n_layers)] + [nn.Sigmoid()]))
        nn.init.xavier_uniform_(self.z_to_feat[0].weight)

    def forward(self, z):
        feats = self.z_to_feat(z).reshape(-1, *self.feat_shape)
        return (self.deconv_stack(feats),)
