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
