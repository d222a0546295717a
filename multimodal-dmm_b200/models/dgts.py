"""Base class shared by deep generative time-series models: the reference's
MultiDGTS surface (models/dgts.py:12-180) on top of the CUDA ops.

`product_of_experts` / `mean_of_experts` are the public tensor-level helpers (the
hot loop fuses them inside the filter kernels); `step` is the generic multimodal +
unimodal ELBO (models/dgts.py:85-130) composed from differentiable CUDA ops, used
when a model cannot take the fully fused `MultiDMM.step` path."""
import torch
import torch.nn as nn

from . import losses


class MultiDGTS(nn.Module):

    def product_of_experts(self, mean, std, mask=None, eps=1e-8):
        """Product of Gaussian experts along dim 0 (models/dgts.py:15-51): precision
        sign(std)/(std^2+eps) so that a negated std divides an expert out; masked
        experts contribute nothing; a NaN mean (0/0) becomes 0."""
        var = std.pow(2) + eps
        prec = std.sign() / var
        if mask is None:
            mask = ~torch.isnan(var).any(dim=-1)
        w = mask.to(mean.dtype).unsqueeze(-1)
        prec = prec * w
        total = prec.sum(dim=0)
        out_mean = torch.nan_to_num((mean * w * prec).sum(dim=0) / total, nan=0.0,
                                    posinf=float('inf'), neginf=float('-inf'))
        return out_mean, total.reciprocal().pow(0.5)

    def mean_of_experts(self, mean, std, mask=None):
        """Moment-matched Gaussian of an equally weighted mixture along dim 0
        (models/dgts.py:53-83)."""
        if mask is None:
            mask = ~torch.isnan(std).any(dim=-1)
        w = mask.to(mean.dtype).unsqueeze(-1)
        mean = mean * w
        first = mean.mean(dim=0)
        second = (std.pow(2) * w).mean(dim=0) + (mean.pow(2).mean(dim=0) - first.pow(2))
        return first, second.pow(0.5)

    def step(self, inputs, mask, kld_mult, rec_mults, targets=None, uni_loss=True, **kwargs):
        """Multimodal ELBO over all given modalities (when the model has more than
        one) plus one ELBO per single modality (models/dgts.py:85-130)."""
        inputs = {m: inputs[m] for m in inputs if m in self.modalities}
        if targets is None:
            targets = inputs
        total = 0
        if len(self.modalities) > 1:
            infer, prior, recon = self.forward(inputs, **kwargs)
            total = total + self.loss(targets, infer, prior, recon, mask, kld_mult, rec_mults)
        if not uni_loss:
            return total
        for m in self.modalities:
            infer, prior, recon = self.forward({m: inputs[m]}, **kwargs)
            total = total + self.loss({m: targets[m]}, infer, prior, recon, mask, kld_mult,
                                      rec_mults)
        return total

    def loss(self, inputs, infer, prior, recon, mask=1, kld_mult=1.0, rec_mults={}, avg=False):
        """kld_mult * KL + reconstruction (models/dgts.py:132-145)."""
        total = kld_mult * self.kld_loss(infer, prior, mask) + \
            self.rec_loss(inputs, recon, mask, rec_mults)
        if avg:
            if torch.is_tensor(mask):
                total = total / mask.sum()
            else:
                shape = inputs[self.modalities[-1]].shape
                total = total / (shape[0] * shape[1])
        return total

    def kld_loss(self, infer, prior, mask=None):
        """models/dgts.py:147-152."""
        return losses.kld_gauss(infer[0], infer[1], prior[0], prior[1], mask)

    def rec_loss(self, inputs, recon, mask=None, rec_mults={}):
        """Weighted sum of per-modality negative log-likelihoods
        (models/dgts.py:154-175)."""
        total = 0.0
        for m in self.modalities:
            if m not in inputs:
                continue
            mult = rec_mults.get(m, 1.0)
            if mult == 0:
                continue
            dist = self.dists[m]
            if dist == 'Bernoulli':
                total = total + mult * losses.nll_bernoulli(recon[m][0], inputs[m], mask)
            elif dist == 'Categorical':
                total = total + mult * losses.nll_categorical(recon[m][0], inputs[m], mask)
            elif dist == 'Normal':
                total = total + mult * losses.nll_gauss(recon[m][0], recon[m][1], inputs[m], mask)
        return total

    def _sample_gauss(self, mean, std):
        """Reparameterised draw (models/dgts.py:177-180); the fused kernels generate
        their own noise, this serves the tensor-level helpers."""
        return torch.randn_like(std) * std + mean
