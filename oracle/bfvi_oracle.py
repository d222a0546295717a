"""CPU oracle for the BFVI ELBO step of the Multimodal Deep Markov Model.

TEST INFRASTRUCTURE ONLY.  This file is a plain-PyTorch (CPU, fp32 or fp64)
restatement of the reference algorithm.  It is the checker for the CUDA path
and the `cpu_baseline` / `--impl reference` arm of bench.py.  Nothing under
`multimodal-dmm_b200/` may import it; the product path has no CPU fallback.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference
(`/root/reference/models/dmm.py`, with the documented `1 - bool` shim) and this
restatement on identical weights / inputs / injected noise and asserts equal
results before it writes `tests/golden/*.pt`; `tests/test_oracle_golden.py`
re-checks the restatement against those committed vectors on every run.

Every function cites the reference lines it restates (paths are relative to the
reference repository root).  The restatement is functional: weights live in a
flat `dict[str, Tensor]` whose keys are the reference `state_dict` keys, and
reparameterisation noise comes from an injected `draw(shape)` callable, which
replaces `MultiDGTS._sample_gauss` (models/dgts.py:177-180).
"""
import math

import torch
import torch.nn.functional as F

MIN_STD_MLP = 1e-3      # GaussianMLP default (models/common.py:27)
POE_EPS = 1e-8          # product_of_experts eps (models/dgts.py:15)


# --------------------------------------------------------------------------
# parameter helpers
# --------------------------------------------------------------------------
def init_params(modalities, dims, dists=None, h_dim=32, z_dim=32,
                z0_mean=0.0, z0_std=1.0, seed=0, dtype=torch.float32,
                scale=1.0):
    """Random parameters with the reference's state_dict keys and shapes.

    Shapes follow models/dmm.py:75-116 and models/common.py:9-68.  The values
    are NOT the reference's initialiser (tests copy a real state_dict when they
    need seeded-init equality); `scale` > 1 makes the nonlinearities matter.
    """
    g = torch.Generator().manual_seed(seed)
    if dists is None:
        dists = ['Normal'] * len(modalities)
    p = {}

    def lin(prefix, n_out, n_in):
        bound = scale / math.sqrt(n_in)
        p[prefix + '.weight'] = ((torch.rand(n_out, n_in, generator=g) * 2 - 1)
                                 * bound).to(dtype)
        p[prefix + '.bias'] = ((torch.rand(n_out, generator=g) * 2 - 1)
                               * bound).to(dtype)

    def gauss_mlp(prefix, n_in, n_out):
        lin(prefix + '.in_to_h.0', h_dim, n_in)
        lin(prefix + '.h_to_mean', n_out, h_dim)
        lin(prefix + '.h_to_std.0', n_out, h_dim)

    p['z0_mean'] = (z0_mean * torch.ones(1, z_dim)
                    + 0.1 * torch.randn(1, z_dim, generator=g)).to(dtype)
    p['z0_log_std'] = (math.log(z0_std) * torch.ones(1, z_dim)
                       + 0.1 * torch.randn(1, z_dim, generator=g)).to(dtype)
    for m, d, dist in zip(modalities, dims, dists):
        d = int(torch.tensor(d).prod()) if not isinstance(d, int) else d
        if dist == 'Categorical':
            p['enc.%s.0.weight' % m] = torch.randn(d, h_dim, generator=g).to(dtype)
            gauss_mlp('enc.%s.2' % m, h_dim, z_dim)
        else:
            gauss_mlp('enc.%s' % m, d, z_dim)
    for m, d, dist in zip(modalities, dims, dists):
        d = int(torch.tensor(d).prod()) if not isinstance(d, int) else d
        if dist == 'Categorical':
            lin('dec.%s.in_to_h.0' % m, h_dim, z_dim)
            lin('dec.%s.h_to_out.0' % m, d, h_dim)
        else:
            gauss_mlp('dec.%s' % m, z_dim, d)
    for direction in ('fwd', 'bwd'):
        pre = 'trans.%s' % direction
        lin(pre + '.z_to_gate.0', h_dim, z_dim)
        lin(pre + '.z_to_gate.2', z_dim, h_dim)
        lin(pre + '.z_lin', z_dim, z_dim)
        lin(pre + '.z_nonlin.0', h_dim, z_dim)
        lin(pre + '.z_nonlin.2', z_dim, h_dim)
        lin(pre + '.z_to_std.0', z_dim, z_dim)
    return p


class NoiseTape(object):
    """Hands out pre-generated noise tensors in call order.

    Stands in for `torch.FloatTensor(size).normal_()` of models/dgts.py:179 so
    the oracle, the reference and the CUDA path all see identical draws.
    """

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.pos = 0

    def __call__(self, shape):
        t = self.tensors[self.pos]
        self.pos += 1
        assert tuple(t.shape) == tuple(shape), \
            "noise draw %d: tape %s vs requested %s" % (
                self.pos - 1, tuple(t.shape), tuple(shape))
        return t


class RandomDraw(object):
    """Fresh N(0,1) draws (used by the timing arm only)."""

    def __init__(self, dtype=torch.float32, seed=0):
        self.g = torch.Generator().manual_seed(seed)
        self.dtype = dtype

    def __call__(self, shape):
        return torch.randn(*shape, generator=self.g, dtype=self.dtype)


# --------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------
class OracleDMM(object):
    """Functional restatement of models/dmm.py:28-554 + models/dgts.py:12-180."""

    def __init__(self, modalities, dims, params, dists=None, h_dim=32,
                 z_dim=32, min_std=1e-3, draw=None):
        self.modalities = list(modalities)
        self.dims = dict(zip(self.modalities, dims))
        self.dists = dict(zip(self.modalities,
                              dists or ['Normal'] * len(self.modalities)))
        self.h_dim, self.z_dim, self.min_std = h_dim, z_dim, min_std
        self.p = params
        self.draw = draw

    # ---- building blocks -------------------------------------------------
    def _linear(self, key, x):
        return F.linear(x, self.p[key + '.weight'], self.p[key + '.bias'])

    def gaussian_mlp(self, prefix, x):
        """models/common.py:25-41 (Softplus beta=1, threshold=20)."""
        h = torch.relu(self._linear(prefix + '.in_to_h.0', x))
        mean = self._linear(prefix + '.h_to_mean', h)
        std = F.softplus(self._linear(prefix + '.h_to_std.0', h)) + MIN_STD_MLP
        return mean, std

    def categorical_mlp(self, prefix, x):
        """models/common.py:9-23."""
        h = torch.relu(self._linear(prefix + '.in_to_h.0', x))
        return (torch.softmax(self._linear(prefix + '.h_to_out.0', h), dim=1),)

    def gtf(self, direction, z):
        """models/common.py:43-68."""
        pre = 'trans.%s' % direction
        gate = torch.sigmoid(self._linear(
            pre + '.z_to_gate.2', torch.relu(self._linear(pre + '.z_to_gate.0', z))))
        z_lin = self._linear(pre + '.z_lin', z)
        z_nonlin = self._linear(
            pre + '.z_nonlin.2', torch.relu(self._linear(pre + '.z_nonlin.0', z)))
        z_std = F.softplus(self._linear(pre + '.z_to_std.0', z_nonlin)) + self.min_std
        z_mean = (1 - gate) * z_lin + gate * z_nonlin
        return z_mean, z_std

    @staticmethod
    def product_of_experts(mean, std, mask=None, eps=POE_EPS):
        """models/dgts.py:15-51.  Expert axis first; sign(std) carries the
        inverse-prior trick; NaN means are zero-filled."""
        var = std.pow(2) + eps
        prec = 1. / var * std.sign()
        if mask is None:
            mask = ~torch.isnan(var).any(dim=-1)
        w = mask.to(mean.dtype).unsqueeze(-1)
        prec = prec * w
        mean = mean * w
        prec_sum = torch.sum(prec, dim=0)
        out_mean = torch.sum(mean * prec, dim=0) / prec_sum
        out_mean = torch.where(torch.isnan(out_mean),
                               torch.zeros_like(out_mean), out_mean)
        out_std = (1. / prec_sum).pow(0.5)
        return out_mean, out_std

    @staticmethod
    def mean_of_experts(mean, std, mask=None):
        """models/dgts.py:53-83 (mixture moment matching over axis 0)."""
        if mask is None:
            mask = ~torch.isnan(std).any(dim=-1)
        w = mask.to(mean.dtype).unsqueeze(-1)
        mean = mean * w
        var = std.pow(2) * w
        out_mean = torch.mean(mean, dim=0)
        out_var = torch.mean(var, dim=0) + (torch.mean(mean.pow(2), dim=0)
                                            - out_mean.pow(2))
        return out_mean, out_var.pow(0.5)

    def sample_gauss(self, mean, std):
        """models/dgts.py:177-180 with the draw injected."""
        eps = self.draw(tuple(std.shape)).to(std.dtype)
        return eps * std + mean

    def prior(self, shape):
        """models/dmm.py:124-129."""
        mean = self.p['z0_mean'].repeat(*shape)
        std = (self.p['z0_log_std'].exp() + self.min_std).repeat(*shape)
        mask = torch.ones(shape[:-1], dtype=torch.bool)
        return mean, std, mask

    # ---- encode / decode -------------------------------------------------
    def encode(self, inputs):
        """models/dmm.py:131-190 (combine=False branch)."""
        first = inputs[list(inputs.keys())[0]]
        t_max, b_dim = first.shape[:2]
        means, stds, masks = [], [], []
        for m in self.modalities:
            if m not in inputs:
                continue
            x = inputs[m]
            mask_m = ~torch.isnan(x).flatten(2, -1).any(dim=-1)
            x = torch.where(torch.isnan(x), torch.zeros_like(x), x).detach()
            if self.dists[m] == 'Categorical':
                idx = x.long().flatten(0, 1)
                h = torch.relu(F.embedding(idx, self.p['enc.%s.0.weight' % m]))
                mu, sd = self.gaussian_mlp('enc.%s.2' % m, h)
            else:
                mu, sd = self.gaussian_mlp('enc.%s' % m, x.flatten(0, 1).flatten(1, -1))
            means.append(mu.reshape(t_max, b_dim, -1))
            stds.append(sd.reshape(t_max, b_dim, -1))
            masks.append(mask_m)
        return torch.stack(means), torch.stack(stds), torch.stack(masks)

    def decode(self, z):
        """models/dmm.py:192-212."""
        t_max, b_dim = z.shape[:2]
        recon = {}
        for m in self.modalities:
            flat = z.reshape(-1, self.z_dim)
            if self.dists[m] == 'Categorical':
                out = self.categorical_mlp('dec.%s' % m, flat)
            else:
                out = self.gaussian_mlp('dec.%s' % m, flat)
            recon[m] = tuple(r.reshape(t_max, b_dim, *r.shape[1:]) for r in out)
        return recon

    # ---- temporal core ---------------------------------------------------
    def z_next(self, z, direction, glb):
        """models/dmm.py:214-258.  z is (K, B, Z)."""
        glb_mean, glb_std = glb
        k = z.shape[0]
        if k == 1:
            q_mean, q_std = self.gtf(direction, z[0])
            return self.product_of_experts(torch.stack([glb_mean, q_mean]),
                                           torch.stack([glb_std, q_std]))
        q_mean, q_std = self.gtf(direction, z.reshape(-1, self.z_dim))
        mean, std = self.product_of_experts(
            torch.stack([glb_mean.repeat(k, 1), q_mean]),
            torch.stack([glb_std.repeat(k, 1), q_std]))
        return self.mean_of_experts(mean.reshape(z.shape), std.reshape(z.shape))

    def z_sample(self, t_max, b_dim, direction='fwd', sample=True,
                 n_particles=1, inclusive=False):
        """models/dmm.py:260-317 for z_init=None."""
        glb_mean, glb_std, _ = self.prior((b_dim, 1))
        means, stds = [], []
        mean_t, std_t = glb_mean, glb_std
        if inclusive:
            means.append(mean_t)
            stds.append(std_t)
        for _ in range(t_max - int(inclusive)):
            if sample or n_particles > 1:
                z_t = self.sample_gauss(mean_t.expand(n_particles, -1, -1),
                                        std_t.expand(n_particles, -1, -1))
            else:
                z_t = mean_t.unsqueeze(0)
            mean_t, std_t = self.z_next(z_t, direction, (glb_mean, glb_std))
            means.append(mean_t)
            stds.append(std_t)
        if direction == 'bwd':
            means.reverse()
            stds.reverse()
        return torch.stack(means), torch.stack(stds)

    def z_filter(self, z_mean, z_std, z_masks, direction='fwd', sample=True,
                 n_particles=1, sample_init=False):
        """models/dmm.py:319-412."""
        t_max, b_dim = z_mean[0].shape[:2]
        glb_mean, glb_std, _ = self.prior((b_dim, 1))
        order = list(range(t_max))
        if direction == 'bwd':
            order.reverse()
        pri_m, pri_s, inf_m, inf_s, samples = {}, {}, {}, {}, {}
        z_t = None
        ones = torch.ones((1, b_dim), dtype=z_masks.dtype)
        for i, t in enumerate(order):
            if i == 0:
                pm, ps = glb_mean, glb_std
            else:
                pm, ps = self.z_next(z_t, direction, (glb_mean, glb_std))
            pri_m[t], pri_s[t] = pm, ps
            im, isd = self.product_of_experts(
                torch.cat([pm.unsqueeze(0), z_mean[:, t]], 0),
                torch.cat([ps.unsqueeze(0), z_std[:, t]], 0),
                torch.cat([ones, z_masks[:, t]], 0))
            inf_m[t], inf_s[t] = im, isd
            if sample or n_particles > 1 or (i == 0 and sample_init):
                z_t = self.sample_gauss(im.expand(n_particles, -1, -1),
                                        isd.expand(n_particles, -1, -1))
                samples[t] = z_t.mean(dim=0)
            else:
                z_t = im.unsqueeze(0)
                samples[t] = im
        st = lambda d: torch.stack([d[t] for t in range(t_max)])
        return (st(inf_m), st(inf_s)), (st(pri_m), st(pri_s)), st(samples)

    def forward(self, inputs, lengths, mode='fsmooth', sample=True,
                sample_init=False, flt_particles=1, smt_particles=1):
        """models/dmm.py:420-494."""
        t_max, b_dim = max(lengths), len(lengths)
        obs_mean, obs_std, obs_mask = self.encode(inputs)
        direction = 'fwd' if mode in ('ffilter', 'bsmooth') else 'bwd'
        flt_init = sample_init if mode in ('ffilter', 'bfilter') else False
        infer, prior, z_samples = self.z_filter(
            obs_mean, obs_std, obs_mask, direction=direction, sample=sample,
            n_particles=flt_particles, sample_init=flt_init)
        if mode in ('fsmooth', 'bsmooth'):
            direction = 'fwd' if mode == 'fsmooth' else 'bwd'
            inv_mean, inv_std, inv_mask = self.prior((t_max, b_dim, 1))
            inv_std = -inv_std
            flt_mean, flt_std = prior
            flt_mask = torch.ones((t_max, b_dim), dtype=torch.bool)
            flt_mask[-1] = False
            infer, prior, z_samples = self.z_filter(
                torch.cat([obs_mean, flt_mean[None], inv_mean[None]], 0),
                torch.cat([obs_std, flt_std[None], inv_std[None]], 0),
                torch.cat([obs_mask, flt_mask[None], inv_mask[None]], 0),
                direction=direction, sample=sample, n_particles=smt_particles,
                sample_init=sample_init)
        recon = self.decode(z_samples)
        return infer, prior, recon

    # ---- losses ----------------------------------------------------------
    def loss(self, targets, infer, prior, recon, mask, kld_mult, rec_mults):
        """models/dgts.py:132-175."""
        total = kld_mult * kld_gauss(infer[0], infer[1], prior[0], prior[1], mask)
        total = total + self.rec_loss(targets, recon, mask, rec_mults)
        return total

    def rec_loss(self, targets, recon, mask, rec_mults):
        """models/dgts.py:154-175."""
        total = 0.0
        for m in self.modalities:
            if m not in targets:
                continue
            mult = rec_mults.get(m, 1.0)
            if mult == 0:
                continue
            if self.dists[m] == 'Bernoulli':
                total = total + mult * nll_bernoulli(recon[m][0], targets[m], mask)
            elif self.dists[m] == 'Categorical':
                total = total + mult * nll_categorical(recon[m][0], targets[m], mask)
            else:
                total = total + mult * nll_gauss(recon[m][0], recon[m][1],
                                                 targets[m], mask)
        return total

    def dgts_step(self, inputs, mask, kld_mult, rec_mults, targets, uni_loss,
                  lengths, **fw):
        """models/dgts.py:85-130."""
        inputs = {m: inputs[m] for m in inputs if m in self.modalities}
        if targets is None:
            targets = inputs
        total = 0
        if len(self.modalities) > 1:
            infer, prior, recon = self.forward(inputs, lengths, **fw)
            total = total + self.loss(targets, infer, prior, recon, mask,
                                      kld_mult, rec_mults)
        if not uni_loss:
            return total
        for m in self.modalities:
            infer, prior, recon = self.forward({m: inputs[m]}, lengths, **fw)
            total = total + self.loss({m: targets[m]}, infer, prior, recon,
                                      mask, kld_mult, rec_mults)
        return total

    def kld_prior(self, n_particles, direction):
        """models/dmm.py:496-501."""
        glb_mean, glb_std, _ = self.prior((1, 1, 1))
        nxt_mean, nxt_std = self.z_sample(1, 1, direction, True, n_particles)
        return kld_gauss(glb_mean, glb_std, nxt_mean, nxt_std)

    def step(self, inputs, mask, kld_mult, rec_mults, targets=None,
             uni_loss=True, lengths=None, f_mode='bfilter', s_mode='fsmooth',
             f_mult=0.5, s_mult=0.5, match_mult=0.01, train_particles=25,
             match_particles=50, **fw):
        """models/dmm.py:503-554."""
        total = 0
        if match_mult > 0:
            n_obs = mask.sum().to(self.p['z0_mean'].dtype)
            total = total + (match_mult * kld_mult * n_obs
                             * self.kld_prior(match_particles, 'fwd'))
            total = total + (match_mult * kld_mult * n_obs
                             * self.kld_prior(match_particles, 'bwd'))
        total = total + f_mult * self.dgts_step(
            inputs, mask, kld_mult, rec_mults, targets, uni_loss, lengths,
            mode=f_mode, **fw)
        total = total + s_mult * self.dgts_step(
            inputs, mask, kld_mult, rec_mults, targets, uni_loss, lengths,
            mode=s_mode, flt_particles=train_particles, **fw)
        return total


# --------------------------------------------------------------------------
# free loss functions (models/losses.py)
# --------------------------------------------------------------------------
def kld_gauss(mean_1, std_1, mean_2, std_2, mask=None):
    """models/losses.py:14-21."""
    el = (2 * torch.log(std_2) - 2 * torch.log(std_1)
          + (std_1.pow(2) + (mean_1 - mean_2).pow(2)) / std_2.pow(2) - 1)
    if mask is not None:
        el = el.masked_select(mask.bool())
    return 0.5 * torch.sum(el)


def _elem_mask(x, mask):
    """models/losses.py:34-38 / 56-60 / 78-82: observed-and-in-sequence."""
    obs = ~torch.isnan(x)
    if mask is None:
        return obs
    shape = list(mask.shape) + [1] * (x.dim() - mask.dim())
    return obs & mask.bool().view(*shape)


def nll_gauss(mean, std, x, mask=None):
    """models/losses.py:68-89."""
    keep = _elem_mask(x, mask)
    x = torch.where(torch.isnan(x), torch.zeros_like(x), x).detach()
    el = 0.5 * ((x - mean) / std).pow(2) + std.log() + 0.5 * math.log(2 * math.pi)
    return torch.sum(el.masked_select(keep))


def nll_bernoulli(theta, x, mask=None):
    """models/losses.py:23-42."""
    keep = _elem_mask(x, mask)
    return F.binary_cross_entropy(theta.masked_select(keep),
                                  x.masked_select(keep), reduction='sum')


def nll_categorical(probs, x, mask=None):
    """models/losses.py:44-66.  Quirk kept: F.nll_loss is fed probabilities,
    not log-probabilities, so the value is -sum(p[label])."""
    keep = _elem_mask(x, mask)
    cols = [probs[:, :, k:k + 1].masked_select(keep)
            for k in range(probs.shape[2])]
    sel = torch.stack(cols, dim=-1)
    return F.nll_loss(sel, x.masked_select(keep).long(), reduction='sum')


# --------------------------------------------------------------------------
# noise bookkeeping shared by tests / bench (draw order of one default step)
# --------------------------------------------------------------------------
def step_sets(n_mods, uni_loss=True):
    """Input sets of one DGTS step in evaluation order (models/dgts.py:119-129):
    the full set first (only when M > 1), then each single modality.  Returned
    as lists of modality indices."""
    sets = []
    if n_mods > 1:
        sets.append(list(range(n_mods)))
    if uni_loss:
        sets.extend([[i] for i in range(n_mods)])
    return sets


def make_step_noise(n_sets, t_max, b_dim, z_dim, train_particles=25,
                    match_particles=50, seed=0, dtype=torch.float32):
    """Noise for one default `step` in the CUDA path's layout.

    match: (2, K_match, Z)         [fwd, bwd] kld_prior draws
    filt : (S, T, B, 1, Z)         f_mode pass (one particle)
    sflt : (S, T, B, K, Z)         s_mode filtering pass (K particles)
    ssmt : (S, T, B, 1, Z)         s_mode smoothing pass
    indexed by the time index t at which the draw is consumed.
    """
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    return {'match': rn(2, match_particles, z_dim),
            'filt': rn(n_sets, t_max, b_dim, 1, z_dim),
            'sflt': rn(n_sets, t_max, b_dim, train_particles, z_dim),
            'ssmt': rn(n_sets, t_max, b_dim, 1, z_dim)}


def step_noise_tape(noise, f_mode='bfilter', s_mode='fsmooth', with_match=True):
    """Serialise `make_step_noise` tensors into the reference's draw order
    (models/dmm.py:541-553): kld_prior fwd, bwd; then per set
    the f_mode pass; then per set the s_mode filtering pass followed by its
    smoothing pass.  Each draw has the reference shape (K, B, Z)."""
    n_sets, t_max = noise['filt'].shape[:2]
    tape = []
    if with_match:      # models/dmm.py:540 draws only when match_mult > 0
        tape = [noise['match'][0].unsqueeze(1), noise['match'][1].unsqueeze(1)]
    fdir_rev = f_mode == 'bfilter'
    sflt_rev = s_mode == 'fsmooth'         # filtering pass runs opposite to smoothing
    order = lambda rev: (range(t_max - 1, -1, -1) if rev else range(t_max))
    for s in range(n_sets):
        for t in order(fdir_rev):
            tape.append(noise['filt'][s, t].permute(1, 0, 2).contiguous())
    for s in range(n_sets):
        for t in order(sflt_rev):
            tape.append(noise['sflt'][s, t].permute(1, 0, 2).contiguous())
        for t in order(not sflt_rev):
            tape.append(noise['ssmt'][s, t].permute(1, 0, 2).contiguous())
    return NoiseTape(tape)
