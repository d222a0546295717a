"""Parity of the Python drop-in API (MultiDMM.step / forward / losses) on CUDA
against the reference's golden outputs and the oracle."""
import pytest
import torch

import bfvi_oracle as bo
from conftest import golden_names, load_golden, rel_err
import multimodal_dmm_b200.models as models

SMALL = golden_names()        # 'medium_dims' (Z=16, H=48) is served by the large-dim family
ELBO_TOL, GRAD_TOL = 1e-4, 1e-3
pytestmark = pytest.mark.gpu


def build(fx):
    m = models.MultiDMM(fx['modalities'], fx['dims'], h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                        min_std=fx['min_std'], device=torch.device('cuda:0'))
    m.load_state_dict(fx['state_dict'])
    return m


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize('name', SMALL)
def test_step_api_matches_reference_golden(name):
    fx = load_golden(name)
    m = build(fx)
    m.train()
    loss = m.step(cuda(fx['inputs']), fx['mask'].cuda(), fx['kld_mult'], fx['rec_mults'],
                  targets=cuda(fx['targets']), lengths=fx['lengths'], noise=fx['noise'],
                  **fx['step_kwargs'])
    assert loss.dim() == 0 and loss.requires_grad
    ref = fx['ref_loss_fp64']
    assert abs(loss.item() - ref) / abs(ref) < ELBO_TOL
    loss /= sum(fx['lengths'])            # trainer.py:242 divides in place
    loss.backward()
    for k, p in m.named_parameters():
        g_ref = fx['ref_grads_fp64'][k]
        g = torch.zeros_like(g_ref) if p.grad is None else p.grad.cpu()
        if g_ref.norm() == 0:
            assert g.norm() == 0, k
        else:
            assert rel_err(g, g_ref) < GRAD_TOL, (k, rel_err(g, g_ref))


def _assemble(draws, t_max, direction):
    order = list(range(t_max - 1, -1, -1)) if direction == 'bwd' else list(range(t_max))
    k, b, z = draws[0].shape
    eps = torch.empty(t_max, b, k, z)
    for i, t in enumerate(order):
        eps[t] = draws[i].permute(1, 0, 2)
    return eps.cuda()


@pytest.mark.parametrize('name', ['spirals_ragged', 'gauss3_ffilter', 'bsmooth_full', 'medium_dims'])
def test_forward_modes_match_reference_golden(name):
    fx = load_golden(name)
    m = build(fx).eval()
    t_max = max(fx['lengths'])
    failures = []
    for key, ref in fx['ref_forward_fp32'].items():
        mode, sample, kf = key.split('/')
        sample, kf = bool(int(sample)), int(kf)
        draws = list(ref['draws'])
        flt_dir = 'fwd' if mode in ('ffilter', 'bsmooth') else 'bwd'
        eps_flt = eps_smt = None
        if sample or kf > 1:
            eps_flt = _assemble(draws[:t_max], t_max, flt_dir)
            draws = draws[t_max:]
        if mode in ('fsmooth', 'bsmooth') and sample:
            eps_smt = _assemble(draws[:t_max], t_max, 'fwd' if mode == 'fsmooth' else 'bwd')
        with torch.no_grad():
            infer, prior, recon = m(cuda(fx['inputs']), lengths=fx['lengths'], mode=mode, sample=sample,
                                    flt_particles=kf, smt_particles=1, noise=(eps_flt, eps_smt))
        names = ['infer_mean', 'infer_std', 'prior_mean', 'prior_std']
        pairs = list(zip(names, list(infer) + list(prior), ref['infer'] + ref['prior']))
        for mod in fx['modalities']:
            pairs += [('recon_' + mod, a, b) for a, b in zip(recon[mod], ref['recon'][mod])]
        for nm, a, b in pairs:
            a = a.cpu()
            bad = ~torch.isclose(a, b, rtol=2e-4, atol=2e-5, equal_nan=True)
            if bad.any():
                idx = bad.nonzero()[:4].tolist()
                failures.append((key, nm, int(bad.sum()), idx, [a[tuple(i)].item() for i in idx],
                                 [b[tuple(i)].item() for i in idx]))
    assert not failures, '\n'.join(str(f) for f in failures[:6])


def test_masks_bit_exact():
    fx = load_golden('spirals_half_missing')
    m = build(fx).eval()
    with torch.no_grad():
        _, _, masks = m.encode(cuda(fx['inputs']))
    for i, mod in enumerate(fx['modalities']):
        assert torch.equal(masks[i].cpu(), ~torch.isnan(fx['inputs'][mod]).any(dim=-1))


def test_composed_forward_is_differentiable_and_matches_oracle():
    """forward() -> loss() -> backward() through the op-level autograd Functions."""
    fx = load_golden('gauss3_ffilter')
    m = build(fx).train()
    t_max, b_dim, z = max(fx['lengths']), len(fx['lengths']), fx['z_dim']
    g = torch.Generator().manual_seed(5)
    k = 6
    eps_flt = torch.randn(t_max, b_dim, k, z, generator=g)
    eps_smt = torch.randn(t_max, b_dim, 1, z, generator=g)
    infer, prior, recon = m(cuda(fx['inputs']), lengths=fx['lengths'], mode='fsmooth',
                            flt_particles=k, noise=(eps_flt.cuda(), eps_smt.cuda()))
    loss = m.loss(cuda(fx['targets']), infer, prior, recon, fx['mask'].cuda(), 0.7, fx['rec_mults'])
    loss.backward()
    params = {kk: v.clone().double().requires_grad_(True) for kk, v in fx['state_dict'].items()}
    tape = [eps_flt[t].permute(1, 0, 2).contiguous() for t in range(t_max - 1, -1, -1)] + \
           [eps_smt[t].permute(1, 0, 2).contiguous() for t in range(t_max)]
    orc = bo.OracleDMM(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=z,
                       min_std=fx['min_std'], draw=bo.NoiseTape(tape))
    cast = lambda d: {kk: v.double() for kk, v in d.items()}
    inf_o, pri_o, rec_o = orc.forward(cast(fx['inputs']), fx['lengths'], mode='fsmooth', flt_particles=k)
    ref = orc.loss(cast(fx['targets']), inf_o, pri_o, rec_o, fx['mask'], 0.7, fx['rec_mults'])
    ref.backward()
    assert abs(loss.item() - ref.item()) / abs(ref.item()) < ELBO_TOL
    for kk, p in m.named_parameters():
        g_ref = params[kk].grad
        if g_ref is None or g_ref.norm() == 0:
            assert p.grad is None or p.grad.norm() == 0, kk
        else:
            assert rel_err(p.grad.cpu(), g_ref) < GRAD_TOL, (kk, rel_err(p.grad.cpu(), g_ref))


def test_losses_match_formulas():
    g = torch.Generator().manual_seed(0)
    m1, m2 = torch.randn(7, 5, 4, generator=g), torch.randn(7, 5, 4, generator=g)
    s1, s2 = torch.rand(7, 5, 4, generator=g) + 0.1, torch.rand(7, 5, 4, generator=g) + 0.1
    mask = torch.rand(7, 5, 1, generator=g) > 0.3
    x = torch.randn(7, 5, 4, generator=g)
    x[torch.rand(7, 5, 4, generator=g) < 0.2] = float('nan')
    cpu = [t.clone().double().requires_grad_(True) for t in (m1, s1, m2, s2)]
    gpu = [t.clone().cuda().requires_grad_(True) for t in (m1, s1, m2, s2)]
    ref = bo.kld_gauss(*cpu, mask) + bo.nll_gauss(cpu[0], cpu[1], x.double(), mask)
    out = models.losses.kld_gauss(*gpu, mask.cuda()) + \
        models.losses.nll_gauss(gpu[0], gpu[1], x.cuda(), mask.cuda())
    ref.backward()
    out.backward()
    assert abs(out.item() - ref.item()) / abs(ref.item()) < 1e-5
    for a, b in zip(gpu, cpu):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-5


def test_trainer_contract_adam_steps():
    """The loop of trainer.py:225-252: step -> /= sum(lengths) -> backward -> clip ->
    Adam.step -> zero_grad; loss must go down and parameters must stay aliased to
    the flat buffer the kernels read."""
    fx = load_golden('spirals_ragged')
    m = build(fx).train()
    m.noise_seed = 123
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    inputs, targets, mask = cuda(fx['inputs']), cuda(fx['targets']), fx['mask'].cuda()
    hist = []
    for _ in range(25):
        b_loss = m.step(inputs, mask, 1.0, fx['rec_mults'], targets=targets, lengths=fx['lengths'])
        hist.append(b_loss.item())
        b_loss /= sum(fx['lengths'])
        b_loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 10.0)
        opt.step()
        opt.zero_grad()
    assert hist[-1] < hist[0]
    assert m.fused_step_available and m.last_launches > 0
    sd = m.state_dict()
    assert set(sd.keys()) == set(fx['state_dict'].keys())


def test_flat_adam_trains_like_torch_adam():
    """optim.FlatAdam (bfvi_adam_step: fused clip + Adam on the flat buffers) against
    torch.optim.Adam + clip_grad_norm_ over the same steps, noise and data."""
    from multimodal_dmm_b200 import optim
    fx = load_golden('spirals_half_missing')
    n = float(sum(fx['lengths']))

    def train(use_flat):
        m = build(fx).train()
        m.noise_seed = 5
        opt = optim.FlatAdam(m, lr=5e-3, weight_decay=1e-4, max_norm=1.0) if use_flat else \
            torch.optim.Adam(m.parameters(), lr=5e-3, weight_decay=1e-4)
        losses = []
        for _ in range(4):
            loss = m.step(cuda(fx['inputs']), fx['mask'].cuda(), fx['kld_mult'], fx['rec_mults'],
                          targets=cuda(fx['targets']), lengths=fx['lengths'])
            (loss / n).backward()
            if not use_flat:
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
            opt.zero_grad()
            losses.append(loss.item())
        return losses, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    l_flat, s_flat = train(True)
    l_ref, s_ref = train(False)
    assert l_flat[-1] < l_flat[0]
    for a, b in zip(l_flat, l_ref):
        assert abs(a - b) <= 1e-4 * abs(b), (l_flat, l_ref)
    for k in s_ref:
        assert torch.allclose(s_flat[k], s_ref[k], rtol=1e-3, atol=1e-5), k
