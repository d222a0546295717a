"""Trainer-level drop-in test (SURVEY.md §4 iii): the reference's OWN `SpiralsTrainer` (spirals.py:16-135,
trainer.py:155-330) runs with `models` swapped for this package on cuda:0, next to the same trainer with the
reference's `models` on the CPU — same seed, same generated spirals dataset, same command-line arguments.

 * identical seeded initialisation  -> the deterministic evaluation (MAP estimate, `flt_particles=1`) of the untrained
   models must agree: KLD / reconstruction loss / MSE per trainer.py:264-323, spirals.py:93-111;
 * one epoch of `Trainer.train` (burst deletions, KL annealing, clip, Adam) keeps them together: the sampling noise of
   the two implementations differs (Philox vs torch.randn), so the epoch loss is compared statistically;
 * the checkpoint written by one loads into the other (`state_dict` keys / shapes).

The reference is imported from its byte-compiled copy under oracle/_ref/ (oracle/build_ref.py), which travels to the
GPU box; the test skips when that copy does not exist.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu


def make_trainer(spirals, trainer, models_module, device, data_dir, save_dir, extra=()):
    trainer.models = models_module                       # trainer.py:21 `import models` is the whole seam
    argv = ['--method', 'bfvi', '--device', device, '--data_dir', data_dir, '--save_dir', save_dir,
            '--batch_size', '20', '--epochs', '1', '--eval_freq', '1', '--save_freq', '1', '--data_workers', '0',
            '--kld_anneal', '1', '--eval_args', '{flt_particles: 1}', '--seed', '3'] + list(extra)
    args = spirals.SpiralsTrainer.parser.parse_args(argv)
    return spirals.SpiralsTrainer(args), args


def evaluate(tr, trainer, args, seed=11):
    from torch.utils.data import DataLoader
    from datasets import multiseq as mseq               # the reference's (oracle/_ref/datasets)
    loader = DataLoader(tr.test_data, batch_size=args.batch_sz_eval, collate_fn=mseq.seq_collate_dict, shuffle=False)
    np.random.seed(seed)                                 # rand_delete draws from numpy (datasets/multiseq.py:405-426)
    with torch.no_grad():
        return tr.evaluate(loader, args)[1]


def test_reference_trainer_with_models_swapped(tmp_path):
    if ref_shim.reference_root() is None:
        pytest.skip('reference not available (oracle/_ref not built)')
    spirals, trainer = ref_shim.import_reference_trainer()
    ref_models = ref_shim.import_reference_models()
    import multimodal_dmm_b200.models as our_models
    from datasets import spirals as ref_spirals_data
    data_dir = str(tmp_path / 'spirals')
    np.random.seed(1)
    ref_spirals_data.gen_dataset(n_examples=60, n_train=40, timesteps=40, data_dir=data_dir)

    t_ref, a_ref = make_trainer(spirals, trainer, ref_models, 'cpu', data_dir, str(tmp_path / 'save_ref'))
    t_our, a_our = make_trainer(spirals, trainer, our_models, 'cuda:0', data_dir, str(tmp_path / 'save_our'))
    assert type(t_our.model).__module__.startswith('multimodal_dmm_b200')
    # same seed -> same initial parameters, key for key
    sd_ref, sd_our = t_ref.model.state_dict(), t_our.model.state_dict()
    assert list(sd_ref) == list(sd_our)
    for k in sd_ref:
        assert torch.equal(sd_ref[k].cpu(), sd_our[k].cpu()), k

    # ---- deterministic evaluation of the untrained models ----
    m_ref, m_our = evaluate(t_ref, trainer, a_ref), evaluate(t_our, trainer, a_our)
    for k in ('kld_loss', 'rec_loss', 'mse', 'mse_std'):
        assert abs(m_our[k] - m_ref[k]) <= 2e-4 * abs(m_ref[k]) + 1e-6, (k, m_our[k], m_ref[k])

    # ---- one epoch of Trainer.train through run_train (train, evaluate, checkpoints, param_hist.tsv) ----
    for tr, args in ((t_ref, a_ref), (t_our, a_our)):
        torch.manual_seed(5)
        np.random.seed(5)
        tr.run_train(args)
        assert os.path.exists(os.path.join(args.save_dir, 'last.pth'))
        assert os.path.exists(os.path.join(args.save_dir, 'best.pth'))
    e_ref, e_our = evaluate(t_ref, trainer, a_ref), evaluate(t_our, trainer, a_our)
    for k in ('kld_loss', 'rec_loss', 'mse'):
        assert np.isfinite(e_our[k])
        assert abs(e_our[k] - e_ref[k]) <= 0.05 * abs(e_ref[k]) + 1e-3, (k, e_our[k], e_ref[k])
    # training moved the parameters the same way (Adam steps of lr 1e-4 on noisy gradients: compare the update)
    for k in sd_ref:
        d_ref = t_ref.model.state_dict()[k].cpu() - sd_ref[k].cpu()
        d_our = t_our.model.state_dict()[k].cpu() - sd_our[k].cpu()
        if d_ref.norm() > 0:
            cos = (d_ref * d_our).sum() / (d_ref.norm() * d_our.norm() + 1e-30)
            assert cos > 0.5, (k, cos.item())

    # ---- checkpoints are interchangeable ----
    ck = torch.load(os.path.join(a_our.save_dir, 'last.pth'), map_location='cpu')
    t_ref.model.load_state_dict(ck['model'])
    ck = torch.load(os.path.join(a_ref.save_dir, 'last.pth'), map_location='cuda:0')
    t_our.model.load_state_dict(ck['model'])
    x_ref, x_our = evaluate(t_ref, trainer, a_ref), evaluate(t_our, trainer, a_our)   # models swapped now
    assert abs(x_ref['mse'] - e_our['mse']) <= 2e-4 * abs(e_our['mse']) + 1e-6
    assert abs(x_our['mse'] - e_ref['mse']) <= 2e-4 * abs(e_ref['mse']) + 1e-6
