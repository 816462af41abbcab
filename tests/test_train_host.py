"""Host-side logic of the training step on CPU: the Keras-2 Nadam schedule, the oracle's Nadam restatement against
torch.optim.NAdam, and the data-parallel gradient exchange under gloo (world_size 2)."""
import os

import numpy as np

from dsen2_b200 import train
from oracle import train_oracle as to


def test_nadam_schedule_matches_keras_recurrence():
    opt = train.Nadam()
    assert (opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, opt.schedule_decay) == (1e-4, 0.9, 0.999, 1e-8, 0.004)
    ms = 1.0
    for t in range(1, 50):
        s = train.nadam_schedule(t, ms, opt)
        mu_t = 0.9 * (1.0 - 0.5 * 0.96 ** (t * 0.004))
        mu_n = 0.9 * (1.0 - 0.5 * 0.96 ** ((t + 1) * 0.004))
        assert s['mu_t'] == mu_t and s['mu_next'] == mu_n
        assert np.isclose(s['sched_new'], ms * mu_t) and np.isclose(s['sched_next'], ms * mu_t * mu_n)
        assert np.isclose(s['bias2'], 1.0 - 0.999 ** t)
        ms = s['sched_new']


def test_oracle_nadam_equals_torch_nadam():
    import torch
    rng = np.random.RandomState(0)
    p = rng.randn(500)
    pt = torch.tensor(p.copy(), requires_grad=True)
    opt = torch.optim.NAdam([pt], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, momentum_decay=0.004)
    m = np.zeros_like(p); v = np.zeros_like(p); ms = 1.0
    for t in range(1, 8):
        g = rng.randn(500)
        pt.grad = torch.tensor(g.copy())
        opt.step()
        p, m, v, ms = to.nadam_reference(p, g, m, v, t, ms, lr=1e-3)
        np.testing.assert_allclose(p, pt.detach().numpy(), rtol=0, atol=1e-10)


def test_oracle_training_reduces_the_loss():
    rng = np.random.RandomState(1)
    from oracle import dsen2net_oracle as no
    w = no.he_uniform_weights(10, 6, 1, 8, seed=0)
    xs = [rng.rand(2, 4, 8, 8).astype(np.float32), rng.rand(2, 6, 8, 8).astype(np.float32)]
    y = (xs[1] + 0.1).astype(np.float32)
    losses, _ = to.train_steps([(xs, y)] * 20, w, lr=1e-2)
    assert losses[-1] < 0.5 * losses[0]


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    w = train.allreduce_gradients(flat)
    ok = w == world and torch.equal(flat, torch.arange(1000, dtype=torch.float32) * sum(range(1, world + 1)))
    # averaged gradient = what Trainer.apply_gradients(1/world) hands to the Nadam kernel
    avg = flat / w
    ok = ok and torch.allclose(avg, torch.arange(1000, dtype=torch.float32) * (world + 1) / 2)
    if rank == 0:
        np.save(os.path.join(tmp, 'ok.npy'), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(str(tmp_path / 'ok.npy'))[0]


def test_allreduce_is_identity_without_a_process_group():
    import torch
    flat = torch.ones(10)
    assert train.allreduce_gradients(flat) == 1 and torch.equal(flat, torch.ones(10))
