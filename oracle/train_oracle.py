"""fp32 CPU restatement of one DSen2 training step.  TEST INFRASTRUCTURE.  PARITY UNPINNED.

Follows ``/root/reference/training/supres_train.py:137-144`` (Nadam(lr, 0.9, 0.999, epsilon=1e-8,
schedule_decay=0.004); loss mean_absolute_error; metric mean_squared_error) and ``:218-230`` (model.fit inner
step) on the graph of ``utils/DSen2Net.py:9-43``.  Keras 2.x's Nadam is third-party code that is not in the
reference tree and not installable here; its recurrence (momentum schedule mu_t = beta_1 (1 - 0.5 * 0.96^(t *
schedule_decay))) is restated in ``nadam_reference`` below and cross-checked against ``torch.optim.NAdam(
momentum_decay=0.004)``, which implements the same update.  No reference artefact pins it: parity unpinned.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _params(weights):
    ps = []
    for k, b in weights:
        ps.append(torch.tensor(np.ascontiguousarray(k.transpose(3, 2, 0, 1)), requires_grad=True))
        ps.append(torch.tensor(b.copy(), requires_grad=True))
    return ps


def forward(xs, ps, scale=0.1):
    x = torch.cat(xs, dim=1)
    x = torch.relu(F.conv2d(x, ps[0], ps[1], padding=1))
    n_res = (len(ps) // 2 - 2) // 2
    for l in range(n_res):
        t = torch.relu(F.conv2d(x, ps[2 + 4 * l], ps[3 + 4 * l], padding=1))
        t = F.conv2d(t, ps[4 + 4 * l], ps[5 + 4 * l], padding=1)
        x = x + t * scale
    return F.conv2d(x, ps[-2], ps[-1], padding=1) + xs[-1]


def loss_and_grads(inputs, y, weights):
    """-> loss, mse, [(dkernel HWIO, dbias)]"""
    ps = _params(weights)
    xs = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in inputs]
    pred = forward(xs, ps)
    d = pred - torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
    loss = d.abs().mean()
    loss.backward()
    grads = [(ps[2 * i].grad.numpy().transpose(2, 3, 1, 0).copy(), ps[2 * i + 1].grad.numpy().copy())
             for i in range(len(ps) // 2)]
    return float(loss.detach()), float((d.detach() ** 2).mean()), grads


def nadam_reference(p, g, m, v, t, m_schedule, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, decay=0.004):
    """Keras-2 ``Nadam.get_updates`` for one tensor at (1-based) iteration t; returns (p, m, v, m_schedule_new)."""
    mu_t = b1 * (1.0 - 0.5 * 0.96 ** (t * decay))
    mu_n = b1 * (1.0 - 0.5 * 0.96 ** ((t + 1) * decay))
    s_new = m_schedule * mu_t
    s_next = s_new * mu_n
    g_prime = g / (1.0 - s_new)
    m_t = b1 * m + (1.0 - b1) * g
    m_prime = m_t / (1.0 - s_next)
    v_t = b2 * v + (1.0 - b2) * g * g
    v_prime = v_t / (1.0 - b2 ** t)
    m_bar = (1.0 - mu_t) * g_prime + mu_n * m_prime
    return p - lr * m_bar / (np.sqrt(v_prime) + eps), m_t, v_t, s_new


def train_steps(batches, weights, lr=1e-4, steps=None):
    """Run Nadam steps on [(inputs, y)]; returns (losses, final weights [(kernel HWIO, bias)])."""
    ps = _params(weights)
    opt = torch.optim.NAdam(ps, lr=lr, betas=(0.9, 0.999), eps=1e-8, momentum_decay=0.004)
    losses = []
    for inputs, y in batches[:steps]:
        xs = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in inputs]
        opt.zero_grad()
        d = forward(xs, ps) - torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
        loss = d.abs().mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    out = [(ps[2 * i].detach().numpy().transpose(2, 3, 1, 0).copy(), ps[2 * i + 1].detach().numpy().copy())
           for i in range(len(ps) // 2)]
    return losses, out
