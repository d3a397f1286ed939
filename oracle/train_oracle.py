"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the training-side arithmetic the reference delegates to tianshou.

PARITY UNPINNED at the tianshou boundary: tianshou 1.0.0 (requirements.txt:5) is not installable in this image and the
reference has no test or golden vector for its training step, so these functions restate the PUBLISHED algorithm of

* ``tianshou.policy.base._nstep_return`` / ``BasePolicy.compute_nstep_return`` (called by ``DQNPolicy.process_fn``
  with ``discount_factor`` and ``estimation_step``, reference l_dgn.py:69-76):
      indices = [idx, next(idx), ...] (n_step entries; next(i) = i on an episode end),
      terminal = indices[-1]; target_q = target_q_fn(terminal) * ~terminated[terminal];
      backward recursion  returns = rew[now] + gamma * returns, reset to 0 / gammas = n+1 where end_flag[now];
      result = target_q * gamma ** gammas + returns
* ``torch.optim.Adam`` (single tensor, amsgrad = False), reference l_dgn.py:66 -- this part IS pinned: torch is installed,
  tests/test_training_host.py::test_adam_restatement_is_pinned_on_torch_optim_adam runs the real optimiser beside it
* ``tianshou DQNPolicy.learn`` loss  mean((returns - Q(obs)[act])^2)  and ``DGNPolicy.learn`` (policies/dgn.py:22-71).

Not product code: melissa_b200 never imports this module.
"""
from __future__ import annotations

import numpy as np


def nstep_return_chain(rew: np.ndarray, terminated: np.ndarray, start: int, n_step: int, gamma: float, target_q_fn):
    """One agent's transition chain (its tianshou sub-buffer, in order): rew[t], terminated[t] (== done here: the
    environment only ever sets ``terminations``, graph.py:330-334).  Returns the n-step return of transition
    ``start`` exactly as tianshou computes it; ``target_q_fn(t)`` values the observation AFTER transition t."""
    L = len(rew)
    nxt = lambda i: i if (terminated[i] or i + 1 >= L) else i + 1
    indices = [start]
    for _ in range(n_step - 1):
        indices.append(nxt(indices[-1]))
    terminal = indices[-1]
    end_flag = terminated.copy()
    end_flag[L - 1] = True                                   # unfinished_index(): the last stored transition
    tq = 0.0 if terminated[terminal] else float(target_q_fn(terminal))
    gammas, returns = n_step, 0.0
    for n in range(n_step - 1, -1, -1):
        now = indices[n]
        if end_flag[now]:
            gammas, returns = n + 1, 0.0
        returns = float(rew[now]) + gamma * returns
    return tq * gamma ** gammas + returns


def adam_reference(p, grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam over a list of gradient arrays applied in order; float64 numpy."""
    p = np.array(p, dtype=np.float64)
    m, v = np.zeros_like(p), np.zeros_like(p)
    for t, g in enumerate(grads, start=1):
        g = np.asarray(g, dtype=np.float64) + weight_decay * p
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p
