"""TEST INFRASTRUCTURE ONLY -- pure-torch restatement of the reference Q-networks.

Reference side (all under graph_env/env/utils/networks/):
  common.py:6-64    build_pyg_batch_time   -> :func:`split_obs`, :func:`radius_graph_mask`
  dgn_r.py:82-129   DGNRNetwork.forward    -> :func:`dgn_r_forward`
  l_dgn.py:92-151   LDGNNetwork.forward    -> :func:`l_dgn_forward`
  hl_dgn.py:82-119  HLDGNNetwork.forward   -> :func:`hl_dgn_forward`
Third-party semantics restated (packages NOT installed here, pins from the reference's
requirements.txt): torch-geometric~=2.2.0 ``GATv2Conv`` / ``TransformerConv`` /
``softmax`` / ``global_*_pool`` / ``radius_graph`` (torch_cluster ``radius`` CUDA kernel),
tianshou==1.0.0 ``MLP`` and ``DQNPolicy.forward/compute_q_value/exploration_noise``.

PARITY UNPINNED at that boundary: the reference has no network test or golden vector,
and PyG/tianshou cannot be installed offline, so this file follows their published
semantics.  Hand-derived small cases are in
tests/test_oracle_net.py.

Graphs are handled densely ([B, N, N] masks) -- mathematically the scatter/gather form
PyG uses, evaluated per target node.  Parameters come in as a state_dict with the
reference's key names.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

RADIUS = 0.20
MAX_NUM_NEIGHBORS = 32          # torch_cluster radius_graph default


def r2_threshold() -> float:
    """torch_cluster passes ``r * r`` computed in double and narrows it to the tensor dtype."""
    return float(np.float32(np.float64(RADIUS) * np.float64(RADIUS)))


def split_obs(obs: torch.Tensor, n_agents: int, input_dim: int = 5):
    """common.py:20-44,63.  obs [bs, N*(input_dim+3)+1] -> pos [bs,N,2], feats [bs,N,F],
    dm [bs,N,1], ctrl [bs] (clamped)."""
    if obs.ndim != 2:
        raise ValueError(f"Expected obs to be 2D, but got shape {obs.shape}")
    bs, dim = obs.shape
    d = input_dim + 2 + 1
    if dim - 1 != n_agents * d:
        raise ValueError(f"Expected {n_agents * d} feature cols for nodes, got {dim - 1}")
    node = obs[:, : dim - 1].reshape(bs, n_agents, d).float()
    pos, feats, dm = node[..., :2], node[..., 2:-1], node[..., -1:]
    ctrl = obs[:, -1].clamp(0, n_agents - 1).long()
    return pos, feats, dm, ctrl


def radius_graph_mask(pos: torch.Tensor) -> torch.Tensor:
    """torch_cluster ``radius_graph(pos, batch, r=0.2, loop=False, max_num_neighbors=32)``.
    Returns bool [B, N(target i), N(source j)].  For each target i candidates j of the same
    graph are scanned in index order, kept while ``dist < r*r`` in fp32 with
    ``dist = fma(dy, dy, dx*dx)`` (nvcc contracts the accumulation), the scan stops after
    33 hits (self included), then self is dropped."""
    p = pos.to(torch.float32)
    dx = (p[:, None, :, 0] - p[:, :, None, 0])          # x_j - y_i  -> [B, i, j]
    dy = (p[:, None, :, 1] - p[:, :, None, 1])
    dx2 = (dx * dx)                                     # fp32 rounding of dx*dx (fma(dx,dx,0))
    d2 = (dy.double() * dy.double() + dx2.double()).float()   # fma(dy,dy,dx2): one rounding
    hit = d2 < r2_threshold()
    rank = hit.cumsum(dim=2)                            # 1-based rank in index order, self included
    hit = hit & (rank <= MAX_NUM_NEIGHBORS + 1)
    n = p.shape[1]
    eye = torch.eye(n, dtype=torch.bool, device=p.device)
    return hit & ~eye


def mlp(sd, prefix, x, n_layers):
    """tianshou MLP: Sequential(Linear, ReLU, ..., Linear) under ``{prefix}.model.{0,2,4}``."""
    for k in range(n_layers):
        x = F.linear(x, sd[f"{prefix}.model.{2 * k}.weight"], sd[f"{prefix}.model.{2 * k}.bias"])
        if k + 1 < n_layers:
            x = F.relu(x)
    return x


def _masked_softmax(e, mask):
    """PyG utils.softmax per target: exp(e - max) / (sum + 1e-16); e [B,i,j,H], mask [B,i,j]."""
    m = mask[..., None]
    neg = torch.finfo(e.dtype).min
    mx = torch.where(m, e, torch.full_like(e, neg)).amax(dim=2, keepdim=True)
    mx = torch.where(m.any(dim=2, keepdim=True), mx, torch.zeros_like(mx))
    ex = torch.where(m, (e - mx).exp(), torch.zeros_like(e))
    return ex / (ex.sum(dim=2, keepdim=True) + 1e-16)


def gatv2_conv(sd, prefix, x, edges, heads):
    """PyG GATv2Conv(concat=True, negative_slope=0.2, add_self_loops=True, bias=True,
    share_weights=False).  x [B,N,D]; edges bool [B,i,j] (j -> i, no self loops)."""
    B, N, _ = x.shape
    xl = F.linear(x, sd[f"{prefix}.lin_l.weight"], sd[f"{prefix}.lin_l.bias"]).view(B, N, heads, -1)
    xr = F.linear(x, sd[f"{prefix}.lin_r.weight"], sd[f"{prefix}.lin_r.bias"]).view(B, N, heads, -1)
    att = sd[f"{prefix}.att"].view(heads, -1)
    mask = edges | torch.eye(N, dtype=torch.bool, device=x.device)
    s = F.leaky_relu(xl[:, None, :, :, :] + xr[:, :, None, :, :], 0.2)       # [B,i,j,H,C]
    e = (s * att).sum(-1)                                                    # [B,i,j,H]
    a = _masked_softmax(e, mask)
    out = torch.einsum("bijh,bjhc->bihc", a, xl).reshape(B, N, -1)
    return out + sd[f"{prefix}.bias"]


def transformer_conv(sd, prefix, x, edges, heads):
    """PyG TransformerConv(concat=True, beta=False, root_weight=False, bias=True): no self
    loops, no skip (``lin_skip`` exists in the state_dict but is unused), isolated node -> 0."""
    B, N, _ = x.shape
    q = F.linear(x, sd[f"{prefix}.lin_query.weight"], sd[f"{prefix}.lin_query.bias"]).view(B, N, heads, -1)
    k = F.linear(x, sd[f"{prefix}.lin_key.weight"], sd[f"{prefix}.lin_key.bias"]).view(B, N, heads, -1)
    v = F.linear(x, sd[f"{prefix}.lin_value.weight"], sd[f"{prefix}.lin_value.bias"]).view(B, N, heads, -1)
    c = q.shape[-1]
    e = torch.einsum("bihc,bjhc->bijh", q, k) / math.sqrt(c)
    a = _masked_softmax(e, edges)
    return torch.einsum("bijh,bjhc->bihc", a, v).reshape(B, N, -1)


def _dueling(sd, z):
    if "out_linear.weight" in sd:                # network built without dueling_param (l_dgn.py:88-90,149)
        return F.linear(z, sd["out_linear.weight"], sd["out_linear.bias"])
    q = mlp(sd, "Q", z, 3)
    v = mlp(sd, "V", z, 3)
    return q - q.mean(dim=1, keepdim=True) + v


def _encode(sd, feats):
    return F.relu(mlp(sd, "encoder", feats, 2))


def _gather(x, ctrl):
    return x[torch.arange(x.shape[0], device=x.device), ctrl]


def l_dgn_forward(sd, obs, n_agents, heads=4):
    pos, feats, dm, ctrl = split_obs(obs, n_agents)
    edges = radius_graph_mask(pos)
    x = _encode(sd, feats)
    x1 = _gather(x, ctrl)
    x = F.relu(gatv2_conv(sd, "conv1", x, edges, heads))
    x2 = _gather(x, ctrl)
    x = x * dm
    x = F.relu(gatv2_conv(sd, "conv2", x, edges, heads))
    x3 = _gather(x, ctrl)
    return _dueling(sd, torch.cat([x1, x2, x3], dim=1))


def dgn_r_forward(sd, obs, n_agents, heads=4):
    pos, feats, dm, ctrl = split_obs(obs, n_agents)
    edges = radius_graph_mask(pos)
    x = _encode(sd, feats)
    x1 = _gather(x, ctrl)
    x = F.relu(transformer_conv(sd, "conv1", x, edges, heads))
    x2 = _gather(x, ctrl)
    x = x * dm
    x = F.relu(transformer_conv(sd, "conv2", x, edges, heads))
    x3 = _gather(x, ctrl)
    return _dueling(sd, torch.cat([x1, x2, x3], dim=1))


def hl_dgn_forward(sd, obs, n_agents, heads=4, aggregator="mean"):
    pos, feats, dm, ctrl = split_obs(obs, n_agents)
    edges = radius_graph_mask(pos)
    x = _encode(sd, feats)
    x = F.relu(gatv2_conv(sd, "conv1", x, edges, heads))
    x = x * dm
    if aggregator == "mean":
        z = x.mean(dim=1)
    elif aggregator == "add":
        z = x.sum(dim=1)
    elif aggregator == "max":
        z = x.amax(dim=1)
    else:
        raise KeyError(aggregator)
    return _dueling(sd, z)


FORWARDS = {"l_dgn": l_dgn_forward, "dgn_r": dgn_r_forward, "hl_dgn": hl_dgn_forward}


def forward_graphs(kind, sd, obs_matrix, ctrl_mask, n_agents, **kw):
    """Batched-rollout view: obs_matrix [B,N,8], ctrl_mask bool [B,N] -> q [B,N,2] (rows of
    non-controlling nodes are 0).  Implemented by expanding to one agent-observation row per
    controlling node, exactly what the reference feeds its network."""
    B, N, _ = obs_matrix.shape
    b_idx, i_idx = torch.nonzero(ctrl_mask, as_tuple=True)
    flat = obs_matrix.reshape(B, N * 8)[b_idx]
    rows = torch.cat([flat, i_idx.to(flat.dtype)[:, None]], dim=1)
    q = torch.zeros(B, N, 2, dtype=torch.float32, device=obs_matrix.device)
    if rows.shape[0]:
        out = []
        for s in range(0, rows.shape[0], 256):
            out.append(FORWARDS[kind](sd, rows[s:s + 256], n_agents, **kw))
        q[b_idx, i_idx] = torch.cat(out).to(torch.float32)
    return q


# ----------------------------------------------------------------------------- policy
def dqn_act(q: np.ndarray, mask: np.ndarray | None = None):
    """tianshou DQNPolicy.compute_q_value + forward: logits + (1-mask)*(min-max-1); argmax."""
    logits = np.array(q, dtype=np.float32)
    if mask is not None:
        min_value = logits.min() - logits.max() - 1.0
        logits = logits + (1 - mask.astype(np.float32)) * min_value
    return logits.argmax(axis=1)


def exploration_noise(act: np.ndarray, eps: float, u_eps: np.ndarray, u_act: np.ndarray,
                      mask: np.ndarray | None = None):
    """tianshou DQNPolicy.exploration_noise with the uniforms supplied by the caller:
    ``rand_mask = u_eps < eps``; ``rand_act = argmax(u_act + mask)``; skipped when eps ~ 0."""
    act = act.copy()
    if np.isclose(eps, 0.0):
        return act
    rand_mask = u_eps < eps
    q = u_act.copy()
    if mask is not None:
        q = q + mask
    rand_act = q.argmax(axis=1)
    act[rand_mask] = rand_act[rand_mask]
    return act


# ------------------------------------------------------------------ random parameters
def init_state_dict(kind: str, seed: int = 9, hidden: int = 128, heads: int = 4, input_dim: int = 5,
                    dtype=torch.float32, dueling: bool = True):
    """Random weights with the reference's parameter names/shapes and
    the libraries' default initialisers (nn.Linear default for tianshou MLP and
    TransformerConv linears; glorot + zero bias for GATv2Conv)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def linear(name, out_f, in_f, glorot=False):
        if glorot:
            a = math.sqrt(6.0 / (in_f + out_f))
            sd[f"{name}.weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a
            sd[f"{name}.bias"] = torch.zeros(out_f)
        else:
            bound = 1.0 / math.sqrt(in_f)
            sd[f"{name}.weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
            sd[f"{name}.bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    hc = hidden * heads
    linear("encoder.model.0", hidden, input_dim)
    linear("encoder.model.2", hidden, hidden)
    convs = [("conv1", hidden)] + ([("conv2", hc)] if kind != "hl_dgn" else [])
    for name, din in convs:
        if kind == "dgn_r":
            for lin in ("lin_key", "lin_query", "lin_value", "lin_skip"):
                linear(f"{name}.{lin}", hc, din)
        else:
            a = math.sqrt(6.0 / (heads + hidden))
            sd[f"{name}.att"] = (torch.rand(1, heads, hidden, generator=g) * 2 - 1) * a
            sd[f"{name}.bias"] = torch.zeros(hc)
            linear(f"{name}.lin_l", hc, din, glorot=True)
            linear(f"{name}.lin_r", hc, din, glorot=True)
    latent = hc if kind == "hl_dgn" else hidden + 2 * hc
    if not dueling:
        linear("out_linear", 2, latent)
        return {k: v.to(dtype) for k, v in sd.items()}
    for head, out in (("Q", 2), ("V", 1)):
        linear(f"{head}.model.0", 128, latent)
        linear(f"{head}.model.2", 128, 128)
        linear(f"{head}.model.4", out, 128)
    return {k: v.to(dtype) for k, v in sd.items()}
