"""CPU: hand-derived checks of oracle/net_oracle.py (parity with PyG/tianshou is UNPINNED:
those packages are not installable here; see the module header)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import net_oracle as no


def _loop_gatv2(sd, p, x, edges, H):
    """Edge-list / per-target formulation written independently of the dense one."""
    N, C = x.shape[0], sd[f"{p}.att"].shape[-1]
    xl = (x @ sd[f"{p}.lin_l.weight"].T + sd[f"{p}.lin_l.bias"]).view(N, H, C)
    xr = (x @ sd[f"{p}.lin_r.weight"].T + sd[f"{p}.lin_r.bias"]).view(N, H, C)
    att = sd[f"{p}.att"].view(H, C)
    out = torch.zeros(N, H, C, dtype=x.dtype)
    for i in range(N):
        src = [j for j in range(N) if edges[i, j]] + [i]
        for h in range(H):
            e = torch.stack([(att[h] * torch.nn.functional.leaky_relu(xl[j, h] + xr[i, h], 0.2)).sum() for j in src])
            a = (e - e.max()).exp()
            a = a / (a.sum() + 1e-16)
            for w, j in zip(a, src):
                out[i, h] += w * xl[j, h]
    return out.reshape(N, H * C) + sd[f"{p}.bias"]


def _loop_transformer(sd, p, x, edges, H):
    N = x.shape[0]
    q = (x @ sd[f"{p}.lin_query.weight"].T + sd[f"{p}.lin_query.bias"]).view(N, H, -1)
    k = (x @ sd[f"{p}.lin_key.weight"].T + sd[f"{p}.lin_key.bias"]).view(N, H, -1)
    v = (x @ sd[f"{p}.lin_value.weight"].T + sd[f"{p}.lin_value.bias"]).view(N, H, -1)
    C = q.shape[-1]
    out = torch.zeros(N, H, C, dtype=x.dtype)
    for i in range(N):
        src = [j for j in range(N) if edges[i, j]]
        if not src:
            continue
        for h in range(H):
            e = torch.stack([(q[i, h] * k[j, h]).sum() / math.sqrt(C) for j in src])
            a = (e - e.max()).exp()
            a = a / (a.sum() + 1e-16)
            for w, j in zip(a, src):
                out[i, h] += w * v[j, h]
    return out.reshape(N, H * C)


def test_dense_convs_equal_edge_list_form():
    torch.manual_seed(0)
    N, H = 7, 4
    edges = torch.rand(N, N) < 0.4
    edges = (edges | edges.T) & ~torch.eye(N, dtype=torch.bool)
    edges[3] = False
    edges[:, 3] = False                       # isolated node
    x = torch.randn(N, 128, dtype=torch.float64)
    sd = {k: v.double() for k, v in no.init_state_dict("l_dgn", seed=1).items()}
    sd["conv1.bias"] = torch.randn(512, dtype=torch.float64)
    got = no.gatv2_conv(sd, "conv1", x[None], edges[None], H)[0]
    torch.testing.assert_close(got, _loop_gatv2(sd, "conv1", x, edges, H), rtol=1e-10, atol=1e-12)
    sd = {k: v.double() for k, v in no.init_state_dict("dgn_r", seed=2).items()}
    got = no.transformer_conv(sd, "conv1", x[None], edges[None], H)[0]
    torch.testing.assert_close(got, _loop_transformer(sd, "conv1", x, edges, H), rtol=1e-10, atol=1e-12)
    assert torch.all(got[3] == 0)             # isolated node -> 0 (no self loops in TransformerConv)


def test_radius_graph_mask_semantics():
    pos = torch.tensor([[[0.0, 0.0], [0.19, 0.0], [0.2, 0.0], [0.1, 0.1], [1.0, 1.0]]])
    m = no.radius_graph_mask(pos)[0]
    assert m[0].tolist() == [False, True, False, True, False]      # strict <, no self loop
    assert m.equal(m.T)
    assert not m[4].any()
    # 40 coincident nodes: every target keeps the first 33 hits in index order (self included), then drops self
    pos = torch.zeros(1, 40, 2)
    m = no.radius_graph_mask(pos)[0]
    assert m[0].nonzero().flatten().tolist() == list(range(1, 33))
    assert m[39].nonzero().flatten().tolist() == list(range(0, 33))
    assert m[10].nonzero().flatten().tolist() == [j for j in range(0, 33) if j != 10]
    assert abs(no.r2_threshold() - 0.04) < 1e-8


def test_forward_shapes_and_snapshot_order():
    N = 6
    torch.manual_seed(3)
    obs = torch.rand(5, N * 8 + 1)
    obs[:, 7:N * 8:8] = (torch.rand(5, N) < 0.7).float()
    obs[:, -1] = torch.tensor([0, 5, 2, 9, -3])                     # clamped to [0, N-1]
    for kind in ("l_dgn", "dgn_r", "hl_dgn"):
        sd = no.init_state_dict(kind, seed=4)
        q = no.FORWARDS[kind](sd, obs, N)
        assert q.shape == (5, 2) and torch.isfinite(q).all()
    with pytest.raises(ValueError):
        no.l_dgn_forward(no.init_state_dict("l_dgn"), torch.rand(2, N * 8), N)
    with pytest.raises(ValueError):
        no.l_dgn_forward(no.init_state_dict("l_dgn"), torch.rand(N * 8 + 1), N)
    # HL-DGN does not depend on the controlling index
    sd = no.init_state_dict("hl_dgn", seed=5)
    o2 = obs.clone()
    o2[:, -1] = 1
    torch.testing.assert_close(no.hl_dgn_forward(sd, obs, N, aggregator="max"), no.hl_dgn_forward(sd, o2, N, aggregator="max"))


def test_forward_graphs_equals_per_agent_rows():
    N, B = 6, 3
    torch.manual_seed(6)
    om = torch.rand(B, N, 8)
    cm = torch.rand(B, N) < 0.5
    sd = no.init_state_dict("l_dgn", seed=7)
    q = no.forward_graphs("l_dgn", sd, om, cm, N)
    for b in range(B):
        for i in range(N):
            if cm[b, i]:
                row = torch.cat([om[b].reshape(-1), torch.tensor([float(i)])])[None]
                torch.testing.assert_close(q[b, i], no.l_dgn_forward(sd, row, N)[0])
            else:
                assert torch.all(q[b, i] == 0)


def test_dqn_act_and_exploration_noise():
    q = np.array([[0.1, 0.2], [0.3, 0.3], [0.5, -1.0]], dtype=np.float32)
    act = no.dqn_act(q, np.ones((3, 2)))
    assert act.tolist() == [1, 0, 0]
    u_eps = np.array([0.01, 0.9, 0.04])
    u_act = np.array([[0.2, 0.7], [0.9, 0.1], [0.6, 0.5]])
    out = no.exploration_noise(act, 0.05, u_eps, u_act, np.ones((3, 2)))
    assert out.tolist() == [1, 0, 0]
    out = no.exploration_noise(act, 0.05, np.array([0.01, 0.9, 0.04]), np.array([[0.8, 0.7], [0.9, 0.1], [0.4, 0.5]]))
    assert out.tolist() == [0, 0, 1]
    assert no.exploration_noise(act, 0.0, u_eps, u_act).tolist() == act.tolist()


def test_transformer_conv_equals_torch_scaled_dot_product_attention():
    """An implementation maintained by somebody else for the one conv that has one in torch itself: PyG's
    TransformerConv(root_weight=False, beta=False) is multi-head scaled dot-product attention of every target over its
    incoming edges -- torch.nn.functional.scaled_dot_product_attention with the edge mask (targets with at least one
    edge; PyG's softmax adds 1e-16 to the denominator, far below fp32 resolution here)."""
    torch.manual_seed(4)
    B, N, D, H, C = 3, 14, 32, 4, 16
    sd = {}
    for lin in ("lin_query", "lin_key", "lin_value"):
        sd[f"c.{lin}.weight"] = torch.randn(H * C, D) / math.sqrt(D)
        sd[f"c.{lin}.bias"] = torch.randn(H * C) * 0.1
    x = torch.randn(B, N, D)
    edges = torch.rand(B, N, N) < 0.3
    edges &= ~torch.eye(N, dtype=torch.bool)
    edges[:, :, 0] |= torch.arange(N)[None, :] != 0           # every target but node 0 has an edge
    edges[:, 0, 1] = True
    got = no.transformer_conv(sd, "c", x, edges, H).view(B, N, H, C)
    q = F.linear(x, sd["c.lin_query.weight"], sd["c.lin_query.bias"]).view(B, N, H, C).transpose(1, 2)
    k = F.linear(x, sd["c.lin_key.weight"], sd["c.lin_key.bias"]).view(B, N, H, C).transpose(1, 2)
    v = F.linear(x, sd["c.lin_value.weight"], sd["c.lin_value.bias"]).view(B, N, H, C).transpose(1, 2)
    want = F.scaled_dot_product_attention(q, k, v, attn_mask=edges[:, None, :, :]).transpose(1, 2)
    assert float((got - want).abs().max()) < 1e-5


def test_net_oracle_against_torch_geometric_when_installed():
    """Auto-pinning: wherever torch_geometric (and torch_cluster for radius_graph) can be imported, the oracle's convs
    and edge rule are compared with the real GATv2Conv / TransformerConv / radius_graph.  Not installable in the build
    image (no index access): skipped there, and the network half of the oracle stays 'parity unpinned'."""
    tg = pytest.importorskip("torch_geometric")
    from torch_geometric.nn import GATv2Conv, TransformerConv
    torch.manual_seed(7)
    B, N, D, H, C = 2, 11, 24, 4, 16
    x = torch.randn(B, N, D)
    edges = torch.rand(B, N, N) < 0.35
    edges &= ~torch.eye(N, dtype=torch.bool)
    bi, ti, sj = torch.nonzero(edges, as_tuple=True)                          # edges[b, target i, source j]
    edge_index = torch.stack([bi * N + sj, bi * N + ti])                      # PyG: row 0 = source, row 1 = target
    xf = x.reshape(B * N, D)
    gat = GATv2Conv(D, C, heads=H)
    sd = {f"c.{k}": v.detach().clone() for k, v in gat.state_dict().items()}
    want = gat(xf, edge_index).view(B, N, H * C)
    assert float((no.gatv2_conv(sd, "c", x, edges, H) - want).abs().max()) < 1e-5
    tr = TransformerConv(D, C, heads=H, root_weight=False)
    sd = {f"c.{k}": v.detach().clone() for k, v in tr.state_dict().items()}
    want = tr(xf, edge_index).view(B, N, H * C)
    assert float((no.transformer_conv(sd, "c", x, edges, H) - want).abs().max()) < 1e-5
    try:
        from torch_geometric.nn import radius_graph
        pos = torch.rand(B, N, 2) * 0.5
        ei = radius_graph(pos.reshape(-1, 2), r=no.RADIUS, batch=torch.arange(B).repeat_interleave(N), loop=False, max_num_neighbors=32)
        got = torch.zeros(B * N, B * N, dtype=torch.bool)
        got[ei[1], ei[0]] = True
        mask = no.radius_graph_mask(pos)
        for b in range(B):
            assert torch.equal(got[b * N:(b + 1) * N, b * N:(b + 1) * N], mask[b])
    except ImportError:                                                       # torch_cluster missing: the conv checks above still count
        pass
