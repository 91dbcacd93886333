"""Oracle: PNA-equivariant message / update blocks.  Test infrastructure only.

Restates hydragnn/models/PNAEqStack.py:240-538 (sub-module names ``pre_nns``, ``post_nns``, ``rbf_emb``,
``rbf_lin``, ``edge_encoder``, ``scalar_message_mlp``, ``update_X``, ``update_V``, ``update_mlp``,
``aggr_module``).  ``DegreeScalerAggregation`` is torch_geometric == 2.6.1 (requirements-pyg.txt:1), absent
from the reference tree; its published algorithm is restated in :class:`DegreeScalerAggregation`
[3P-memory, SURVEY Appendix B.1] -- the 20-way aggregation is therefore "parity unpinned" (the reference's own
equivariance test replaces it with mean/identity, tests/test_forces_equivariant.py:48-78).
"""
import math

import torch
from torch import nn

from .geometry import segment_sum


def sanitize_degree(deg):
    """``PNAEqStack._sanitize_degree`` (PNAEqStack.py:75-90)."""
    deg = torch.as_tensor(deg, dtype=torch.float32)
    if deg.numel() == 0:
        return deg.new_ones((1,))
    finite = torch.isfinite(deg)
    max_finite = deg[finite].max() if finite.any() else deg.new_tensor(1.0)
    deg = torch.nan_to_num(deg, nan=1.0, neginf=1.0, posinf=float(max_finite))
    return deg.clamp_min(1.0)


class DegreeScalerAggregation(nn.Module):
    """PyG 2.6.1 ``DegreeScalerAggregation(aggr, scaler, deg)`` (train_norm=False)."""

    def __init__(self, aggr, scaler, deg):
        super().__init__()
        self.aggr, self.scaler = list(aggr), list(scaler)
        deg = deg.to(torch.float)
        n = int(deg.sum())
        bins = torch.arange(deg.numel(), dtype=torch.float)
        self.register_buffer("avg_deg_lin", torch.tensor(float((bins * deg).sum()) / n))
        self.register_buffer("avg_deg_log", torch.tensor(float(((bins + 1).log() * deg).sum()) / n))

    def forward(self, x, index, dim_size):
        cnt = torch.bincount(index, minlength=dim_size).to(x.dtype)
        cnt1 = cnt.clamp(min=1)[:, None]
        idx = index[:, None].expand_as(x)
        outs = []
        for a in self.aggr:
            if a == "mean":
                outs.append(segment_sum(x, index, dim_size) / cnt1)
            elif a in ("min", "max"):
                red = "amin" if a == "min" else "amax"
                outs.append(x.new_zeros(dim_size, x.shape[1]).scatter_reduce(0, idx, x, reduce=red, include_self=False))
            elif a == "std":
                mean = segment_sum(x, index, dim_size) / cnt1
                mean2 = segment_sum(x * x, index, dim_size) / cnt1
                std = (mean2 - mean * mean).clamp(min=1e-5).sqrt()
                outs.append(std.masked_fill(std <= math.sqrt(1e-5), 0.0))
            else:
                raise ValueError(a)
        out = torch.cat(outs, dim=-1)
        deg = cnt.clamp(min=1)[:, None]
        res = []
        for s in self.scaler:
            if s == "identity":
                res.append(out)
            elif s == "amplification":
                res.append(out * (torch.log(deg + 1) / self.avg_deg_log))
            elif s == "attenuation":
                res.append(out * (self.avg_deg_log / torch.log(deg + 1)))
            elif s == "linear":
                res.append(out * (deg / self.avg_deg_lin))
            elif s == "inverse_linear":
                res.append(out * (self.avg_deg_lin / deg))
            else:
                raise ValueError(s)
        return torch.cat(res, dim=-1)


X_AGGREGATORS = ["mean", "min", "max", "std"]
X_SCALERS = ["identity", "amplification", "attenuation", "linear", "inverse_linear"]


def rbf_basis(dist, num_radial, cutoff):
    """``rbf_BasisLayer.forward`` (PNAEqStack.py:479-538): sinc expansion (with the d -> 0 guard) x cosine cutoff.
    ``dist`` is [E]."""
    n = torch.arange(1, num_radial + 1, device=dist.device, dtype=dist.dtype)
    d = dist.unsqueeze(-1)
    sinc = torch.sin(d * n * math.pi / cutoff) / d.clamp_min(1e-9)
    sinc = torch.where(d.abs() < 1e-9, (n * math.pi / cutoff).expand_as(sinc), sinc)
    fc = torch.where(dist < cutoff, 0.5 * (torch.cos(math.pi * dist / cutoff) + 1), torch.zeros_like(dist))
    return sinc * fc.unsqueeze(-1)


class PainnMessage(nn.Module):
    def __init__(self, node_size, deg, edge_dim, num_radial):
        super().__init__()
        F = node_size
        self.node_size, self.edge_dim, self.num_radial = F, edge_dim, num_radial
        self.F_in = self.F_out = F
        self.towers = 1
        self.aggr_module = DegreeScalerAggregation(X_AGGREGATORS, X_SCALERS, deg)
        self.pre_nns = nn.ModuleList([nn.Sequential(nn.Linear((4 if edge_dim else 3) * F, F))])
        self.post_nns = nn.ModuleList([nn.Sequential(nn.Linear((len(X_AGGREGATORS) * len(X_SCALERS) + 1) * F, F))])
        self.rbf_emb = nn.Sequential(nn.Linear(num_radial, F), nn.Tanh())
        if edge_dim is not None:
            self.edge_encoder = nn.Linear(edge_dim, F)
        self.rbf_lin = nn.Linear(num_radial, 3 * F, bias=False)
        self.scalar_message_mlp = nn.Sequential(nn.Linear(F, F), nn.Tanh(), nn.Linear(F, F), nn.SiLU(), nn.Linear(F, 3 * F))

    def forward(self, x, v, edge, edge_rbf, edge_vec, edge_attr=None):
        src, dst = edge[:, 0], edge[:, 1]                           # :341
        feats = [x[src], x[dst], self.rbf_emb(edge_rbf)]            # :349-366
        if edge_attr is not None:
            feats.append(self.edge_encoder(edge_attr))
        m = self.pre_nns[0](torch.cat(feats, dim=-1))
        f = self.scalar_message_mlp(m) * self.rbf_lin(edge_rbf)     # :372-381
        g_v, g_e, m_s = torch.split(f, self.node_size, dim=-1)
        m_v = v[dst] * g_v.unsqueeze(1) + g_e.unsqueeze(1) * edge_vec.unsqueeze(-1)   # :391-393 (vec NOT re-divided)
        agg = self.aggr_module(m_s, src, x.shape[0])                # :396-400
        dx = self.post_nns[0](torch.cat([x, agg], dim=-1))
        dv = torch.zeros_like(v).index_add_(0, src, m_v)
        return x + dx, v + dv


class PainnUpdate(nn.Module):
    def __init__(self, node_size, last_layer=False):
        super().__init__()
        self.last_layer = last_layer
        self.update_X = nn.Linear(node_size, node_size)
        self.update_V = nn.Linear(node_size, node_size)
        self.update_mlp = nn.Sequential(nn.Linear(2 * node_size, node_size), nn.SiLU(),
                                        nn.Linear(node_size, (2 if last_layer else 3) * node_size))

    def forward(self, x, v):
        F = v.shape[-1]
        Xv, Vv = self.update_X(v), self.update_V(v)
        a = self.update_mlp(torch.cat([torch.linalg.norm(Vv, dim=1), x], dim=-1))
        inner = (Xv * Vv).sum(dim=1)
        if self.last_layer:
            a_xv, a_xx = torch.split(a, F, dim=-1)
            return x + a_xv * inner + a_xx, None
        a_vv, a_xv, a_xx = torch.split(a, F, dim=-1)
        return x + a_xv * inner + a_xx, v + a_vv.unsqueeze(1) * Xv
