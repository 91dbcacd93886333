"""Oracle: PaiNN message / update blocks.  Test infrastructure only.

Restates hydragnn/models/PAINNStack.py:194-328 with the same sub-module names
(``scalar_message_mlp``, ``filter_layer``, ``edge_filter``, ``update_U``,
``update_V``, ``update_mlp``).
"""
import torch
from torch import nn

from .geometry import cosine_cutoff, sinc_expansion


class PainnMessage(nn.Module):
    def __init__(self, node_size, num_radial, cutoff, edge_dim=None):
        super().__init__()
        self.node_size, self.num_radial, self.cutoff, self.edge_dim = node_size, num_radial, cutoff, edge_dim
        self.scalar_message_mlp = nn.Sequential(
            nn.Linear(node_size, node_size), nn.SiLU(), nn.Linear(node_size, 3 * node_size))
        self.filter_layer = nn.Linear(num_radial, 3 * node_size)
        if edge_dim is not None:
            self.edge_filter = nn.Sequential(
                nn.Linear(edge_dim, node_size), nn.SiLU(), nn.Linear(node_size, 3 * node_size))

    def forward(self, s, v, edge, diff, dist, edge_attr=None):
        # edge is [E, 2]; messages are read at edge[:, 1] and summed into edge[:, 0] (:247-266)
        F = self.node_size
        W = self.filter_layer(sinc_expansion(dist, self.num_radial, self.cutoff))   # :239-241
        W = W * cosine_cutoff(dist, self.cutoff)                                    # :242
        if edge_attr is not None:
            W = W * self.edge_filter(edge_attr)                                     # :243-244
        phi = self.scalar_message_mlp(s)                                            # :246
        f = W * phi[edge[:, 1]]
        g_v, g_e, m_s = torch.split(f, F, dim=1)
        # Q2: diff is already unit length and is divided by dist again (:257)
        m_v = v[edge[:, 1]] * g_v.unsqueeze(1) + g_e.unsqueeze(1) * (diff / dist).unsqueeze(-1)
        ds = torch.zeros_like(s).index_add_(0, edge[:, 0], m_s)
        dv = torch.zeros_like(v).index_add_(0, edge[:, 0], m_v)
        return s + ds, v + dv


class PainnUpdate(nn.Module):
    def __init__(self, node_size, last_layer=False):
        super().__init__()
        self.last_layer = last_layer
        self.update_U = nn.Linear(node_size, node_size)
        self.update_V = nn.Linear(node_size, node_size)
        self.update_mlp = nn.Sequential(
            nn.Linear(2 * node_size, node_size), nn.SiLU(),
            nn.Linear(node_size, (2 if last_layer else 3) * node_size))

    def forward(self, s, v):
        F = v.shape[-1]
        Uv, Vv = self.update_U(v), self.update_V(v)            # Linear WITH bias on [N,3,F] (:299-300)
        a = self.update_mlp(torch.cat([torch.linalg.norm(Vv, dim=1), s], dim=1))
        inner = (Uv * Vv).sum(dim=1)
        if self.last_layer:                                    # :318-328 -- v is dropped
            a_sv, a_ss = torch.split(a, F, dim=1)
            return s + a_sv * inner + a_ss, None
        a_vv, a_sv, a_ss = torch.split(a, F, dim=1)
        return s + a_sv * inner + a_ss, v + a_vv.unsqueeze(1) * Uv
