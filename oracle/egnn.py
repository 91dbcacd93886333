"""Oracle: EGNN ``E_GCL`` layer.  Test infrastructure only.

Restates hydragnn/models/EGCLStack.py:180-300 with the same sub-module names
(``edge_mlp``, ``node_mlp``, ``coord_mlp``) so state dicts interchange.
"""
import torch
from torch import nn

from .geometry import edge_vectors_and_lengths, segment_mean, segment_sum


class EGCL(nn.Module):
    def __init__(self, input_channels, output_channels, hidden_channels, edge_attr_dim=0,
                 equivariant=False):
        super().__init__()
        ed = edge_attr_dim or 0
        self.equivariant = bool(equivariant)
        # EGCLStack.py:208-213 -- activation is the hard-coded nn.ReLU default arg (:188)
        self.edge_mlp = nn.Sequential(
            nn.Linear(2 * input_channels + 1 + ed, hidden_channels), nn.ReLU(),
            nn.Linear(hidden_channels, hidden_channels), nn.ReLU())
        # :215-221
        self.node_mlp = nn.Sequential(
            nn.Linear(hidden_channels + input_channels, hidden_channels), nn.ReLU(),
            nn.Linear(hidden_channels, output_channels))
        if self.equivariant:  # :225-237
            last = nn.Linear(hidden_channels, 1, bias=False)
            nn.init.xavier_uniform_(last.weight, gain=0.001)
            self.coord_mlp = nn.Sequential(
                nn.Linear(hidden_channels, hidden_channels), nn.ReLU(), last, nn.Tanh())

    def forward(self, x, coord, edge_index, edge_attr=None, edge_shifts=None):
        row, col = edge_index[0], edge_index[1]
        # :280-282 -- normalised with eps = 1.0 (Q3); "radial" is the length, not its square
        coord_diff, radial = edge_vectors_and_lengths(coord, edge_index, edge_shifts,
                                                      normalize=True, eps=1.0)
        feats = [x[row], x[col], radial]
        if edge_attr is not None:
            feats.append(edge_attr)
        m = self.edge_mlp(torch.cat(feats, dim=1))                       # :245-250
        if self.equivariant:                                             # :268-276
            trans = torch.clamp(coord_diff * self.coord_mlp(m), min=-100, max=100)
            coord = coord + segment_mean(trans, row, coord.shape[0])
        agg = segment_sum(m, row, x.shape[0])                            # :257-258
        out = self.node_mlp(torch.cat([x, agg], dim=1))                  # :262-263
        return out, coord
