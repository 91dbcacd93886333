"""Oracle: GPS layer (local MPNN + dense multi-head self-attention).  Test infrastructure only.

Restates ``GPSConv`` (hydragnn/globalAtt/gps.py:32-152) for ``attn_type == "multihead"`` with the reference's
quirk Q1: ``graph_batch`` is never passed by ``Base`` (hydragnn/models/Base.py:709-713), so
``to_dense_batch(x, None)`` turns the WHOLE mini-batch into one sequence and attention mixes graphs.
Third-party pieces: torch ``nn.MultiheadAttention`` (present), PyG ``BatchNorm`` = a module holding
``self.module = torch.nn.BatchNorm1d(channels)`` [3P-memory] (hence the ``norm1.module.weight`` key names).
"""
import torch
from torch import nn
import torch.nn.functional as F


class PyGBatchNorm(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.module = nn.BatchNorm1d(channels)

    def forward(self, x):
        return self.module(x)


class GPSConv(nn.Module):
    def __init__(self, channels, conv, heads=1, dropout=0.0):
        super().__init__()
        self.channels, self.conv, self.heads, self.dropout = channels, conv, heads, dropout
        self.attn = nn.MultiheadAttention(channels, heads, batch_first=True)
        self.mlp = nn.Sequential(nn.Linear(channels, 2 * channels), nn.ReLU(), nn.Dropout(dropout),
                                 nn.Linear(2 * channels, channels), nn.Dropout(dropout))
        self.norm1, self.norm2, self.norm3 = PyGBatchNorm(channels), PyGBatchNorm(channels), PyGBatchNorm(channels)

    def forward(self, x, equiv, run_conv):
        """``run_conv(x, equiv) -> (h, equiv)`` executes the wrapped local MPNN."""
        h, equiv = run_conv(x, equiv)                                        # :113-115
        h = F.dropout(h, p=self.dropout, training=self.training) + x
        h1 = self.norm1(h)
        seq = x.unsqueeze(0)                                                 # to_dense_batch(x, None)  (Q1)
        a, _ = self.attn(seq, seq, seq, need_weights=False)
        a = F.dropout(a[0], p=self.dropout, training=self.training) + x
        h2 = self.norm2(a)
        out = h1 + h2
        out = out + self.mlp(out)
        return self.norm3(out), equiv
