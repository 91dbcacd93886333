"""Oracle: edge geometry.  Test infrastructure only (see oracle/__init__.py)."""
import math

import torch


def edge_vectors_and_lengths(pos, edge_index, shifts=None, normalize=False, eps=1e-9):
    """Restates ``get_edge_vectors_and_lengths``
    (hydragnn/utils/model/operations.py:21-36).

    ``vec = pos[edge_index[1]] - pos[edge_index[0]] + shifts``; ``len = ||vec||``
    kept as ``[E, 1]``; with ``normalize`` the vector is divided by ``len + eps``
    (note: *plus* eps, which is how EGNN's ``eps=1.0`` quirk Q3 arises).
    """
    snd, rcv = edge_index[0], edge_index[1]
    vec = pos[rcv] - pos[snd]
    if shifts is not None:
        vec = vec + shifts
    length = torch.linalg.norm(vec, dim=-1, keepdim=True)
    if normalize:
        return vec / (length + eps), length
    return vec, length


def sinc_expansion(dist, num_radial, cutoff):
    """``sin(n*pi*d/rc)/d`` for n = 1..R  (hydragnn/models/PAINNStack.py:331-338)."""
    n = torch.arange(1, num_radial + 1, device=dist.device)
    return torch.sin(dist * n * math.pi / cutoff) / dist


def cosine_cutoff(dist, cutoff):
    """Behler-Parrinello cutoff, zero for d >= rc
    (hydragnn/models/PAINNStack.py:341-352)."""
    val = 0.5 * (torch.cos(math.pi * dist / cutoff) + 1.0)
    return torch.where(dist < cutoff, val, torch.zeros_like(val))


def segment_sum(src, index, num_segments):
    """``unsorted_segment_sum`` (hydragnn/models/EGCLStack.py:294-300)."""
    out = src.new_zeros((num_segments,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def segment_mean(src, index, num_segments):
    """``unsorted_segment_mean`` (hydragnn/utils/model/model.py:441-448):
    sum divided by the per-segment element count clamped to >= 1."""
    total = segment_sum(src, index, num_segments)
    cnt = segment_sum(torch.ones_like(src), index, num_segments)
    return total / cnt.clamp(min=1)


def graph_pool(x, batch, num_graphs, mode):
    """``global_{mean,add,max}_pool`` as used by ``Base`` (hydragnn/models/Base.py:147-170).

    [3P-memory] torch_geometric 2.6.1: ``scatter(x, batch, dim=0, dim_size=G,
    reduce=...)``; mean divides by max(count, 1); an empty graph gives 0.
    """
    if mode in ("add", "sum"):
        return segment_sum(x, batch, num_graphs)
    if mode == "mean":
        cnt = torch.bincount(batch, minlength=num_graphs).clamp(min=1).to(x.dtype)
        return segment_sum(x, batch, num_graphs) / cnt[:, None]
    if mode == "max":
        out = x.new_full((num_graphs, x.shape[1]), float("-inf"))
        out = out.scatter_reduce(0, batch[:, None].expand_as(x), x, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    raise ValueError("Unsupported graph_pooling: " + str(mode))
