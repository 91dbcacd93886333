"""Oracle: radius graphs (open and periodic).  Test infrastructure only.

Follows ``get_radius_graph*`` / ``RadiusGraphPBC``
(hydragnn/preprocess/graph_samples_checks_and_updates.py:112-141, 144-417).
The neighbour searches themselves live in third-party wheels that are absent
from the reference tree, so their published algorithms are restated:

* torch_cluster == 1.6.3 (requirements-pyg.txt:4) ``radius_graph`` -- reached via
  PyG ``RadiusGraph`` (``graph_samples...py:113-117``).  [3P-memory] semantics
  restated here (CUDA kernel order): ``radius(x, x, r, batch, batch,
  k if loop else k + 1)`` visits candidates of the same graph in ascending index,
  accepts ``d2 < r*r`` (strict, arithmetic in the dtype of ``pos``), stops after
  the cap, then ``radius_graph`` drops ``row == col`` when ``loop=False``.
  Result: ``edge_index = [neighbour; query]`` grouped by query ascending,
  neighbours ascending.  d2 is accumulated as ``(dx*dx + dy*dy) + dz*dz`` with
  every operation rounded (no FMA contraction) -- the CUDA kernel uses
  ``__fmul_rn/__fadd_rn`` so the accept test is bit-identical.
* vesin == 0.4.2 / ase == 3.26.0 (requirements-base.txt:16,6)
  ``ase_neighbor_list("ijS")``: full list of (i, j, S) with
  ``|pos[j] - pos[i] + S @ cell| < cutoff`` in float64, never the zero-shift self
  pair.

Ordering under ``max_neighbours`` truncation is "parity unpinned" (the
reference has no golden vectors); counts and edge sets are pinned by the
reference's known-answer tests (tests/test_periodic_boundary_conditions.py:82-127,
tests/test_rotational_invariance.py:70-116) which tests/ re-run on this oracle.
"""
import numpy as np
import torch


def _graph_ptr(batch, n):
    if batch is None:
        return np.array([0, n], dtype=np.int64)
    b = batch.cpu().numpy()
    assert (np.diff(b) >= 0).all(), "batch must be sorted"
    g = int(b.max()) + 1 if n > 0 else 0
    return np.concatenate([[0], np.cumsum(np.bincount(b, minlength=g))]).astype(np.int64)


def radius_graph(pos, r, batch=None, loop=False, max_num_neighbors=32):
    """Open-boundary radius graph; returns ``edge_index [2, E]`` int64."""
    p = pos.detach().cpu().numpy()
    dt = p.dtype
    n = p.shape[0]
    ptr = _graph_ptr(batch, n)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    r2 = dt.type(r) * dt.type(r)
    rows, cols = [], []
    for g in range(len(ptr) - 1):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        if hi <= lo:
            continue
        q = p[lo:hi]
        d = q[None, :, :] - q[:, None, :]            # d[i, j] = x[j] - y[i]
        sq = d * d
        d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]  # rounded after every op
        ok = d2 < r2
        rank = np.cumsum(ok, axis=1) - 1             # ascending-index visitation
        ok &= rank < cap
        if not loop:
            ok &= ~np.eye(hi - lo, dtype=bool)
        qi, nj = np.nonzero(ok)                      # row-major: query asc, nbr asc
        rows.append(nj + lo)
        cols.append(qi + lo)
    if rows:
        ei = np.stack([np.concatenate(rows), np.concatenate(cols)]).astype(np.int64)
    else:
        ei = np.zeros((2, 0), dtype=np.int64)
    return torch.from_numpy(ei)


# ----------------------------------------------------------------------------------
# periodic
# ----------------------------------------------------------------------------------

def _neighbor_list_ijS(pos64, cell64, pbc, cutoff):
    """Restated vesin/ASE ``neighbor_list("ijS")`` (brute force over images)."""
    n = pos64.shape[0]
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros((0, 3), dtype=np.int64), np.zeros(0)
    pbc = [bool(b) for b in pbc]
    nimg = [0, 0, 0]
    if any(pbc):
        vol = abs(np.linalg.det(cell64))
        inv = np.linalg.inv(cell64)
        frac = pos64 @ inv
        for k in range(3):
            if not pbc[k]:
                continue
            a, b = cell64[(k + 1) % 3], cell64[(k + 2) % 3]
            height = vol / np.linalg.norm(np.cross(a, b))
            spread = frac[:, k].max() - frac[:, k].min()
            nimg[k] = int(np.ceil(cutoff / height + spread)) + 1
    rng = [np.arange(-m, m + 1) for m in nimg]
    S = np.stack(np.meshgrid(*rng, indexing="ij"), -1).reshape(-1, 3).astype(np.int64)
    out_i, out_j, out_S, out_len = [], [], [], []
    c2 = cutoff * cutoff
    base = pos64[None, :, :] - pos64[:, None, :]            # [i, j] = pos[j] - pos[i]
    for s in S:
        # every operation individually rounded, in this order (the CUDA kernel does the same)
        sh = (float(s[0]) * cell64[0] + float(s[1]) * cell64[1]) + float(s[2]) * cell64[2]
        v = base + sh[None, None, :]
        d2 = (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]
        ok = d2 < c2
        if not s.any():
            ok &= ~np.eye(n, dtype=bool)                    # never the zero-shift self pair
        ii, jj = np.nonzero(ok)
        out_i.append(ii)
        out_j.append(jj)
        out_S.append(np.broadcast_to(s, (ii.size, 3)))
        out_len.append(np.sqrt(d2[ii, jj]))
    return (np.concatenate(out_i).astype(np.int64), np.concatenate(out_j).astype(np.int64),
            np.concatenate(out_S).astype(np.int64), np.concatenate(out_len))


def _shift_vectors(S, cell64):
    Sf = S.astype(np.float64)
    return (Sf[:, 0:1] * cell64[0][None] + Sf[:, 1:2] * cell64[1][None]) + Sf[:, 2:3] * cell64[2][None]


def limit_neighbors(src, dst, length, shifts, k):
    """``RadiusGraphPBC._limit_neighbors`` (graph_samples...py:266-298): stable
    lexsort by (dst, length), keep the first ``k`` per dst.  Ties in length are
    broken canonically by (src, Sx, Sy, Sz) (vesin's raw order is unspecified)."""
    order = np.lexsort((shifts[:, 2], shifts[:, 1], shifts[:, 0], src, length, dst))
    src, dst, length, shifts = src[order], dst[order], length[order], shifts[order]
    m = dst.size
    if m == 0:
        return src, dst, length, shifts
    start = np.ones(m, dtype=bool)
    start[1:] = dst[1:] != dst[:-1]
    first = np.flatnonzero(start)
    rank = np.arange(m) - first[np.cumsum(start) - 1]
    keep = rank < k
    return src[keep], dst[keep], length[keep], shifts[keep]


def radius_graph_pbc(pos, cell, pbc, r, loop=False, max_num_neighbors=32):
    """``RadiusGraphPBC.__call__`` (graph_samples...py:149-256) for ONE sample.

    Returns ``(edge_index [2,E] int64, edge_shifts [E,3] in pos.dtype)``.
    Mixed PBC: the reference inflates non-periodic cell vectors so no image is
    found across them (:356-414); enumerating no images along those axes gives
    the same list.  Deviation: where the reference adds an edge from a *random*
    node to an isolated one after 3 failed radius growths (:312-320,
    ``np.random.choice``), this restatement uses node ``(m + 1) % n`` so results
    are reproducible.
    """
    p64 = pos.detach().cpu().numpy().astype(np.float64)
    c64 = np.asarray(cell.detach().cpu().numpy() if torch.is_tensor(cell) else cell, dtype=np.float64)
    pbc_l = [bool(b) for b in (pbc.tolist() if torch.is_tensor(pbc) else pbc)]
    n = p64.shape[0]
    cutoff = float(r)
    for attempt in range(3):
        # the reference recomputes |pos[dst]-pos[src]+S@cell| with numpy (:179-183); same values
        src, dst, S, length = _neighbor_list_ijS(p64, c64, pbc_l, cutoff)
        if not loop:  # _remove_true_self_loops (:258-264) -- a no-op on a vesin list
            keep = ~((src == dst) & (S == 0).all(1))
            src, dst, length, S = src[keep], dst[keep], length[keep], S[keep]
        src, dst, length, S = limit_neighbors(src, dst, length, S, max_num_neighbors)
        if np.unique(dst).size == n:
            break
        if attempt < 2:
            cutoff *= 1.25
        else:  # _ensure_connected (:300-322)
            missing = np.setdiff1d(np.arange(n), np.unique(dst))
            for m in missing:
                src = np.append(src, (m + 1) % n if n > 1 else 0)
                dst = np.append(dst, m)
                length = np.append(length, cutoff - 1e-8)
                S = np.vstack([S, np.zeros((1, 3), dtype=S.dtype)])
    if loop:  # appended after truncation (:221-232)
        ar = np.arange(n, dtype=np.int64)
        src, dst = np.concatenate([src, ar]), np.concatenate([dst, ar])
        S = np.vstack([S, np.zeros((n, 3), dtype=S.dtype)])
    ei = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
    np_dt = np.float32 if pos.dtype == torch.float32 else np.float64
    shifts = torch.from_numpy(_shift_vectors(S, c64).astype(np_dt))
    return ei, shifts


def canonical_sort(edge_index, shifts=None):
    """Canonical order used by the bit-exact parity checks: dst asc, src asc
    (then shift lexicographic)."""
    ei = edge_index.cpu().numpy()
    keys = [ei[0], ei[1]]
    if shifts is not None:
        s = shifts.cpu().numpy()
        keys = [s[:, 2], s[:, 1], s[:, 0]] + keys
    order = np.lexsort(tuple(keys))
    out = torch.from_numpy(ei[:, order])
    return (out, shifts[torch.from_numpy(order)]) if shifts is not None else out
