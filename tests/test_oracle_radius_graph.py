"""CPU: the reference's own known-answer / property tests, run against the oracle's
radius graphs (tests/test_periodic_boundary_conditions.py:82-127,
tests/test_rotational_invariance.py:70-116 in the reference)."""
import numpy as np
import torch

from oracle.radius_graph import radius_graph, radius_graph_pbc


def _check_pbc(pos, cell, r, expected, expected_loops, k=100000):
    n = pos.shape[0]
    for loop, exp in ((False, expected), (True, expected_loops)):
        ei, sh = radius_graph_pbc(pos, cell, [True, True, True], r, loop=loop, max_num_neighbors=k)
        assert ei.shape[1] == exp * n
        vec = pos[ei[1]] - pos[ei[0]] + sh
        d = vec.norm(dim=-1)
        assert ((d <= r) & (d >= 0)).all()
        if not loop:
            assert (np.bincount(ei[1].numpy(), minlength=n) == exp).all()


def test_periodic_h2():
    # ci_periodic.json: radius 0.9 ... but H2 bond here is 0.745 A, box 3 A -> 1 neighbour / atom
    cell = torch.eye(3) * 3.0
    pos = torch.tensor([[1.0, 1.0, 1.0], [1.43, 1.43, 1.43]])
    _check_pbc(pos, cell, 0.9, 1, 2)


def bcc_supercell(a=3.6, reps=5, dtype=torch.float64):
    base = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]], dtype=dtype) * a
    cells = torch.stack(torch.meshgrid(*[torch.arange(reps, dtype=dtype)] * 3, indexing="ij"), -1).reshape(-1, 3) * a
    pos = (cells[:, None, :] + base[None]).reshape(-1, 3)
    return pos, torch.eye(3, dtype=dtype) * a * reps


def test_periodic_bcc_large():
    pos, cell = bcc_supercell()
    assert pos.shape[0] == 250
    _check_pbc(pos, cell, 5.0, 14, 15)


def test_pbc_nearest_k_truncation_and_mixed_pbc():
    pos, cell = bcc_supercell(reps=3)
    ei, sh = radius_graph_pbc(pos, cell, [True, True, True], 5.0, max_num_neighbors=8)
    assert (np.bincount(ei[1].numpy(), minlength=pos.shape[0]) == 8).all()   # the 8 first-shell atoms
    d = (pos[ei[1]] - pos[ei[0]] + sh).norm(dim=-1)
    assert torch.allclose(d, torch.full_like(d, 3.6 * 3 ** 0.5 / 2))
    # slab: no images along z -> surface atoms lose neighbours, all shifts have S_z = 0
    ei2, sh2 = radius_graph_pbc(pos, cell, [True, True, False], 5.0, max_num_neighbors=100)
    assert (sh2[:, 2] == 0).all() and ei2.shape[1] < 14 * pos.shape[0]


def _bct():
    p = []
    for x in range(4):
        for y in range(2):
            for z in range(2):
                p.append([x * 5.218, y * 5.218, z * 7.058])
                p.append([(x + .5) * 5.218, (y + .5) * 5.218, (z + .5) * 7.058])
    return torch.tensor(p)


def _normalize_rotation(pos):
    """[3P-memory B.6] PyG NormalizeRotation(max_points=-1, sort=False)."""
    c = pos - pos.mean(0, keepdim=True)
    _, _, v = torch.linalg.svd(c, full_matrices=False)
    return pos @ v.t()


def _edge_set(pos, r=7.0, k=100000):
    ei = radius_graph(pos, r, None, False, k)
    d = (pos[ei[1]] - pos[ei[0]]).norm(dim=-1)
    return {(int(a), int(b)): float(x) for a, b, x in zip(ei[0], ei[1], d)}


def test_rotational_invariance():
    g = torch.Generator().manual_seed(0)
    for dtype, tol in ((torch.float32, 1e-4), (torch.float64, 1e-12)):
        samples = [_bct().to(dtype)] + [3 * torch.randn(10, 3, generator=g, dtype=dtype) for _ in range(10)]
        for pos in samples:
            a, b = _edge_set(pos), _edge_set(_normalize_rotation(pos))
            assert a.keys() == b.keys()
            assert max(abs(a[k] - b[k]) for k in a) < tol


def test_radius_graph_semantics():
    # ordering: grouped by query ascending, neighbours ascending; cap = k (+1 with the self match, then dropped)
    pos = torch.arange(6, dtype=torch.float32)[:, None] * torch.tensor([[1.0, 0, 0]])
    ei = radius_graph(pos, 2.5, None, False, 2)
    # query 3 sees 1,2,(3),4,5 within 2.5 -> first k+1 = 3 matches {1,2,3}, minus self -> {1,2}
    assert ei[0][ei[1] == 3].tolist() == [1, 2]
    # query 0 sees (0),1,2 -> {1,2}
    assert ei[0][ei[1] == 0].tolist() == [1, 2]
    # batch separation
    batch = torch.tensor([0, 0, 0, 1, 1, 1])
    ei = radius_graph(pos, 10.0, batch, False, 32)
    assert ((ei[0] < 3) == (ei[1] < 3)).all() and ei.shape[1] == 12
    # strict '<'
    ei = radius_graph(pos, 1.0, None, False, 32)
    assert ei.shape[1] == 0
    assert radius_graph(pos[:0], 1.0).shape == (2, 0)
