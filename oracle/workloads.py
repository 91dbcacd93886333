"""Oracle-side helpers for the synthetic workloads (test infrastructure only)."""
import torch

from hydragnn_b200.synthetic import ARCH, WORKLOADS, make_samples  # noqa: F401

from .radius_graph import radius_graph, radius_graph_pbc


def add_edges_cpu(batch, name):
    """Build the edges of a synthetic batch with the oracle's radius graphs (per sample for PBC, as
    the reference does at preprocessing, hydragnn/preprocess/serialized_dataset_loader.py:134-150)."""
    w = WORKLOADS[name]
    if w.get("pbc"):
        eis, shs = [], []
        n = w["n"]
        for g in range(batch.num_graphs):
            ei, sh = radius_graph_pbc(batch.pos[g * n:(g + 1) * n], batch.cell[g], batch.pbc[g], w["radius"], False,
                                      w["max_neighbours"])
            eis.append(ei + g * n)
            shs.append(sh)
        batch.edge_index, batch.edge_shifts = torch.cat(eis, 1), torch.cat(shs)
    else:
        batch.edge_index = radius_graph(batch.pos, w["radius"], batch.batch, False, w["max_neighbours"])
        batch.edge_shifts = torch.zeros(batch.edge_index.shape[1], 3)
    return batch
