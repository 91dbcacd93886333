"""Oracle-side helpers for the synthetic workloads (test infrastructure only)."""
import torch

from hydragnn_b200.synthetic import ARCH, WORKLOADS, make_samples  # noqa: F401

from .radius_graph import radius_graph, radius_graph_pbc


def add_edges_cpu(batch, name):
    """Build the edges of a synthetic batch with the oracle's radius graphs (per sample for PBC, as
    the reference does at preprocessing, hydragnn/preprocess/serialized_dataset_loader.py:134-150)."""
    w = WORKLOADS[name]
    if w.get("pbc") or w.get("pbc_box"):
        eis, shs = [], []
        ptr = batch.ptr.tolist()
        for g in range(batch.num_graphs):
            ei, sh = radius_graph_pbc(batch.pos[ptr[g]:ptr[g + 1]], batch.cell[g], batch.pbc[g], w["radius"], False,
                                      w["max_neighbours"])
            eis.append(ei + ptr[g])
            shs.append(sh)
        batch.edge_index, batch.edge_shifts = torch.cat(eis, 1), torch.cat(shs)
    else:
        batch.edge_index = radius_graph(batch.pos, w["radius"], batch.batch, False, w["max_neighbours"])
        batch.edge_shifts = torch.zeros(batch.edge_index.shape[1], 3)
    if w.get("pe_dim"):                                    # serialized_dataset_loader.py:186-189
        batch.rel_pe = (batch.pe[batch.edge_index[0]] - batch.pe[batch.edge_index[1]]).abs()
    return batch


def arch_for(name, batch):
    """ARCH[name] with the data-dependent knobs filled in from an edge-carrying batch: MACE ``avg_num_neighbors`` (measured) and
    the PNA in-degree histogram ``pna_deg``."""
    kw = dict(ARCH[name])
    n, e = batch.pos.shape[0], batch.edge_index.shape[1]
    if kw["mpnn_type"] == "MACE":
        kw["avg_num_neighbors"] = e / n
    if kw["mpnn_type"] == "PNAEq":
        kw["pna_deg"] = torch.bincount(torch.bincount(batch.edge_index[1].cpu(), minlength=n)).tolist()
    return kw
