"""Graph sample / mini-batch containers.

Host-side mirror of the slice of ``torch_geometric.data.Data`` / ``Batch`` that the
reference's hot path touches (hydragnn/train/train_validate_test.py:74-84 iterates
``data.items()`` and calls ``data.to(device)``; hydragnn/models/Base.py:697-846 reads
``x, pos, edge_index, edge_shifts, edge_attr, batch, dataset_name``;
hydragnn/models/create.py:626-738 reads ``energy, forces``).  torch_geometric is not
installed in this image, so the engine ships its own container with the same surface.
"""
import copy

import torch

# attributes concatenated along dim 0 with one row per node / per edge / per graph
_EDGE_KEYS = ("edge_attr", "edge_shifts", "rel_pe")


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    # -- mapping-style access used by the reference train loop ---------------------
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def items(self):
        return [(k, self.__dict__[k]) for k in self.keys()]

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.__dict__ and self.__dict__[key] is not None

    def __getattr__(self, name):      # only reached for attributes that were never set
        if name in ("edge_attr", "edge_shifts", "batch", "pe", "rel_pe", "y", "y_loc", "graph_attr",
                    "cell", "pbc", "energy", "forces", "edge_index", "x", "pos"):
            return None
        raise AttributeError(name)

    @property
    def num_nodes(self):
        for k in ("x", "pos"):
            v = self.__dict__.get(k)
            if torch.is_tensor(v):
                return v.shape[0]
        return 0

    @property
    def num_edges(self):
        ei = self.__dict__.get("edge_index")
        return 0 if ei is None else ei.shape[1]

    @property
    def num_graphs(self):
        n = self.__dict__.get("_num_graphs")
        if n is not None:
            return n
        b = self.__dict__.get("batch")
        return 1 if b is None else int(b.max()) + 1

    def to(self, device=None, dtype=None, non_blocking=False):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                if dtype is not None and torch.is_floating_point(v):
                    v = v.to(dtype=dtype)
                if device is not None:
                    v = v.to(device, non_blocking=non_blocking)
                self.__dict__[k] = v
        return self

    def clone(self):
        out = self.__class__()
        for k, v in self.__dict__.items():
            if k.startswith("_hgb"):          # cached engine plans are tied to this object's tensors
                continue
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def __repr__(self):
        parts = []
        for k, v in self.items():
            parts.append("%s=%s" % (k, list(v.shape) if torch.is_tensor(v) else v))
        return "%s(%s)" % (self.__class__.__name__, ", ".join(parts))


class Batch(Data):
    """Mini-batch of graphs stored as one disjoint graph (``batch`` is sorted)."""

    @classmethod
    def from_data_list(cls, samples):
        out = cls()
        keys = []
        for s in samples:
            for k in s.keys():
                if k not in keys:
                    keys.append(k)
        offs, n = [], 0
        for s in samples:
            offs.append(n)
            n += s.num_nodes
        for k in keys:
            vals = [getattr(s, k) for s in samples]
            if not all(torch.is_tensor(v) for v in vals):
                out.__dict__[k] = vals
                continue
            if k == "edge_index":
                out.edge_index = torch.cat([v + o for v, o in zip(vals, offs)], dim=1)
            elif k in ("cell",):
                out.__dict__[k] = torch.stack([v.reshape(3, 3) for v in vals])
            elif k in ("pbc",):
                out.__dict__[k] = torch.stack([v.reshape(3) for v in vals])
            elif vals[0].dim() == 0:
                out.__dict__[k] = torch.stack(vals)
            else:
                out.__dict__[k] = torch.cat(vals, dim=0)
        out.batch = torch.cat([torch.full((s.num_nodes,), i, dtype=torch.long) for i, s in enumerate(samples)])
        out.ptr = torch.tensor(offs + [n], dtype=torch.long)
        out._num_graphs = len(samples)
        return out


def collate_to_device(samples, device, non_blocking=True):
    """``Batch.from_data_list(samples).to(device)`` with the index work done on the GPU: every tensor field is
    packed into ONE pinned staging buffer and copied once; ``ptr`` is a device scan of the node counts, ``batch`` and the
    node offsets of ``edge_index`` come from two small kernels instead of per-sample Python arithmetic."""
    from . import _lib, ops
    device = torch.device(device)
    out = Batch()
    g = len(samples)
    keys = []
    for s in samples:
        for k in s.keys():
            if k not in keys:
                keys.append(k)

    def staged(parts, dim=0, stack=False):
        t = torch.stack(parts) if stack else torch.cat(parts, dim=dim)
        return t.pin_memory().to(device, non_blocking=non_blocking)

    counts = torch.tensor([s.num_nodes for s in samples], dtype=torch.int32)
    nptr = ops.exclusive_scan(counts.pin_memory().to(device, non_blocking=non_blocking))
    n = int(counts.sum())
    for k in keys:
        vals = [getattr(s, k) for s in samples]
        if not all(torch.is_tensor(v) for v in vals):
            out.__dict__[k] = vals
        elif k == "edge_index":
            ecnt = torch.tensor([v.shape[1] for v in vals], dtype=torch.int32)
            eptr = ops.exclusive_scan(ecnt.pin_memory().to(device, non_blocking=non_blocking))
            local = staged(vals, dim=1).contiguous()
            e = local.shape[1]
            ei = torch.empty_like(local)
            _lib.call("hgb_collate_offset_edges", ops._p(local), ops._p(eptr), ops._p(nptr), g, e, ops._p(ei), ops._stream())
            out.edge_index = ei
        elif k == "cell":
            out.cell = staged([v.reshape(3, 3) for v in vals], stack=True)
        elif k == "pbc":
            out.pbc = staged([v.reshape(3) for v in vals], stack=True)
        elif vals[0].dim() == 0:
            out.__dict__[k] = staged(vals, stack=True)
        else:
            out.__dict__[k] = staged(vals)
    batch = torch.empty(n, dtype=torch.int64, device=device)
    _lib.call("hgb_collate_batch_vector", ops._p(nptr), g, n, ops._p(batch), ops._stream())
    out.batch, out.ptr, out._num_graphs = batch, nptr.to(torch.int64), g
    return out
