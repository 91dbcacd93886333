"""Training-step timing of the MACE path on the SURVEY C4 shape (one GPU).  usage: python profiles/mace_bench.py [graphs] [steps]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hydragnn_b200 as hb
from hydragnn_b200.synthetic import ARCH, make_samples

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda")
b = make_samples("oc20_mace", G).to(dev); b._num_graphs = G
b = hb.get_radius_graph_pbc(6.0, 128)(b)
n, e = b.pos.shape[0], b.edge_index.shape[1]
kw = dict(ARCH["oc20_mace"], avg_num_neighbors=e / n)
model = hb.set_precision(hb.create_model(**kw), prec)
model = hb.get_distributed_model(model)
opt = hb.FlatAdamW(model, lr=1e-3)
hi = hb.get_head_indices(model, b)
for _ in range(3):
    loss, _ = hb.train_step(model, opt, b, head_index=hi)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(steps):
    loss, _ = hb.train_step(model, opt, b, head_index=hi)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / steps
print(json.dumps({"workload": "oc20_mace", "graphs": G, "atoms": n, "edges": e, "precision": prec, "ms_per_step": ms,
                  "atoms_per_s": n / ms * 1e3, "loss": float(loss), "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
