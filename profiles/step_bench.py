"""Training-step timing of any synthetic workload (one GPU, eager).  usage: python profiles/step_bench.py <workload> [graphs] [steps] [precision]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hydragnn_b200 as hb
from hydragnn_b200.synthetic import ARCH, WORKLOADS, make_samples

name = sys.argv[1]
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
dev = torch.device("cuda")
w = WORKLOADS[name]
b = make_samples(name, G).to(dev); b._num_graphs = G
pbc = w.get("pbc") or w.get("pbc_box")
b = (hb.get_radius_graph_pbc if pbc else hb.get_radius_graph)(w["radius"], w["max_neighbours"])(b)
n, e = b.pos.shape[0], b.edge_index.shape[1]
kw = dict(ARCH[name])
if kw["mpnn_type"] == "MACE":
    kw["avg_num_neighbors"] = e / n
if kw["mpnn_type"] == "PNAEq":
    deg = torch.bincount(b.edge_index[1], minlength=n)
    kw["pna_deg"] = torch.bincount(deg).tolist()
if kw.get("global_attn_engine"):
    b.rel_pe = (b.pe[b.edge_index[0]] - b.pe[b.edge_index[1]]).abs()      # serialized_dataset_loader.py:186-189
mlip = bool(kw.get("enable_interatomic_potential"))
model = hb.get_distributed_model(hb.set_precision(hb.create_model(**kw), prec))
opt = hb.FlatAdamW(model, lr=1e-3)
hi = None if mlip else hb.get_head_indices(model, b)
run = lambda: hb.train_step(model, opt, b, compute_grad_energy=mlip, head_index=hi)
for _ in range(3):
    loss, _ = run()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(steps):
    loss, _ = run()
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / steps
print(json.dumps({"workload": name, "graphs": G, "atoms": n, "edges": e, "precision": prec, "mlip": mlip, "ms_per_step": ms,
                  "atoms_per_s": n / ms * 1e3, "loss": float(loss), "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
