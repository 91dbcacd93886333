#!/usr/bin/env python
"""bench.py -- atoms/sec of a full training step on synthetic radius-graph batches (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5                     # engine arm, one B200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1     # CPU arm (oracle on the host cores)

One "step" = neighbour build (radius graph + CSR plans) -> forward -> loss -> backward -> flat gradient
all-reduce -> fused AdamW, on one batch of the QM9-shape PaiNN workload (configs[1] of BASELINE.json:
9-atom molecules, r = 7, k = 5, PaiNN F = 64 L = 2 R = 5, graph energy head, MSE).  Weak scaling: every rank
owns ``--graphs`` graphs per step; ``value`` = atoms processed by all ranks / max-over-ranks device time.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="qm9_painn")
    ap.add_argument("--graphs", type=int, default=16384, help="graphs per GPU per step")
    ap.add_argument("--nbatches", type=int, default=4, help="distinct pre-generated batches cycled through")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--ref-graphs", type=int, default=1024, help="graphs per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16 = the named config: tensor-core (tcgen05, TF32-in/fp32-acc) Linears; fp32 = exact SIMT kernels")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# -----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's training step on the host cores
# -----------------------------------------------------------------------------------------------------------
def cpu_step_rate(workload, graphs, steps, warmup):
    """atoms/s of the pure-torch oracle (same model, same batch shape, fp32, all host threads).  The radius graph
    is built once outside the timed steps, as the reference does at preprocessing."""
    import oracle
    from oracle.workloads import ARCH, add_edges_cpu, make_samples
    kw = ARCH[workload]
    model = oracle.base.create_model(**kw)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    batch = add_edges_cpu(make_samples(workload, graphs, seed=4321), workload)
    mlip = kw.get("enable_interatomic_potential", False)
    hi = [torch.arange(graphs)]

    def step():
        opt.zero_grad()
        if mlip:
            batch.pos.requires_grad_(True)
            loss, _ = model.energy_force_loss(model(batch), batch)
        else:
            loss, _ = model.loss(model(batch), batch.y, hi)
        loss.backward()
        opt.step()
        return float(loss)

    # "all the host threads it can use": ATen's intra-op pool does not scale to 128 threads on these small
    # tensors, so time one step at several pool sizes and keep the fastest (reported as `cores`).
    ncpu = os.cpu_count() or 1
    best = None
    for nt in sorted({ncpu, max(1, ncpu // 2), 32, 16, 8}):
        if nt > ncpu:
            continue
        torch.set_num_threads(nt)
        step()
        t0 = time.perf_counter()
        step()
        d1 = time.perf_counter() - t0
        if best is None or d1 < best[0]:
            best = (d1, nt)
    torch.set_num_threads(best[1])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    atoms = batch.pos.shape[0]
    return atoms / dt, dt, atoms, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, dt, atoms, threads = cpu_step_rate(args.workload, args.ref_graphs, args.steps, max(args.warmup, 1))
    sample = "%d graphs (%d atoms) per step, %d steps; edges prebuilt as the reference does at preprocessing" % (
        args.ref_graphs, atoms, args.steps)
    line = {"impl": "reference", "metric": "atoms_per_sec_training_step", "value": rate, "unit": "atoms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": PaiNN F=64 L=2 R=5, 9-atom graphs r=7 k=5, graph energy head, MSE, AdamW",
                       "graphs_per_step": args.ref_graphs, "note": "pure-torch oracle restating the reference's PyG path (PyG is not installable here)"},
            "cpu_baseline": {"value": rate, "unit": "atoms/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# -----------------------------------------------------------------------------------------------------------
# engine arm
# -----------------------------------------------------------------------------------------------------------
def run_engine(args):
    import hydragnn_b200 as hb
    from hydragnn_b200 import _lib, ops, radius
    from hydragnn_b200.synthetic import ARCH, WORKLOADS, make_samples

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, kw = WORKLOADS[args.workload], ARCH[args.workload]
    mlip = kw.get("enable_interatomic_potential", False)
    model = hb.get_distributed_model(hb.set_precision(hb.create_model(**kw), args.precision))
    opt = hb.FlatAdamW(model, lr=1e-3)
    G = args.graphs

    # host batches (pinned) and their static device twins; one CUDA graph per batch
    host, devb, hidx = [], [], None
    keys = ["x", "pos", "y"] + (["energy", "forces"] if mlip else [])
    for b in range(args.nbatches):
        cpu = make_samples(args.workload, G, seed=1234 + 1000 * rank + b)
        hb_ = {k: cpu[k].pin_memory() for k in keys}
        host.append(hb_)
        d = hb.Batch(**{k: torch.empty_like(v, device=dev) for k, v in hb_.items()})
        d.batch, d.ptr, d._num_graphs = cpu.batch.to(dev), cpu.ptr.to(dev).int(), G     # topology of the batch: resident
        for k in keys:
            d[k].copy_(hb_[k])
        devb.append(d)
    n_atoms = devb[0].pos.shape[0]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    hidx = None if mlip else hb.get_head_indices(model, devb[0])
    gptr = devb[0].ptr

    def step(d, known_e=None):
        """the full hot path on one resident batch"""
        ei, rowptr = radius.radius_graph(d.pos, w["radius"], gptr, G, False, w["max_neighbours"], known_e=known_e)
        d.edge_index = ei
        d._hgb_col_sorted = (ei, rowptr)                      # what hb.get_radius_graph(...)(d) records: edges grouped by target
        d.__dict__.pop("_hgb_plan", None)                     # plans are rebuilt every step (new edges)
        opt.zero_grad()
        m = model.module
        if mlip:
            d.pos.requires_grad_(True)
            loss, _ = m.energy_force_loss(model(d), d)
        else:
            loss, _ = m.loss(model(d), d.y, hidx)
        loss.backward()
        flat = opt.gather_grads()
        return loss.detach(), flat

    # eager warm-up (also measures E per batch and the number of libhgb launches per step)
    n_edges = []
    for d in devb:
        step(d)
        n_edges.append(int(d.edge_index.shape[1]))
    torch.cuda.synchronize()
    c0 = _lib.launch_count()
    loss, flat = step(devb[0])
    if ws > 1:
        dist.all_reduce(flat)
    opt.step(1.0 / ws)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - c0

    use_graph = not args.no_graph
    graphs, losses = [], []
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for d, e in zip(devb, n_edges):
                step(d, known_e=e)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for d, e in zip(devb, n_edges):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                l, _ = step(d, known_e=e)
                if ws == 1:
                    opt.step(1.0)
            graphs.append(g)
            losses.append(l)
        g_opt = None
        if ws > 1:
            g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_opt):
                opt.step(1.0 / ws)

    loss_host = torch.zeros(1).pin_memory()
    # e2e: the inputs of step i+1 travel host -> device on a copy stream while step i computes (what a prefetching loader
    # does); every step still pays its own H2D copy and its own loss read-back inside the timed region.
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(args.nbatches)]      # inputs of batch b are on the device
    released = [torch.cuda.Event() for _ in range(args.nbatches)]   # the step that used batch b has been enqueued and finished
    state = {"prefetched": -1}

    def issue_copy(i):
        b = i % args.nbatches
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(released[b])
            for k in keys:
                devb[b][k].detach().copy_(host[b][k], non_blocking=True)
            ready[b].record(copy_stream)
        state["prefetched"] = i

    for ev in released:
        ev.record()

    def run_step(i, e2e):
        b = i % args.nbatches
        d = devb[b]
        if e2e:
            if state["prefetched"] != i:
                issue_copy(i)
            torch.cuda.current_stream().wait_event(ready[b])
        if use_graph:
            graphs[b].replay()
            if ws > 1:
                dist.all_reduce(opt.flat_g)
                g_opt.replay()
            l = losses[b]
        else:
            l, flat = step(d)
            if ws > 1:
                dist.all_reduce(flat)
            opt.step(1.0 / ws)
        if e2e:
            released[b].record()
            loss_host.copy_(l.reshape(1), non_blocking=True)
            if args.nbatches > 1:
                issue_copy(i + 1)

    def timed(e2e, with_clocks):
        for i in range(args.warmup):
            run_step(i, e2e)
        torch.cuda.synchronize()
        state["prefetched"] = -1            # the first timed step issues (and waits for) its own copy
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if with_clocks and rank == 0 else None
        if sampler:
            sampler.start()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            run_step(i, e2e)
        t1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if ws > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / args.steps, clocks

    ms_dev, clocks = timed(False, True)
    ms_e2e, _ = timed(True, False)
    value = n_atoms * ws / (ms_dev * 1e-3)
    e2e = n_atoms * ws / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (fused PaiNN message, F = hidden_dim layer), timed alone with an L2 flush
    roof = None
    if rank == 0 and kw["mpnn_type"] == "PAINN":
        roof = painn_message_roofline(model.module, devb[0], ops, dev)

    cpu_base = None
    if rank == 0 and not args.skip_cpu_baseline:
        rate, dt, atoms, threads = cpu_step_rate(args.workload, args.ref_graphs, 3, 1)
        cpu_base = {"value": rate, "unit": "atoms/s", "cores": threads, "kind": "port",
                    "sample": "%d graphs (%d atoms) per step, 3 timed steps after 1 warm-up; edges prebuilt" % (args.ref_graphs, atoms)}

    if rank == 0:
        line = {"metric": "atoms_per_sec_training_step", "value": value, "unit": "atoms/s", "n_gpus": ws, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload + ": PaiNN F=64 L=2 R=5, 9-atom graphs r=7 k=5, graph energy head, MSE, AdamW"
                           if args.workload == "qm9_painn" else args.workload,
                           "precision": "bf16 config -> fp32 parameters/activations, large-M Linears on tcgen05 kind::tf32 with fp32 "
                                        "accumulation (>= bf16 autocast of the reference)" if args.precision == "bf16" else "fp32",
                           "graphs_per_gpu": G, "atoms_per_gpu": n_atoms, "edges_per_gpu": n_edges[0],
                           "parallelism": "dp%d (graphs sharded by rank, one flat gradient all-reduce)" % ws,
                           "step": "radius graph + CSR plans + fwd + loss + bwd + all-reduce + fused AdamW",
                           "launch": "cuda-graph replay per batch" if use_graph else "eager",
                           "e2e_pipeline": "every step copies its x/pos/y from pinned host memory (copy stream, issued one step ahead) "
                                           "and reads its loss back",
                           "l2": "%d distinct batches cycled; per-step working set (activations + saved tensors) >> 126 MB L2" % args.nbatches},
                "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
                "e2e": {"value": e2e, "unit": "atoms/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
                "roofline": roof, "cpu_baseline": cpu_base}
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


def painn_message_roofline(m, d, ops, dev):
    """CUDA-event timing of hgb_painn_message_fwd (the segmented-scatter kernel of SURVEY.md 8d) alone on the real layer-2
    shapes (F = hidden_dim), L2 flushed between launches.

    `achieved` uses the COMPULSORY bytes of the shared-memory-tiled kernel ("roofline bytes"): every tensor
    crosses HBM once -- N*(3F phi + 3F v + F s read, F s_out + 3F v_out written)*4 + E*64 (edge records) + (N+1)*4.
    `survey_formula` is SURVEY.md 8(d)'s scatter figure, which charges every edge its own gathered rows
    (E*(6F*4 + 8 + 48) + N*(8F*4 + 4)); with the gathers served from shared memory it exceeds the HBM peak."""
    hbm, src = peaks()
    from hydragnn_b200.stacks import Base
    plan = Base.plan_for(d)
    n, e = plan.num_nodes, plan.num_edges
    f = m.hidden_dim
    r = m.num_radial
    conv = m.graph_convs[-1]
    with torch.no_grad():
        _, ln, unit = ops.EdgeGeomFn.apply(d.pos.detach(), None, plan, 1e-9)
        epack = ops.PainnEdgeEmbedFn.apply(unit, ln, r, m.radius)
        rec = ops.painn_edge_records(epack, plan, "row")
        s = torch.randn(n, f, device=dev)
        v = torch.randn(n, 3, f, device=dev)
        phi = torch.randn(n, 3 * f, device=dev)
        msg = conv.module_0
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(13):
            flush.zero_()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            ops.PainnMessageFn.apply(phi, s, v, epack, msg.filter_layer.weight, msg.filter_layer.bias, None, plan, rec)
            t1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(t0.elapsed_time(t1))
    ms = sum(ts) / len(ts)
    alg = n * 11 * f * 4 + e * 64 + (n + 1) * 4
    survey = e * (6 * f * 4 + 8 + 48) + n * (8 * f * 4 + 4)
    ach = alg / (ms * 1e-3) / 1e9
    return {"kernel": "painn_message_fwd_tiled_kernel<false,5,64>", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
            "frac": ach / hbm, "traffic": NCU_TRAFFIC_BYTES if (n, e, f) == NCU_TRAFFIC_SHAPE else None,
            "traffic_source": "profiles/r01_ncu_painn_message_final_details.csv (dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
            "peak_source": src + " (burst copy figure; kernel timed alone, L2 flushed)",
            "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms,
            "survey_formula": {"bytes_per_launch": survey, "achieved": survey / (ms * 1e-3) / 1e9, "frac": survey / (ms * 1e-3) / 1e9 / hbm}}


# measured once with `ncu --set full` on the bench shapes (N, E, F): 315.2 MB read + 120.7 MB written
NCU_TRAFFIC_SHAPE = (147456, 786432, 64)
NCU_TRAFFIC_BYTES = 435925760


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
