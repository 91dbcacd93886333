"""Data-parallel training step of the engine.

Mirror of the hot loop ``train()`` in hydragnn/train/train_validate_test.py:629-801 and of the DDP wrap in
hydragnn/utils/distributed/distributed.py:396-481, redesigned for one NVSwitch box:

* parameters and gradients live in ONE flat fp32 buffer each (the modules hold views), so the optimizer is a
  single fused AdamW kernel and the only collective of the step is ONE ``all_reduce`` over the flat gradient
  (graphs shard by rank with no other communication -- SURVEY 8e); the 1/world_size scaling is folded into
  the AdamW kernel;
* no mpi4py anywhere on the step path (the reference's per-epoch ``MPI.allreduce(nbatch, MIN)``,
  :672, becomes a ``dist.all_reduce(MIN)``);
* ``GraphedTrainStep`` captures forward + loss + backward + flatten + optimizer of a fixed-shape batch in a CUDA
  graph (all kernels are launched through ctypes on the capturing stream; nothing synchronises).
"""
import os

import torch
import torch.distributed as dist

from . import ops

PRECISION_MAP = {"bf16": torch.float32, "fp32": torch.float32}     # parameters stay fp32 (reference :43-49)


def resolve_precision(precision):
    prec = {"bfloat16": "bf16", "float32": "fp32", "float": "fp32", None: "fp32"}.get(precision, precision)
    prec = str(prec).lower()
    if prec in ("fp64", "float64", "double"):
        raise ValueError("Unsupported precision fp64: the b200 engine computes in fp32 / bf16")
    if prec not in PRECISION_MAP:
        raise ValueError("Unsupported precision %s. Choose from %s." % (precision, list(PRECISION_MAP)))
    return prec, PRECISION_MAP[prec], (torch.bfloat16 if prec == "bf16" else None)


def move_batch_to_device(data, param_dtype=torch.float32, device=None):
    """hydragnn/train/train_validate_test.py:74-84: cast floats, move everything."""
    for key, value in data.items():
        if torch.is_tensor(value) and torch.is_floating_point(value):
            data[key] = value.to(dtype=param_dtype)
    return data.to(device, non_blocking=True)


def get_head_indices(model, data):
    """Single-head / all-graph-head cases of hydragnn/train/train_validate_test.py:494-557."""
    m = getattr(model, "module", model)
    if m.num_heads == 1:
        return [torch.arange(data.y.shape[0], device=data.y.device)]
    if all(t == "graph" for t in m.head_type):
        dims, tot = m.head_dims, sum(m.head_dims)
        g = data.num_graphs
        base = torch.arange(g, device=data.y.device) * tot
        out, off = [], 0
        for d in dims:
            out.append((base[:, None] + off + torch.arange(d, device=data.y.device)[None, :]).reshape(-1))
            off += d
        return out
    y_loc = data.y_loc.to(data.y.device)
    start = (torch.cumsum(y_loc[:, -1], 0) - y_loc[:, -1]).view(-1, 1)
    out = []
    for ih in range(m.num_heads):
        lo, hi = (start + y_loc[:, ih:ih + 1]).flatten().tolist(), (start + y_loc[:, ih + 1:ih + 2]).flatten().tolist()
        out.append(torch.cat([torch.arange(a, b, device=data.y.device) for a, b in zip(lo, hi)]))
    return out


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics over one flat buffer (``hydragnn/utils/optimizer/optimizer.py:12-40`` default).  After
    construction every parameter of ``model`` is a view into ``self.flat_p`` (so checkpoints / ``state_dict`` are unchanged).

    It IS a ``torch.optim.Optimizer`` (one param group), so ``ReduceLROnPlateau`` and the reference's checkpoint helpers accept
    it; ``state_dict()`` / ``load_state_dict()`` speak torch.optim.AdamW's format (per-parameter ``step`` / ``exp_avg`` /
    ``exp_avg_sq``), so optimizer checkpoints move between the reference and the engine in both directions.  The learning rate
    and the 1/world gradient scale are read by the kernel from a DEVICE vector that ``step()`` refreshes whenever
    ``param_groups[0]["lr"]`` changed -- a CUDA-graph-captured step therefore follows a scheduler."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        params = [p for p in (model.parameters() if isinstance(model, torch.nn.Module) else model) if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.params = params
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self.slices = []
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p.data)
            self.slices.append((off, k))
            off += k
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.hyper_dev = torch.tensor([lr, 1.0], dtype=torch.float32, device=dev)      # {lr, grad_scale} read by the kernel
        self._hyper_host = (float(lr), 1.0)

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def backward(self, loss):
        """``loss.backward()`` with the weight-gradient kernels of leaf parameters left running on the side stream until the flat
        gradient is gathered (ops.deferred_weight_gradients), then ``gather_grads()``."""
        with ops.deferred_weight_gradients():
            loss.backward()
        return self.gather_grads()

    def gather_grads(self):
        """autograd's per-parameter gradients -> the flat buffer (parameters nobody used contribute zeros)."""
        ops.join_side_streams()                     # weight-gradient kernels run on a side stream (ops.fork_join)
        gs = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        torch.cat(gs, out=self.flat_g)
        return self.flat_g

    def sync_hyper(self, grad_scale=None):
        """Push lr / grad_scale to the device if they changed (a tiny async H2D copy, outside any captured graph)."""
        want = (float(self.param_groups[0]["lr"]), self._hyper_host[1] if grad_scale is None else float(grad_scale))
        if want != self._hyper_host:
            self.hyper_dev.copy_(torch.tensor(want, dtype=torch.float32), non_blocking=False)
            self._hyper_host = want

    def step(self, grad_scale=1.0, closure=None):
        capturing = self.flat_p.is_cuda and torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.sync_hyper(grad_scale)
        elif float(grad_scale) != self._hyper_host[1]:
            raise RuntimeError("FlatAdamW: call sync_hyper(grad_scale) before capturing a step with a new gradient scale")
        ops.adamw_step(self.flat_p, self.flat_g, self.m, self.v, self.step_dev, self.param_groups[0]["lr"], self.betas[0],
                       self.betas[1], self.eps, self.weight_decay, grad_scale, hyper_dev=self.hyper_dev)

    # ---- torch.optim.AdamW checkpoint format ---------------------------------------------------------------
    def state_dict(self):
        state = {}
        for i, (off, k) in enumerate(self.slices):
            shp = self.params[i].shape
            state[i] = {"step": self.step_dev.detach().clone().reshape(()).cpu(), "exp_avg": self.m[off:off + k].view(shp).clone(),
                        "exp_avg_sq": self.v[off:off + k].view(shp).clone()}
        g = self.param_groups[0]
        group = {k: v for k, v in g.items() if k != "params"}
        group.update(betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, amsgrad=False, maximize=False,
                     params=list(range(len(self.params))))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        if "state" not in sd:                                       # round-1 flat format {m, v, step, lr}
            self.m.copy_(sd["m"])
            self.v.copy_(sd["v"])
            self.step_dev.copy_(sd["step"])
            self.param_groups[0]["lr"] = sd["lr"]
            return
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError("FlatAdamW.load_state_dict: %d parameters in the checkpoint, %d in the model" % (len(order), len(self.params)))
        step = None
        for j, (off, k) in zip(order, self.slices):
            st = sd["state"].get(j, sd["state"].get(str(j)))
            if st is None:
                self.m[off:off + k].zero_()
                self.v[off:off + k].zero_()
                continue
            self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            step = float(st["step"]) if step is None else max(step, float(st["step"]))
        self.step_dev.fill_(0.0 if step is None else step)
        g0 = groups[0]
        self.param_groups[0]["lr"] = g0["lr"]
        self.betas, self.eps, self.weight_decay = tuple(g0.get("betas", self.betas)), g0.get("eps", self.eps), g0.get("weight_decay", self.weight_decay)


class DistributedModel(torch.nn.Module):
    """Stand-in for the DDP wrapper: ``.module`` is the model; gradient averaging is done on the flat buffer
    by ``train_step`` (one all-reduce), not by per-bucket hooks."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, data):
        return self.module(data)


def get_distributed_model(model, verbosity=0, **_):
    if dist.is_initialized():
        # same starting point on every rank (DDP broadcasts rank 0's parameters at wrap time)
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    return DistributedModel(model)


def world():
    return (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)


def train_step(model, opt, data, compute_grad_energy=False, head_index=None):
    """forward -> loss -> backward -> flat all-reduce -> fused AdamW.  Returns (loss, tasks_loss)
    (hydragnn/train/train_validate_test.py:702-769 without the tracing scaffolding)."""
    m = model.module
    opt.zero_grad()
    if compute_grad_energy:
        data.pos.requires_grad_(True)
        pred = model(data)
        loss, tasks = m.energy_force_loss(pred, data)
    else:
        if head_index is None:
            head_index = get_head_indices(model, data)
        pred = model(data)
        loss, tasks = m.loss(pred, data.y, head_index)
    flat = opt.backward(loss)
    ws, _ = world()
    if ws > 1:
        dist.all_reduce(flat)                                       # the only collective of the step
    opt.step(grad_scale=1.0 / ws)
    return loss.detach(), [t.detach() for t in tasks]


def _capturable_allreduce(flat):
    """The step's only collective, on the CURRENT stream (NCCL collectives are CUDA-graph capturable)."""
    dist.all_reduce(flat)


class GraphedTrainStep:
    """CUDA-graph capture of ``train_step`` for a fixed-shape batch living in static device buffers: forward + loss + backward
    + flatten + the flat NCCL all-reduce + fused AdamW are ONE graph (`hydragnn/utils/distributed/distributed.py:479` and
    `train_validate_test.py:737-769` in one replay).  ``refill(data)`` copies a new batch of the same shape -- positions,
    features, targets AND topology (``edge_index`` / ``batch``) -- into the static buffers; the index plans (CSR views of
    ``edge_index``, graph offsets) are rebuilt INSIDE the captured region, so a refilled topology is honoured.  With
    ``neighbour_build=(radius, max_neighbours)`` the radius graph itself is part of the captured step (edge count promised by the
    warm-up run and verified on the device -- ``ops.check_guard``)."""

    def __init__(self, model, opt, static_data, compute_grad_energy=False, warmup=3, capture_allreduce=True):
        self.model, self.opt, self.data, self.mlip = model, opt, static_data, compute_grad_energy
        self.head_index = None if compute_grad_energy else get_head_indices(model, static_data)
        self.ws, _ = world()
        self.capture_allreduce = bool(capture_allreduce)
        self.opt.sync_hyper(1.0 / self.ws)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if self.ws > 1:
                    dist.all_reduce(self.opt.flat_g)
                self.opt.step(1.0 / self.ws)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_opt = None
        self.g_fb = torch.cuda.CUDAGraph()
        one_graph = self.ws == 1 or self.capture_allreduce
        with ops.capture_graph(self.g_fb):
            self._fwd_bwd()
            if one_graph:
                if self.ws > 1:
                    _capturable_allreduce(self.opt.flat_g)
                self.opt.step(1.0 / self.ws)
        if not one_graph:
            self.g_opt = torch.cuda.CUDAGraph()
            with ops.capture_graph(self.g_opt):
                self.opt.step(1.0 / self.ws)

    def _fwd_bwd(self):
        m = self.model.module
        d = self.data
        for k in ("_hgb_plan", "_hgb_gcsr"):          # index plans are part of the step (stale CSR after refill)
            d.__dict__.pop(k, None)
        self.opt.zero_grad()
        if self.mlip:
            d.pos.requires_grad_(True)
            pred = self.model(d)
            loss, _ = m.energy_force_loss(pred, d)
        else:
            pred = self.model(d)
            loss, _ = m.loss(pred, d.y, self.head_index)
        self.opt.backward(loss)
        self.loss = loss.detach()

    def refill(self, data):
        for k, v in data.items():
            if torch.is_tensor(v):
                dst = self.data[k]
                if dst.shape != v.shape:
                    raise ValueError("GraphedTrainStep.refill: %s has shape %s, the captured step has %s" % (k, tuple(v.shape), tuple(dst.shape)))
                dst.detach().copy_(v, non_blocking=True)

    def run(self):
        self.opt.sync_hyper()                       # a scheduler may have changed the learning rate
        self.g_fb.replay()
        if self.g_opt is not None:
            dist.all_reduce(self.opt.flat_g)
            self.g_opt.replay()
        return self.loss


FAST_TRAIN = os.environ.get("HGB_FAST_TRAIN", "1") == "1"


def _train_fast(loader, model, opt, compute_grad_energy, neighbour_build):
    """The epoch through ONE capacity-padded captured step (hydragnn_b200/padded.py): no per-step host synchronisation; the
    epoch sums stay on the device until the end."""
    from .padded import PaddedGraphStep
    dev = next(model.parameters()).device
    nbatch = len(loader)
    if dist.is_initialized():
        t = torch.tensor([nbatch], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        nbatch = int(t)
    fast = getattr(opt, "_hgb_fast", None)
    total = tasks_tot = None
    nsamp = 0
    for ibatch, data in enumerate(loader):
        if ibatch >= nbatch:
            break
        if fast is None or fast.model is not model or fast.mlip != bool(compute_grad_energy) or fast.nb != neighbour_build:
            fast = opt._hgb_fast = PaddedGraphStep(model, opt, data, compute_grad_energy, neighbour_build)
        g = fast.load(data)
        loss, tasks = fast.run()
        total = loss * g if total is None else total + loss * g
        tasks_tot = tasks * g if tasks_tot is None else tasks_tot + tasks * g
        nsamp += g
    fast.check()                                              # one host read per epoch: device guards of the captured steps
    return total / nsamp, tasks_tot / nsamp


def train(loader, model, opt, verbosity=0, profiler=None, use_deepspeed=False, compute_grad_energy=False, precision="fp32",
          fast=None, neighbour_build=None):
    """Epoch loop with the reference's signature and return values (train_error, tasks_error).

    ``fast`` (default: on for models ``padded.supported`` accepts): run every batch through one capacity-padded CUDA-graph step
    instead of eager launches -- same arithmetic, no per-step host sync.  ``neighbour_build`` = (radius, max_neighbours) builds
    the radius graph on the device inside that step; by default the batches carry ``edge_index`` as the reference's do."""
    resolve_precision(precision)
    dev = next(model.parameters()).device
    from . import padded
    if fast is None:
        fast = FAST_TRAIN and dev.type == "cuda" and padded.supported(model) and isinstance(opt, FlatAdamW)
    if fast:
        model.train()
        train_error, tasks_error = _train_fast(loader, model, opt, compute_grad_energy, neighbour_build)
        ws, _ = world()
        if ws > 1:
            dist.all_reduce(train_error)
            dist.all_reduce(tasks_error)
            train_error, tasks_error = train_error / ws, tasks_error / ws
        return train_error, tasks_error
    total, tasks_tot, nsamp = None, None, 0
    model.train()
    nbatch = len(loader)
    if dist.is_initialized():
        t = torch.tensor([nbatch], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        nbatch = int(t)
    for ibatch, data in enumerate(loader):
        if ibatch >= nbatch:
            break
        data = move_batch_to_device(data, torch.float32, dev)
        loss, tasks = train_step(model, opt, data, compute_grad_energy)
        g = data.num_graphs
        total = loss * g if total is None else total + loss * g
        tl = torch.stack(tasks) * g
        tasks_tot = tl if tasks_tot is None else tasks_tot + tl
        nsamp += g
    train_error, tasks_error = total / nsamp, tasks_tot / nsamp
    ws, _ = world()
    if ws > 1:
        dist.all_reduce(train_error)
        dist.all_reduce(tasks_error)
        train_error, tasks_error = train_error / ws, tasks_error / ws
    return train_error, tasks_error


@torch.no_grad()
def _eval_no_force(model, data, head_index):
    pred = model(data)
    return model.module.loss(pred, data.y, head_index), pred


def validate(loader, model, verbosity=0, reduce_ranks=True, compute_grad_energy=False, precision="fp32"):
    """hydragnn/train/train_validate_test.py:805-873 (losses only)."""
    dev = next(model.parameters()).device
    model.eval()
    total, tasks_tot, nsamp = None, None, 0
    for data in loader:
        data = move_batch_to_device(data, torch.float32, dev)
        if compute_grad_energy:
            with torch.enable_grad():
                data.pos.requires_grad_(True)
                pred = model(data)
                loss, tasks = model.module.energy_force_loss(pred, data, create_graph=False)
            loss, tasks = loss.detach(), [t.detach() for t in tasks]
        else:
            (loss, tasks), _ = _eval_no_force(model, data, get_head_indices(model, data))
        g = data.num_graphs
        total = loss * g if total is None else total + loss * g
        tl = torch.stack(tasks) * g
        tasks_tot = tl if tasks_tot is None else tasks_tot + tl
        nsamp += g
    err, terr = total / nsamp, tasks_tot / nsamp
    ws, _ = world()
    if reduce_ranks and ws > 1:
        dist.all_reduce(err)
        dist.all_reduce(terr)
        err, terr = err / ws, terr / ws
    return err, terr
