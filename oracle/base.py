"""Oracle: model assembly (encoder loop, pooling, multi-head decoder).

Test infrastructure only.  Restates ``Base`` (hydragnn/models/Base.py:36-982),
``EGCLStack`` (hydragnn/models/EGCLStack.py:22-152), ``PAINNStack``
(hydragnn/models/PAINNStack.py:27-191) and the ``create_model`` dispatch
(hydragnn/models/create.py:112-584) for the EGNN and PAINN stacks without GPS.
Parameter names follow the reference (PyG ``Sequential`` names its children
``module_<i>`` [3P-memory B.5]) so state dicts interchange with the engine.
"""
import torch
from torch import nn

from .egnn import EGCL
from .geometry import edge_vectors_and_lengths, graph_pool
from .painn import PainnMessage, PainnUpdate
from . import pnaeq
from .gps import GPSConv


def activation(name):
    """hydragnn/utils/model/model.py:30-46."""
    table = {
        "relu": nn.ReLU, "selu": nn.SELU, "prelu": nn.PReLU, "elu": nn.ELU,
        "lrelu_01": lambda: nn.LeakyReLU(0.1), "lrelu_025": lambda: nn.LeakyReLU(0.25),
        "lrelu_05": lambda: nn.LeakyReLU(0.5), "sigmoid": nn.Sigmoid,
    }
    return table[name]() if name in table else None


def loss_function(name):
    """hydragnn/utils/model/model.py:49-62."""
    F = torch.nn.functional
    if name == "mse":
        return F.mse_loss
    if name == "mae":
        return F.l1_loss
    if name == "rmse":
        return lambda a, b: torch.sqrt(F.mse_loss(a, b))
    raise ValueError("oracle supports mse / mae / rmse, got " + str(name))


def normalize_heads(output_heads):
    """``update_multibranch_heads`` (hydragnn/utils/model/model.py:314-349)."""
    out = {}
    for key, val in output_heads.items():
        out[key] = val if isinstance(val, list) else [{"type": "branch-0", "architecture": val}]
    return out


class _Conv(nn.Module):
    """Stand-in for the PyG ``Sequential`` built by ``get_conv``; holds children under
    the names PyG would give them."""

    def __init__(self, kind, mods):
        super().__init__()
        self.kind = kind
        for i, m in enumerate(mods):
            if m is not None:
                self.add_module("module_%d" % i, m)


class OracleModel(nn.Module):
    def __init__(self, mpnn_type, input_dim, hidden_dim, output_dim, output_type, output_heads,
                 activation_function="relu", loss_function_type="mse", task_weights=None,
                 num_conv_layers=2, num_nodes=None, edge_dim=None, num_radial=None, radius=None,
                 equivariance=False, graph_pooling="mean", pna_deg=None, global_attn_engine=None,
                 global_attn_type=None, global_attn_heads=0, pe_dim=0, dropout=0.25, **_unused):
        super().__init__()
        self.use_global_attn = bool(global_attn_engine)
        if self.use_global_attn and (global_attn_engine != "GPS" or global_attn_type != "multihead"):
            raise ValueError("oracle supports global_attn_engine='GPS' with global_attn_type='multihead'")
        self.global_attn_heads, self.pe_dim, self.dropout = global_attn_heads, pe_dim, dropout
        if mpnn_type == "PNAEq":
            assert pna_deg is not None, "PNAEq requires degree input."
            self.deg = pnaeq.sanitize_degree(pna_deg)
        if mpnn_type not in ("EGNN", "PAINN", "PNAEq"):
            raise ValueError("Unknown mpnn_type: {0}".format(mpnn_type))
        self.mpnn_type, self.input_dim, self.hidden_dim = mpnn_type, input_dim, hidden_dim
        self.head_dims, self.head_type = list(output_dim), list(output_type)
        self.num_heads = len(self.head_dims)
        self.config_heads = normalize_heads(output_heads)
        self.activation_function = activation(activation_function)
        self.loss_function = loss_function(loss_function_type)
        w = list(task_weights if task_weights is not None else [1.0] * self.num_heads)
        if len(w) != self.num_heads:
            raise ValueError("Inconsistent number of loss weights and tasks")
        tot = sum(abs(t) for t in w)
        self.loss_weights = [t / tot for t in w]                           # Base.py:121-132
        mode = graph_pooling.lower()
        self.graph_pooling = "add" if mode == "sum" else mode
        self.num_conv_layers, self.num_radial, self.radius = num_conv_layers, num_radial, radius
        self.equivariance = bool(equivariance)
        if mpnn_type == "EGNN":
            self.edge_dim = 0 if edge_dim is None else edge_dim            # EGCLStack.py:33-35
        else:
            self.edge_dim = edge_dim                                        # PAINNStack.py:43
        self.use_edge_attr = self.edge_dim is not None and self.edge_dim > 0   # Base.py:135-141
        if self.use_global_attn:                                               # Base.py:179-215
            self.embed_dim = self.edge_embed_dim = hidden_dim
            self.pos_emb = nn.Linear(pe_dim, hidden_dim, bias=False)
            if input_dim:
                self.node_emb = nn.Linear(input_dim, hidden_dim, bias=False)
                self.node_lin = nn.Linear(2 * hidden_dim, hidden_dim, bias=False)
            self.rel_pos_emb = nn.Linear(pe_dim, hidden_dim, bias=False)
            if self.use_edge_attr:
                self.edge_emb = nn.Linear(self.edge_dim, hidden_dim, bias=False)
                self.edge_lin = nn.Linear(2 * hidden_dim, hidden_dim, bias=False)
        else:
            self.embed_dim, self.edge_embed_dim = input_dim, self.edge_dim

        # --- conv stack: first layer at embed_dim (= input_dim without GPS, Q4), last layer flagged (EGCLStack.py:45-70)
        self.graph_convs = nn.ModuleList()
        for i in range(num_conv_layers):
            last = i == num_conv_layers - 1
            conv = self._get_conv(self.embed_dim if i == 0 else hidden_dim, hidden_dim, last)
            if self.use_global_attn:                                           # Base._apply_global_attn :234-247
                conv = GPSConv(hidden_dim, conv, heads=global_attn_heads, dropout=dropout)
            self.graph_convs.append(conv)

        # --- decoder (Base.py:590-691), single or multi branch
        act = self.activation_function
        self.heads_NN = nn.ModuleList()          # registered before graph_shared, as in Base.__init__:83
        self.convs_node_hidden, self.batch_norms_node_hidden = nn.ModuleDict(), nn.ModuleDict()       # Base.py:88-91
        self.convs_node_output, self.batch_norms_node_output = nn.ModuleDict(), nn.ModuleDict()
        self.graph_shared = nn.ModuleDict()
        self.num_branches = 1
        if "graph" in self.config_heads:
            self.num_branches = len(self.config_heads["graph"])
            for br in self.config_heads["graph"]:
                a = br["architecture"]
                layers = [nn.Linear(hidden_dim, a["dim_sharedlayers"]), act]
                for _ in range(a["num_sharedlayers"] - 1):
                    layers += [nn.Linear(a["dim_sharedlayers"], a["dim_sharedlayers"]), act]
                self.graph_shared[br["type"]] = nn.Sequential(*layers)
        if "node" in self.config_heads:
            self._init_node_conv(num_nodes)
        inode = 0
        for ih in range(self.num_heads):
            head = nn.ModuleDict()
            if self.head_type[ih] == "graph":
                for br in self.config_heads["graph"]:
                    a = br["architecture"]
                    dims = [a["dim_sharedlayers"]] + list(a["dim_headlayers"][: a["num_headlayers"]])
                    layers = []
                    for d0, d1 in zip(dims[:-1], dims[1:]):
                        layers += [nn.Linear(d0, d1), act]
                    layers.append(nn.Linear(dims[-1], self.head_dims[ih]))
                    head[br["type"]] = nn.Sequential(*layers)
            elif self.head_type[ih] == "node":
                for br in self.config_heads["node"]:
                    a = br["architecture"]
                    if a["type"] in ("mlp", "mlp_per_node"):                       # Base.py:648-664
                        per_node = a["type"] == "mlp_per_node"
                        if per_node:
                            assert num_nodes is not None, "num_nodes must be provided for mlp_per_node; use 'mlp' for variable-size graphs"
                        head[br["type"]] = _MLPNode(hidden_dim, self.head_dims[ih], a["dim_headlayers"], act,
                                                    num_mlp=num_nodes if per_node else 1, num_nodes=num_nodes if per_node else None)
                    elif a["type"] == "conv":                                       # Base.py:665-680: the SAME modules, listed again
                        key, mods = br["type"], nn.ModuleList()
                        for conv, bn in zip(self.convs_node_hidden[key], self.batch_norms_node_hidden[key]):
                            mods.append(conv)
                            mods.append(bn)
                        mods.append(self.convs_node_output[key][inode])
                        mods.append(self.batch_norms_node_output[key][inode])
                        head[key] = mods
                        inode += 1
                    else:
                        raise ValueError("Unknown head NN structure for node features" + a["type"])
            else:
                raise ValueError("Unknown head type" + str(self.head_type[ih]))
            self.heads_NN.append(head)

    def _init_node_conv(self, num_nodes):
        """Base._init_node_conv (:508-588): conv-type node heads share their hidden convolutions between heads."""
        from .gps import PyGBatchNorm
        cfgs = self.config_heads["node"]
        if any(br["architecture"]["type"] != "conv" for br in cfgs):
            return
        node_heads = [i for i, t in enumerate(self.head_type) if t == "node"]
        if not node_heads:
            return
        for br in cfgs:
            a = br["architecture"]
            hid = a["dim_headlayers"]
            ch, bh, co, bo = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
            ch.append(self._get_conv(self.hidden_dim, hid[0], False))
            bh.append(PyGBatchNorm(hid[0]))
            for k in range(a["num_headlayers"] - 1):
                ch.append(self._get_conv(hid[k], hid[k + 1], False))
                bh.append(PyGBatchNorm(hid[k + 1]))
            for ih in node_heads:
                co.append(self._get_conv(hid[-1], self.head_dims[ih], True))
                bo.append(PyGBatchNorm(self.head_dims[ih]))
            key = br["type"]
            self.convs_node_hidden[key], self.batch_norms_node_hidden[key] = ch, bh
            self.convs_node_output[key], self.batch_norms_node_output[key] = co, bo

    # EGCLStack.get_conv :72-109 / PAINNStack.get_conv :76-147
    def _get_conv(self, fin, fout, last):
        ed = self.edge_embed_dim                                        # hidden_dim under GPS, else the stack's edge_dim
        if self.mpnn_type == "EGNN":
            return _Conv("egnn", [EGCL(fin, fout, self.hidden_dim, edge_attr_dim=ed or self.edge_dim,
                                       equivariant=self.equivariance and not last)])
        if self.mpnn_type == "PNAEq":                                   # PNAEqStack.get_conv :119-192
            msg = pnaeq.PainnMessage(fin, self.deg, ed, self.num_radial)
            upd = pnaeq.PainnUpdate(fin, last_layer=last)
        else:
            msg = PainnMessage(fin, self.num_radial, self.radius, edge_dim=ed)
            upd = PainnUpdate(fin, last_layer=last)
        s_out = nn.Sequential(nn.Linear(fin, fout), nn.Tanh(), nn.Linear(fout, fout))
        v_out = None if last else nn.Linear(fin, fout)
        return _Conv("painn", [msg, upd, s_out, v_out])

    def forward(self, data):
        x, pos, ei = data.x, data.pos, data.edge_index.to(torch.long)
        shifts = getattr(data, "edge_shifts", None)
        if shifts is None:                                                   # Base.py:466-469
            shifts = torch.zeros(ei.shape[1], 3, dtype=pos.dtype, device=pos.device)
        eattr = data.edge_attr if self.use_edge_attr else None
        if self.use_global_attn:                                               # Base._embedding :477-491
            xe = self.pos_emb(data.pe)
            if self.input_dim:
                xe = self.node_lin(torch.cat((self.node_emb(x.float()), xe), 1))
            e = self.rel_pos_emb(data.rel_pe)
            if self.use_edge_attr:
                e = self.edge_lin(torch.cat((self.edge_emb(eattr), e), 1))
            x, eattr = xe, e

        def layer(conv, fn):
            """run one conv, through GPSConv when global attention is on"""
            if self.use_global_attn:
                return lambda x_, q_: conv(x_, q_, lambda a, b: fn(conv.conv, a, b))
            return lambda x_, q_: fn(conv, x_, q_)

        if self.mpnn_type == "EGNN":
            equiv = pos
            run_conv = lambda c, a, b: c.module_0(a, b, ei, eattr, shifts)
            for conv in self.graph_convs:
                x, equiv = layer(conv, run_conv)(x, equiv)
                x = self.activation_function(x)                              # Base.py:726
        elif self.mpnn_type == "PNAEq":
            vec, dist = edge_vectors_and_lengths(pos, ei, shifts, normalize=True)    # PNAEqStack.py:202-205
            rbf = pnaeq.rbf_basis(dist.squeeze(-1), self.num_radial, self.radius)
            edge = ei.t()
            v = torch.zeros(x.shape[0], 3, x.shape[1], dtype=x.dtype, device=x.device)

            def pna_conv(c, a, b):
                a, b2 = c.module_0(a, b, edge, rbf, vec, eattr)
                a, b3 = c.module_1(a, b2)
                a = c.module_2(a)
                return a, (c.module_3(b3) if b3 is not None else b2)

            run_conv = pna_conv
            for conv in self.graph_convs:
                x, v = layer(conv, pna_conv)(x, v)
                x = self.activation_function(x)
            equiv = v
        else:
            diff, dist = edge_vectors_and_lengths(pos, ei, shifts, normalize=True)   # PAINNStack.py:157-159
            edge = ei.t()
            v = torch.zeros(x.shape[0], 3, x.shape[1], dtype=x.dtype, device=x.device)

            def painn_conv(c, a, b):
                a, b2 = c.module_0(a, b, edge, diff, dist, eattr)
                a, b3 = c.module_1(a, b2)
                a = c.module_2(a)
                return a, (c.module_3(b3) if b3 is not None else b2)

            run_conv = painn_conv
            for conv in self.graph_convs:
                x, v = layer(conv, painn_conv)(x, v)
                x = self.activation_function(x)
            equiv = v
        batch = getattr(data, "batch", None)
        if batch is None:
            batch = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
        G = int(batch.max()) + 1
        xg = graph_pool(x, batch, G, self.graph_pooling)                     # Base.py:733-738
        ds = getattr(data, "dataset_name", None)
        outs = []
        for ih, (hd, head, kind) in enumerate(zip(self.head_dims, self.heads_NN, self.head_type)):
            if self.num_branches == 1:
                if kind == "graph":
                    outs.append(head["branch-0"](self.graph_shared["branch-0"](xg))[:, :hd])
                elif isinstance(head["branch-0"], nn.ModuleList):          # conv-type node head (Base.py:800-810)
                    a, b = x, equiv
                    mods = head["branch-0"]
                    for conv, bn in zip(mods[0::2], mods[1::2]):
                        a, b = run_conv(conv, a, b)
                        a = self.activation_function(bn(a))
                    outs.append(a[:, :hd])
                else:
                    outs.append(head["branch-0"](x, batch)[:, :hd])
                continue
            # multi-branch masking (Base.py:770-780, 816-840)
            ids = ds[:, 0]
            if kind == "graph":
                out = x.new_zeros(G, hd)
                for b in ids.unique():
                    m = ids == b
                    key = "branch-%d" % int(b)
                    out[m] = head[key](self.graph_shared[key](xg[m]))[:, :hd]
            else:
                out = x.new_zeros(x.shape[0], hd)
                for b in ids.unique():
                    m = (ids == b)[batch]
                    if isinstance(head["branch-%d" % int(b)], nn.ModuleList):
                        raise ValueError("oracle: conv-type node heads with several branches are not restated")
                    out[m] = head["branch-%d" % int(b)](x[m], batch[m])[:, :hd]
            outs.append(out)
        return outs

    def loss(self, pred, value, head_index):
        """``loss_hpweighted`` (Base.py:879-906)."""
        tot, tasks = 0, []
        for ih in range(self.num_heads):
            tgt = value[head_index[ih]].reshape(pred[ih].shape)
            li = self.loss_function(pred[ih], tgt)
            tot = tot + li * self.loss_weights[ih]
            tasks.append(li)
        return tot, tasks


class _MLPNode(nn.Module):
    """``MLPNode`` (Base.py:912-979): one shared MLP ('mlp') or one MLP per node position ('mlp_per_node', graphs of exactly
    ``num_nodes`` atoms: node i of every graph goes through ``mlp[i]``)."""

    def __init__(self, fin, fout, hidden, act, num_mlp=1, num_nodes=None):
        super().__init__()
        self.num_nodes, self.fout = num_nodes, fout
        self.mlp = nn.ModuleList()
        for _ in range(num_mlp):
            dims = [fin] + list(hidden)
            layers = []
            for d0, d1 in zip(dims[:-1], dims[1:]):
                layers += [nn.Linear(d0, d1), act]
            layers.append(nn.Linear(dims[-1], fout))
            self.mlp.append(nn.Sequential(*layers))

    def forward(self, x, batch=None):
        if self.num_nodes is None:
            return self.mlp[0](x)
        outs = x.new_zeros(x.shape[0], self.fout)
        for i in range(self.num_nodes):
            outs[i::self.num_nodes] = self.mlp[i](x[i::self.num_nodes])
        return outs


def create_model(**kw):
    """Mirror of ``create_model`` (hydragnn/models/create.py:112-766): seeds the RNG
    (:164) and wraps the stack for MLIP training when asked (:586-756)."""
    from .mlip import MLIPWrapper
    torch.manual_seed(0)
    if kw.get("mpnn_type") == "MACE":                      # create.py:542-582 -> MACEStack
        from .mace import MACEOracle
        model = MACEOracle(**{k: v for k, v in kw.items() if k != "mpnn_type"})
    else:
        model = OracleModel(**kw)
    if kw.get("enable_interatomic_potential", False):
        model = MLIPWrapper(model, kw.get("energy_weight", 0.0), kw.get("energy_peratom_weight", 0.0),
                            kw.get("force_weight", 0.0))
    return model
