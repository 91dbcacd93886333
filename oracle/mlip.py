"""Oracle: interatomic-potential wrapper and loss.  Test infrastructure only.

Restates ``EnhancedModelWrapper.energy_force_loss``
(hydragnn/models/create.py:590-738).
"""
import torch
from torch import nn

from .geometry import segment_sum


class MLIPWrapper(nn.Module):
    def __init__(self, model, energy_weight, energy_peratom_weight, force_weight):
        super().__init__()
        self.model = model
        self.energy_weight, self.energy_peratom_weight, self.force_weight = (
            energy_weight, energy_peratom_weight, force_weight)

    def __getattr__(self, name):                      # delegation, create.py:600-622
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("model"), name)

    def forward(self, data):
        return self.model(data)

    def energy_force_loss(self, pred, data, create_graph=True):
        assert data.pos.requires_grad, "data.pos does not have grad"
        assert self.num_heads == 1, "Force predictions require exactly one head."
        G = int(data.batch.max()) + 1
        if self.head_type[0] == "node":               # :651-657  (torch_scatter.scatter_add)
            e_pred = segment_sum(pred[0], data.batch, G).squeeze().float()
        else:                                         # :658-668
            if self.model.graph_pooling != "add":
                raise ValueError("Graph head force loss requires sum pooling (graph_pooling='add').")
            e_pred = pred[0].squeeze().float()
        e_true = data.energy.squeeze().float()
        lf = self.loss_function
        tasks = [lf(e_pred, e_true)]
        if self.energy_weight <= 0 and self.energy_peratom_weight <= 0 and self.force_weight <= 0:
            raise ValueError("All interatomic potential loss weights are zero")
        tot = 0
        if self.energy_weight > 0:
            tot = tot + lf(e_pred, e_true) * self.energy_weight
        natoms = torch.bincount(data.batch)           # :699-706
        tasks.append(lf(e_pred / natoms, e_true / natoms))
        if self.energy_peratom_weight > 0:
            tot = tot + lf(e_pred / natoms, e_true / natoms) * self.energy_peratom_weight
        f_pred = -torch.autograd.grad(                # :718-728
            e_pred, data.pos, grad_outputs=torch.ones_like(e_pred),
            retain_graph=e_pred.requires_grad, create_graph=create_graph)[0].float()
        tasks.append(lf(f_pred, data.forces.float()))
        if self.force_weight > 0:
            tot = tot + lf(f_pred, data.forces.float()) * self.force_weight
        return tot, tasks
