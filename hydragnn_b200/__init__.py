"""hydragnn-b200: a Blackwell-native message-passing engine behind HydraGNN's ``mpnn_type`` plugin API.

Public surface (mirrors the slice of ``hydragnn`` that sits on the per-step hot path):

    create_model, create_model_config      -- hydragnn.models.create
    Data, Batch                            -- torch_geometric.data stand-ins
    get_radius_graph[_pbc][_config]        -- hydragnn.preprocess.graph_samples_checks_and_updates
    train, validate, train_step, FlatAdamW, get_distributed_model -- hydragnn.train / hydragnn.utils.distributed

The CUDA library is loaded lazily on first use; importing the package works on a CPU-only host.
"""
from .data import Batch, Data, collate_to_device  # noqa: F401
from .create import create_model, create_model_config, get_device, set_precision  # noqa: F401
from .radius import (get_radius_graph, get_radius_graph_config, get_radius_graph_pbc,  # noqa: F401
                     get_radius_graph_pbc_config, RadiusGraph, RadiusGraphPBC)
from .train import (FlatAdamW, GraphedTrainStep, get_distributed_model, get_head_indices, train,  # noqa: F401
                    train_step, validate)

from .padded import PaddedGraphStep  # noqa: F401

__version__ = "0.2.0"
