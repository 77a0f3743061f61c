"""tec_mollm_b200 -- B200-native (sm_100a) drop-in for TEC-MoLLM's GATv2 SpatialEncoder and its haversine
graph builder.  Python here is plumbing (tensors, autograd wiring, torch.distributed); every number is
computed by the hand-written CUDA library ``lib/libtecgat.so`` behind the C ABI in ``include/tecgat.h``."""
from .gatv2 import GATv2Conv, GraphPlan, tile_nodes_for  # noqa: F401
from .spatial_encoder import SpatialEncoder  # noqa: F401
from .embedding import SpatioTemporalEmbedding  # noqa: F401
from .temporal import MultiScaleConvEmbedder, Multi_Scale_Conv_Block  # noqa: F401
from . import graph  # noqa: F401
from . import dist  # noqa: F401

__all__ = ["GATv2Conv", "GraphPlan", "SpatialEncoder", "SpatioTemporalEmbedding", "Multi_Scale_Conv_Block", "MultiScaleConvEmbedder", "graph", "dist", "tile_nodes_for"]
