"""GNN building blocks of the MeshGraphNet path (reference: physicsnemo/models/gnn_layers)."""
from .distributed_graph import (  # noqa: F401
    DistributedGraph,
    GraphPartition,
    partition_graph_by_coordinate_bbox,
    partition_graph_nodewise,
    partition_graph_with_id_mapping,
    partition_graph_with_matrix_decomposition,
)
from .graph import CuGraphCSC  # noqa: F401
from .halo_partition import HaloPartition, partition_ids_by_slabs, partition_with_halo  # noqa: F401
from .mesh_edge_block import MeshEdgeBlock  # noqa: F401
from .mesh_graph_decoder import MeshGraphDecoder  # noqa: F401
from .mesh_graph_encoder import MeshGraphEncoder  # noqa: F401
from .mesh_graph_mlp import MeshGraphEdgeMLPConcat, MeshGraphEdgeMLPSum, MeshGraphMLP  # noqa: F401
from .mesh_node_block import MeshNodeBlock  # noqa: F401
from .utils import aggregate_and_concat, concat_efeat, sum_efeat  # noqa: F401
