"""modulus_b200 -- B200-native (sm_100a) MeshGraphNet message-passing path.

Keeps the reference's public names for this path
(physicsnemo.models.meshgraphnet.MeshGraphNet, physicsnemo.models.gnn_layers.*,
physicsnemo.distributed.{indexed_all_to_all_v, mark_module_as_shared, ...}) on top of one
C-ABI CUDA library (include/mgn_b200.h).  There is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
