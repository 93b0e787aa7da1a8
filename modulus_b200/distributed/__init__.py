"""Model-parallel communication of the partitioned-graph path
(reference exports: physicsnemo/distributed/__init__.py:19-33)."""
from .autograd import all_gather_v, gather_v, indexed_all_to_all_v, scatter_v  # noqa: F401
from .manager import (  # noqa: F401
    DistributedManager,
    PhysicsNeMoUndefinedGroupError,
    PhysicsNeMoUninitializedDistributedManagerWarning,
)
from .utils import (  # noqa: F401
    mark_module_as_shared,
    reduce_loss,
    reduce_shared_gradients,
    unmark_module_as_shared,
)
