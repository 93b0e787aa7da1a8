"""GraphCast message-passing pieces that sit on the MeshGraphNet operator seam."""
from .graph_cast_processor import GraphCastProcessor  # noqa: F401
from .partition import get_lat_lon_partition_separators  # noqa: F401
