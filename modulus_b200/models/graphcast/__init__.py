"""GraphCast message-passing pieces that sit on the MeshGraphNet operator seam."""
from .graph_cast_processor import GraphCastProcessor  # noqa: F401
