"""Import stub: physicsnemo/models/graphcast/graph_cast_processor.py:21 imports transformer_engine at
module level for its graph-transformer variant only; the message-passing processor never touches it."""
from . import pytorch  # noqa: F401
