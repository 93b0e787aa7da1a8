from .meshgraphnet import MeshGraphNet  # noqa: F401
