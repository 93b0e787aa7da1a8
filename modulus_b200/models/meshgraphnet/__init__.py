from .meshgraphnet import MeshGraphNet, MeshGraphNetProcessor  # noqa: F401
