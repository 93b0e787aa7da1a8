from .activations import get_activation, activation_name  # noqa: F401
