"""Activation lookup for the MeshGraphNet path.

Mirrors `get_activation` of the reference (physicsnemo/models/layers/activations.py:173-222):
the string table returns torch.nn modules so parameter-free state_dicts stay identical.
`activation_name` maps a module back to the id the CUDA kernels implement.
"""
import torch.nn as nn

ACT2FN = {
    "relu": nn.ReLU,
    "leaky_relu": (nn.LeakyReLU, {"negative_slope": 0.1}),
    "elu": nn.ELU,
    "silu": nn.SiLU,
    "gelu": nn.GELU,
    "sigmoid": nn.Sigmoid,
    "tanh": nn.Tanh,
    "identity": nn.Identity,
}


def get_activation(activation: str) -> nn.Module:
    try:
        activation = activation.lower()
        module = ACT2FN[activation]
    except (KeyError, AttributeError):
        raise KeyError(
            f"Activation function {activation} not found. Available options are: {list(ACT2FN.keys())}"
        )
    if isinstance(module, tuple):
        return module[0](**module[1])
    return module()


def activation_name(fn) -> str:
    """Kernel id of an activation module; raises for anything the kernels do not implement."""
    if fn is None or isinstance(fn, nn.Identity):
        return "identity"
    if isinstance(fn, str):
        return fn.lower()
    table = {nn.ReLU: "relu", nn.SiLU: "silu", nn.Tanh: "tanh", nn.Sigmoid: "sigmoid", nn.ELU: "elu"}
    for cls, name in table.items():
        if type(fn) is cls:
            if cls is nn.ELU and fn.alpha != 1.0:
                break
            return name
    if type(fn) is nn.GELU and getattr(fn, "approximate", "none") == "none":
        return "gelu"
    if type(fn) is nn.LeakyReLU and abs(fn.negative_slope - 0.1) < 1e-12:
        return "leaky_relu"
    raise NotImplementedError(f"modulus_b200 has no CUDA kernel for activation {fn!r}")
