"""dgl.function stand-ins: message/reduce descriptors consumed by DGLGraph.update_all."""


def copy_e(e, out):
    return ("copy_e", e, out)


def sum(msg, out):  # noqa: A001
    return ("sum", msg, out)


def mean(msg, out):
    return ("mean", msg, out)
