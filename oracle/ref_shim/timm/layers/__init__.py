def to_2tuple(x):  # pragma: no cover
    return (x, x) if not isinstance(x, (tuple, list)) else tuple(x)
