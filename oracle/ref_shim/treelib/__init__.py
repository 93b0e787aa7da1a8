"""Stub of `treelib` so the unmodified reference (`physicsnemo/distributed/config.py:19`)
imports in this container.  Test infrastructure only."""


class Tree:  # pragma: no cover - never exercised on the MeshGraphNet path
    def __init__(self, *a, **k):
        pass
