"""Golden fixture for the step BEFORE the path: mesh cells -> bidirected graph -> edge
features -> normalisation, produced by the reference's own static methods
(physicsnemo/datapipes/gnn/vortex_shedding_dataset.py:307-349) under the import shims of oracle/ref_shim
(tensorflow / dgl stubs; `dgl.to_bidirected` is DGL's documented semantics restated in the shim).

    python tests/golden/make_golden_datapipe.py        (build container only; needs /root/reference)

  ref_datapipe_graph.pt    2-D Delaunay mesh (seeded points) and the same connectivity with 3-D coordinates:
                           cells, pos, reference edge list (src, dst), edge features, normalised features
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(1, "/root/reference")
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.spatial import Delaunay  # noqa: E402

from physicsnemo.datapipes.gnn.vortex_shedding_dataset import VortexSheddingDataset as D  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    pts = rng.random((120, 2)).astype(np.float32)
    cells = Delaunay(pts).simplices.astype(np.int64)
    out = {"cells": torch.from_numpy(cells)}
    for dim in (2, 3):
        pos = torch.from_numpy(pts if dim == 2 else np.concatenate([pts, rng.random((120, 1)).astype(np.float32)], 1))
        src, dst = D.cell_to_adj(cells)
        g = D.create_graph(src, dst, dtype=torch.int32)
        g = D.add_edge_features(g, pos)
        ef = g.edata["x"].clone()
        mu = ef.mean(0)
        std = ef.std(0)
        efn = D.normalize_edge(g, mu, std)
        s, d = g.edges()
        out[f"dim{dim}"] = dict(pos=pos, src=s.clone(), dst=d.clone(), edge_features=ef, mu=mu, std=std,
                                edge_features_normalized=efn.clone())
    path = os.path.join(HERE, "ref_datapipe_graph.pt")
    torch.save(out, path)
    print(f"wrote ref_datapipe_graph.pt: {os.path.getsize(path) / 1024:.1f} KiB, E = {out['dim2']['src'].numel()}")


if __name__ == "__main__":
    main()
