"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (`physicsnemo`, NVIDIA/modulus @ 878fd04) is imported from /root/reference with
the import shims of oracle/ref_shim (stubs for treelib/s3fs/timm and a pure-torch `dgl`
stand-in that only restates the gather and the segment sum).  Nothing here is imported by
the product or by the GPU box; the .pt files it writes are the only thing that travels.

Fixtures
  kat1_meshgraphnet_output.pt   the reference's own golden vector
                                (test/models/data/meshgraphnet_output.pth, [40,2] fp32)
  ref_mgn_<case>.pt             inputs, CSC graph, state_dict, output and ALL gradients of the
                                reference MeshGraphNet on small seeded cases
  ref_mgn_h128_L15.pt           default-size model (hidden 128, 15 layers): weights are NOT stored
                                (regenerated from the seed, same RNG stream), outputs + selected
                                gradients are
  ref_partitions.pt             GraphPartition fields produced by the reference partitioners
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import dgl  # noqa: E402  (the shim)
from physicsnemo.models.meshgraphnet import MeshGraphNet  # noqa: E402
from physicsnemo.models.gnn_layers import (  # noqa: E402
    partition_graph_by_coordinate_bbox,
    partition_graph_nodewise,
    partition_graph_with_id_mapping,
)
from physicsnemo.models.gnn_layers.distributed_graph import partition_graph_with_matrix_decomposition  # noqa: E402

from modulus_b200.mesh import random_graph_csc, triangle_grid_mesh  # noqa: E402

torch.set_num_threads(8)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def kat1():
    gold = torch.load("/root/reference/test/models/data/meshgraphnet_output.pth")
    # replay the recipe (test/models/meshgraphnet/test_meshgraphnet.py:41-65) as a self check
    torch.manual_seed(0)
    np.random.seed(0)
    model = MeshGraphNet(input_dim_nodes=4, input_dim_edges=3, output_dim=2)
    graphs = []
    for _ in range(2):
        src = torch.tensor([np.random.randint(20) for _ in range(10)])
        dst = torch.tensor([np.random.randint(20) for _ in range(10)])
        graphs.append(dgl.graph((src, dst)))
    graph = dgl.batch(graphs)
    nf = torch.randn(40, 4)
    ef = torch.randn(20, 3)
    with torch.no_grad():
        out = model(nf, ef, graph)
    err = (out - gold[0]).abs().max().item()
    assert err < 1e-5, err
    src, dst = graph.edges()
    save("kat1_meshgraphnet_output.pt", {"output": gold[0].clone(), "src": src.clone(), "dst": dst.clone(),
                                         "replay_max_abs_err": err})


def coo_graph(src, dst, n):
    return dgl.graph((src, dst), num_nodes=n)


def run_ref(model, nf, ef, src, dst, n_nodes, target):
    g = coo_graph(src, dst, n_nodes)
    nf = nf.clone().requires_grad_(True)
    ef = ef.clone().requires_grad_(True)
    model.zero_grad()
    out = model(nf, ef, g)
    loss = torch.nn.functional.mse_loss(out, target)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    return out.detach().clone(), loss.detach().clone(), grads, nf.grad.clone(), ef.grad.clone()


def small_cases():
    cases = {
        "relu_sum": dict(kw=dict(), seed=11),
        "relu_mean": dict(kw=dict(aggregation="mean"), seed=12),
        "silu_sum": dict(kw=dict(mlp_activation_fn="silu"), seed=13),
        "concat_trick": dict(kw=dict(do_concat_trick=True), seed=14),
    }
    for name, c in cases.items():
        torch.manual_seed(c["seed"])
        n_nodes, d_n, d_e, d_out = 37, 5, 3, 2
        offsets, indices = random_graph_csc(n_nodes, n_nodes, 1, 5, seed=c["seed"])
        deg = offsets[1:] - offsets[:-1]
        dst = torch.repeat_interleave(torch.arange(n_nodes), deg)
        src = indices.clone()
        model = MeshGraphNet(d_n, d_e, d_out, processor_size=2, hidden_dim_processor=32, hidden_dim_node_encoder=32,
                             hidden_dim_edge_encoder=32, hidden_dim_node_decoder=32, **c["kw"])
        nf = torch.randn(n_nodes, d_n)
        ef = torch.randn(src.numel(), d_e)
        target = torch.randn(n_nodes, d_out)
        out, loss, grads, gnf, gef = run_ref(model, nf, ef, src, dst, n_nodes, target)
        save(f"ref_mgn_{name}.pt", dict(
            kwargs=dict(input_dim_nodes=d_n, input_dim_edges=d_e, output_dim=d_out, processor_size=2,
                        hidden_dim_processor=32, hidden_dim_node_encoder=32, hidden_dim_edge_encoder=32,
                        hidden_dim_node_decoder=32, **c["kw"]),
            offsets=offsets, indices=indices, n_nodes=n_nodes, node_features=nf, edge_features=ef, target=target,
            state_dict={k: v.clone() for k, v in model.state_dict().items()},
            output=out, loss=loss, grads=grads, grad_node_features=gnf, grad_edge_features=gef))


def default_size_cases():
    """hidden 128: one message-passing layer (1e-4 bar) and 15 layers (1e-3 bar) on a small triangle
    mesh.  Weights are regenerated from the seed by the consumer (oracle.make_state_dict or
    modulus_b200 MeshGraphNet under torch.manual_seed) -- the same RNG stream as the reference."""
    for L, seed in ((1, 21), (15, 22)):
        mesh = triangle_grid_mesh(12, 13)
        n_nodes = mesh["num_nodes"]
        offsets, indices = mesh["offsets"], mesh["indices"]
        deg = offsets[1:] - offsets[:-1]
        dst = torch.repeat_interleave(torch.arange(n_nodes), deg)
        src = indices.clone()
        torch.manual_seed(seed)
        model = MeshGraphNet(6, 3, 3, processor_size=L)
        nf = torch.randn(n_nodes, 6)
        ef = torch.randn(src.numel(), 3)
        target = torch.randn(n_nodes, 3)
        out, loss, grads, gnf, gef = run_ref(model, nf, ef, src, dst, n_nodes, target)
        keep = [k for k in grads if k.startswith(("edge_encoder.model.0", "node_decoder.model.4",
                                                   "processor.processor_layers.0.edge_mlp.model.0",
                                                   "processor.processor_layers.0.edge_mlp.model.5",
                                                   f"processor.processor_layers.{2 * L - 1}.node_mlp.model.4"))]
        sel = {k: grads[k] for k in keep}
        norms = {k: float(v.double().norm()) for k, v in grads.items()}
        save(f"ref_mgn_h128_L{L}.pt", dict(
            seed=seed, processor_size=L, grid=(12, 13), offsets=offsets, indices=indices, n_nodes=n_nodes,
            node_features=nf, edge_features=ef, target=target, output=out, loss=loss,
            grads_selected=sel, grad_norms=norms, grad_node_features=gnf, grad_edge_features=gef,
            weight_checksum=float(sum(v.double().sum() for v in model.state_dict().values()))))


def gp_to_dict(gp):
    d = {}
    for k, v in vars(gp).items():
        if k == "device":
            continue
        if isinstance(v, torch.Tensor):
            d[k] = v.clone()
        elif isinstance(v, list):
            d[k] = [x.clone() if isinstance(x, torch.Tensor) else
                    ([int(y) for y in x] if isinstance(x, list) else (int(x) if x is not None else None)) for x in v]
        else:
            d[k] = int(v) if isinstance(v, (bool, int)) or torch.is_tensor(v) else v
    return d


def partitions():
    out = {}
    # bipartite random graph (recipe of test/models/test_distributed_graph.py:26-51, smaller)
    off_b, idx_b = random_graph_csc(61, 43, 1, 6, seed=42)
    # square graph
    off_s, idx_s = random_graph_csc(50, 50, 1, 5, seed=7)
    out["graphs"] = {"bipartite": (off_b, idx_b, 61, 43), "square": (off_s, idx_s, 50, 50)}
    res = {}
    for gname, (off, idx, ns, nd) in out["graphs"].items():
        ns_eff = int(idx.max()) + 1
        for P in (2, 3, 4):
            for r in range(P):
                res[(gname, "nodewise", P, r)] = gp_to_dict(partition_graph_nodewise(off, idx, P, r, "cpu"))
                g = torch.Generator().manual_seed(100 + P)
                ms = torch.randint(0, P, (ns_eff,), generator=g)
                md = torch.randint(0, P, (nd,), generator=g)
                ms[:P] = torch.arange(P)
                md[:P] = torch.arange(P)
                res[(gname, "mapping", P, r)] = gp_to_dict(
                    partition_graph_with_id_mapping(off, idx, ms, md, P, r, "cpu"))
                res[(gname, "mapping", P, r)]["_mapping_src"] = ms
                res[(gname, "mapping", P, r)]["_mapping_dst"] = md
                if P in (2, 4):
                    g = torch.Generator().manual_seed(200 + P)
                    cs = torch.rand(ns_eff, 2, generator=g) * 2 - 1
                    cd = torch.rand(nd, 2, generator=g) * 2 - 1
                    if P == 2:
                        cmin = [[None, None], [0.0, None]]
                        cmax = [[0.0, None], [None, None]]
                    else:
                        cmin = [[0, 0], [None, 0], [None, None], [0, None]]
                        cmax = [[None, None], [0, None], [0, 0], [None, 0]]
                    d = gp_to_dict(partition_graph_by_coordinate_bbox(off, idx, cs, cd, cmin, cmax, P, r, "cpu"))
                    d.update(_src_coordinates=cs, _dst_coordinates=cd, _cmin=cmin, _cmax=cmax)
                    res[(gname, "bbox", P, r)] = d
                if gname == "square":
                    res[(gname, "matrix_decomp", P, r)] = gp_to_dict(
                        partition_graph_nodewise(off, idx, P, r, "cpu", matrix_decomp=True))
    out["partitions"] = res
    save("ref_partitions.pt", out)


if __name__ == "__main__":
    kat1()
    small_cases()
    default_size_cases()
    partitions()
