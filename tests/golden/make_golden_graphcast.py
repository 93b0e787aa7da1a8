"""Golden fixtures for the GraphCast blocks on the MeshGraphNet operator seam,
produced by the UNMODIFIED reference under the import shims of oracle/ref_shim:

    python tests/golden/make_golden_graphcast.py      (build container only; needs /root/reference)

  ref_graphcast_blocks.pt   MeshGraphEncoder / MeshGraphDecoder on a random bipartite graph
                            (N_src != N_dst, recipe of test/models/graphcast/test_graphcast_snmg.py
                            scaled down) and GraphCastProcessor on a square graph; for each case the
                            constructor kwargs, state_dict, inputs, outputs and ALL gradients.
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

import torch  # noqa: E402

import dgl  # noqa: E402  (the shim)
from physicsnemo.models.gnn_layers.mesh_graph_decoder import MeshGraphDecoder  # noqa: E402
from physicsnemo.models.gnn_layers.mesh_graph_encoder import MeshGraphEncoder  # noqa: E402
import importlib.util  # noqa: E402

# load the processor module by file: the package __init__ pulls in GraphCastNet and its data utilities
_spec = importlib.util.spec_from_file_location(
    "_ref_graph_cast_processor", "/root/reference/physicsnemo/models/graphcast/graph_cast_processor.py")
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
GraphCastProcessor = _mod.GraphCastProcessor

from modulus_b200.mesh import random_graph_csc  # noqa: E402

torch.set_num_threads(8)


def coo(offsets, indices, n_dst):
    deg = offsets[1:] - offsets[:-1]
    return indices.clone(), torch.repeat_interleave(torch.arange(n_dst), deg)


def grads_of(model, inputs, outs):
    loss = sum((o * torch.linspace(-1, 1, o.numel()).view_as(o)).sum() for o in outs)
    loss.backward()
    return ({k: p.grad.clone() for k, p in model.named_parameters()}, [x.grad.clone() for x in inputs])


def main():
    out = {}
    n_src, n_dst = 61, 43
    offsets, indices = random_graph_csc(n_src, n_dst, 1, 6, seed=42)
    src, dst = coo(offsets, indices, n_dst)
    E = src.numel()
    for trick in (False, True):
        for agg in ("sum", "mean"):
            tag = f"{'trick' if trick else 'concat'}_{agg}"
            # ---- encoder: grid (src) -> mesh (dst)
            torch.manual_seed(31 + trick)
            kw = dict(aggregation=agg, input_dim_src_nodes=12, input_dim_dst_nodes=20, input_dim_edges=8,
                      output_dim_src_nodes=12, output_dim_dst_nodes=20, output_dim_edges=24, hidden_dim=32,
                      hidden_layers=1, do_concat_trick=trick)
            model = MeshGraphEncoder(**kw)
            g = dgl.DGLGraph(src, dst, num_src=n_src, num_dst=n_dst, bipartite=True)
            ef = torch.randn(E, 8, requires_grad=True)
            grid = torch.randn(n_src, 12, requires_grad=True)
            mesh = torch.randn(n_dst, 20, requires_grad=True)
            grid_o, mesh_o = model(ef, grid, mesh, g)
            pg, ig = grads_of(model, [ef, grid, mesh], [grid_o, mesh_o])
            out[f"encoder_{tag}"] = dict(kwargs=kw, state_dict={k: v.clone() for k, v in model.state_dict().items()},
                                         inputs=[ef.detach(), grid.detach(), mesh.detach()],
                                         outputs=[grid_o.detach(), mesh_o.detach()], param_grads=pg, input_grads=ig)
            # ---- decoder: mesh (src) -> grid (dst)
            torch.manual_seed(41 + trick)
            kw = dict(aggregation=agg, input_dim_src_nodes=20, input_dim_dst_nodes=12, input_dim_edges=8,
                      output_dim_dst_nodes=12, output_dim_edges=16, hidden_dim=32, hidden_layers=1,
                      do_concat_trick=trick)
            model = MeshGraphDecoder(**kw)
            g = dgl.DGLGraph(src, dst, num_src=n_src, num_dst=n_dst, bipartite=True)
            ef = torch.randn(E, 8, requires_grad=True)
            grid = torch.randn(n_dst, 12, requires_grad=True)   # destination side
            mesh = torch.randn(n_src, 20, requires_grad=True)   # source side
            grid_o = model(ef, grid, mesh, g)
            pg, ig = grads_of(model, [ef, grid, mesh], [grid_o])
            out[f"decoder_{tag}"] = dict(kwargs=kw, state_dict={k: v.clone() for k, v in model.state_dict().items()},
                                         inputs=[ef.detach(), grid.detach(), mesh.detach()],
                                         outputs=[grid_o.detach()], param_grads=pg, input_grads=ig)
    out["bipartite_graph"] = dict(offsets=offsets, indices=indices, n_src=n_src, n_dst=n_dst)

    # ---- processor on a square graph
    n = 50
    off_s, idx_s = random_graph_csc(n, n, 1, 5, seed=7)
    src, dst = coo(off_s, idx_s, n)
    for trick in (False, True):
        torch.manual_seed(51 + trick)
        kw = dict(aggregation="sum", processor_layers=3, input_dim_nodes=32, input_dim_edges=32, hidden_dim=32,
                  hidden_layers=1, do_concat_trick=trick)
        model = GraphCastProcessor(**kw)
        g = dgl.graph((src, dst), num_nodes=n)
        ef = torch.randn(src.numel(), 32, requires_grad=True)
        nf = torch.randn(n, 32, requires_grad=True)
        ef_o, nf_o = model(ef, nf, g)
        pg, ig = grads_of(model, [ef, nf], [ef_o, nf_o])
        out[f"processor_{'trick' if trick else 'concat'}"] = dict(
            kwargs=kw, state_dict={k: v.clone() for k, v in model.state_dict().items()},
            inputs=[ef.detach(), nf.detach()], outputs=[ef_o.detach(), nf_o.detach()], param_grads=pg, input_grads=ig)
    out["square_graph"] = dict(offsets=off_s, indices=idx_s, n=n)
    path = os.path.join(HERE, "ref_graphcast_blocks.pt")
    torch.save(out, path)
    print(f"wrote ref_graphcast_blocks.pt: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} entries")


if __name__ == "__main__":
    main()
