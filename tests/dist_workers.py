"""Worker functions of the multi-rank tests (spawned with torch.multiprocessing; one process per rank).

They replay the reference's own distributed tests:
  collectives            test/distributed/test_autograd.py:29-208   (scatter_v, gather_v, all_gather_v, indexed_all_to_all_v:
                                                                     values and gradients)
  DistributedGraph       test/models/test_distributed_graph.py:186-336 (all scatter_features x get_on_all_ranks combinations
                                                                     for src / dst / edge features, halo exchange against a
                                                                     global scatter_reduce, nodewise and lat-lon-bbox partitions)
  model                  test/models/meshgraphnet/test_meshgraphnet_snmg.py:56-248 (partitioned == single device: output and
                                                                     every weight gradient)

Transport: NCCL when every rank has its own GPU; otherwise gloo -- on host tensors for the CPU tier, and with several ranks
SHARING cuda:0 for the single-GPU tier (modulus_b200.distributed.utils stages device buffers through the host for any
non-NCCL backend; every kernel, index map and protocol step of the partitioned product path is the one NCCL runs with).
"""
import os
import socket
import traceback

import torch


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port, use_cuda):
    from modulus_b200.distributed import DistributedManager

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if use_cuda:
        ndev = torch.cuda.device_count()
        local = rank % ndev
        backend = "nccl" if ndev >= world else "gloo"
        os.environ["LOCAL_RANK"] = str(local)
        DistributedManager.setup(rank, world, local_rank=local, port=str(port), backend=backend)
    else:
        os.environ["LOCAL_RANK"] = "0"
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        DistributedManager.setup(rank, world, local_rank=0, port=str(port), backend="gloo")
    dm = DistributedManager()
    assert dm.is_initialized() and dm.distributed
    return dm


def run(rank, world, port, use_cuda, what, result_dir, kwargs):
    """entry point of every spawned rank: runs `what`, stores its result (or the traceback) for the parent"""
    from modulus_b200.distributed import DistributedManager

    out = {"ok": False}
    try:
        dm = _setup(rank, world, port, use_cuda)
        out["result"] = globals()[what](dm, **kwargs)
        out["ok"] = True
    except Exception:
        out["error"] = traceback.format_exc()
    torch.save(out, os.path.join(result_dir, f"r{rank}.pt"))
    try:
        DistributedManager.cleanup()
    except Exception:
        pass


def launch(world, use_cuda, what, tmp_path, **kwargs):
    """spawn `world` ranks, return their results; raises with the first failing rank's traceback"""
    import torch.multiprocessing as mp

    mp.spawn(run, args=(world, free_port(), use_cuda, what, str(tmp_path), kwargs), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt"), weights_only=False) for r in range(world)]
    for r, o in enumerate(res):
        assert o["ok"], f"rank {r}:\n{o.get('error')}"
    return [o["result"] for o in res]


# ------------------------------------------------------------------------------------------------------------------
# collectives (test_autograd.py)
# ------------------------------------------------------------------------------------------------------------------
def collectives(dm):
    from modulus_b200.distributed.autograd import all_gather_v, gather_v, indexed_all_to_all_v, scatter_v

    rank, world, dev = dm.rank, dm.world_size, dm.device
    dim = 4
    sizes = [r + 2 for r in range(world)]
    blocks = (torch.arange(world, device=dev, dtype=torch.float32) + 1).view(-1, 1).expand(-1, dim).contiguous()
    stacked = blocks.repeat_interleave(torch.tensor(sizes, device=dev), dim=0)

    # scatter_v: rank 0 holds the stacked tensor, rank r receives block r; gradient returns to rank 0
    t = stacked.clone().requires_grad_(True)
    got = scatter_v(t, sizes, dim=0, src=0, group=None)
    assert torch.allclose(got, torch.full((sizes[rank], dim), float(rank + 1), device=dev))
    got.backward(gradient=-torch.ones_like(got))
    if rank == 0:
        assert torch.allclose(t.grad, -torch.ones_like(t))

    # gather_v: the reverse
    t = torch.full((rank + 2, dim), float(rank + 1), device=dev, requires_grad=True)
    got = gather_v(t, sizes, dim=0, dst=0, group=None)
    if rank == 0:
        assert torch.allclose(got, stacked)
    got.backward(gradient=-torch.ones_like(got))
    assert torch.allclose(t.grad, -torch.ones_like(t))

    # all_gather_v: everybody gets the stacked tensor, gradients are summed over ranks
    t = torch.full((rank + 2, dim), float(rank + 1), device=dev, requires_grad=True)
    got = all_gather_v(t, sizes, dim=0, group=None)
    assert torch.allclose(got, stacked)
    got.backward(gradient=-torch.ones_like(got))
    assert torch.allclose(t.grad, -torch.ones_like(t) * world)

    # indexed_all_to_all_v (test_autograd.py:157-208): rank p holds (p+1) rows of every value 1..world and sends the rows
    # of value r+1 to rank r
    t = (torch.arange(1, world + 1, device=dev, dtype=torch.float32).view(-1, 1).expand(-1, dim).contiguous()
         .repeat_interleave(repeats=rank + 1, dim=0)).requires_grad_(True)
    szs = [[r + 1 for _ in range(world)] for r in range(world)]
    idx = [torch.nonzero(t[:, 0] == (r + 1)).view(-1) for r in range(world)]
    got = indexed_all_to_all_v(t, idx, szs, dim=0, use_fp32=True, group=None)
    n_expected = sum(szs[r][rank] for r in range(world))
    assert got.shape == (n_expected, dim) and torch.allclose(got, torch.full_like(got, float(rank + 1)))
    got.backward(gradient=-torch.ones_like(got))
    assert torch.allclose(t.grad, -torch.ones_like(t))

    # a genuinely indexed case: duplicates and a permutation, checked against the single-process definition
    g = torch.Generator().manual_seed(100 + rank)
    rows = 7 + rank
    x_all = [torch.randn(7 + r, dim, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    idx_all = [[torch.randint(0, 7 + p, (3 + (p + 2 * r) % 4,), generator=torch.Generator().manual_seed(17 * p + r))
                for r in range(world)] for p in range(world)]
    szs = [[int(idx_all[p][r].numel()) for r in range(world)] for p in range(world)]
    x = x_all[rank].to(dev).requires_grad_(True)
    got = indexed_all_to_all_v(x, [i.to(dev) for i in idx_all[rank]], szs, dim=0, use_fp32=True, group=None)
    want = torch.cat([x_all[p][idx_all[p][rank]] for p in range(world)])
    assert torch.allclose(got.cpu(), want)
    w_all = [torch.randn(sum(szs[p][r] for p in range(world)), dim, generator=torch.Generator().manual_seed(900 + r))
             for r in range(world)]
    got.backward(gradient=w_all[rank].to(dev))
    g_want = torch.zeros(rows, dim)
    for r in range(world):  # rank r received my rows idx_all[rank][r] at offset sum_{p<rank} szs[p][r]
        off = sum(szs[p][r] for p in range(rank))
        g_want.index_add_(0, idx_all[rank][r], w_all[r][off:off + szs[rank][r]])
    assert torch.allclose(x.grad.cpu(), g_want, atol=1e-6)
    return True


# ------------------------------------------------------------------------------------------------------------------
# DistributedGraph (test_distributed_graph.py)
# ------------------------------------------------------------------------------------------------------------------
def _scatter_reduce(feat, offsets, indices):
    n_dst = offsets.numel() - 1
    dst = torch.repeat_interleave(torch.arange(n_dst, device=feat.device), offsets[1:] - offsets[:-1])
    out = torch.zeros((n_dst,) + tuple(feat.shape[1:]), dtype=feat.dtype, device=feat.device)
    return out.index_add(0, dst, feat[indices.long()])


def distributed_graph(dm, partition_scheme):
    from modulus_b200.mesh import random_graph_csc
    from modulus_b200.models.gnn_layers import DistributedGraph, partition_graph_by_coordinate_bbox
    from modulus_b200.models.graphcast import get_lat_lon_partition_separators

    dev, world = dm.device, dm.world_size
    n_src, n_dst, C = 4321, 1234, 64
    offsets, indices = random_graph_csc(n_src, n_dst, 2, 8, seed=42, device=dev)
    g = torch.Generator().manual_seed(42)
    src_feat = (10 * torch.rand((n_src, C), generator=g) + 16).to(dev).requires_grad_(True)
    dst_feat = (10 * torch.rand((n_dst, C), generator=g) + 8).to(dev).requires_grad_(True)
    edge_feat = (10 * torch.rand((indices.numel(), C), generator=g) + 4).to(dev).requires_grad_(True)
    gp = None
    if partition_scheme == "lat_lon_bbox":
        x_src, x_dst = torch.rand((n_src, 2), generator=g), torch.rand((n_dst, 2), generator=g)
        for x in (x_src, x_dst):
            x[:, 0] = x[:, 0] * 180 - 90
            x[:, 1] = x[:, 1] * 360 - 180
        lo, hi = get_lat_lon_partition_separators(world)
        gp = partition_graph_by_coordinate_bbox(offsets, indices, x_src.to(dev), x_dst.to(dev), lo, hi, world, dm.rank, dev)
    dg = DistributedGraph(offsets, indices, partition_size=world, graph_partition=gp)
    P, rank = dg.partition_size, dg.partition_rank

    def roundtrip(feat, to_partition, to_global):
        for scatter_features in (False, True):
            for get_on_all_ranks in (False, True):
                feat.grad = None
                local = to_partition(feat, scatter_features=scatter_features)
                glob = to_global(local, get_on_all_ranks=get_on_all_ranks)
                loss = glob.sum() / (P if get_on_all_ranks else 1)
                loss.backward()
                if get_on_all_ranks or rank == 0:
                    assert torch.allclose(glob, feat)
                if scatter_features:
                    if rank == 0:
                        assert torch.allclose(feat.grad, torch.ones_like(feat))
                else:
                    with torch.no_grad():
                        assert torch.allclose(to_partition(feat.grad, scatter_features=False), torch.ones_like(local))

    roundtrip(src_feat, dg.get_src_node_features_in_partition, dg.get_global_src_node_features)
    roundtrip(dst_feat, dg.get_dst_node_features_in_partition, dg.get_global_dst_node_features)
    roundtrip(edge_feat, dg.get_edge_features_in_partition, dg.get_global_edge_features)

    # halo exchange + local aggregation == aggregation on the global graph, values and source gradients
    for scatter_features in (False, True):
        for get_on_all_ranks in (False, True):
            ref_src = src_feat.detach().clone().requires_grad_(True)
            src_feat.grad = None
            local = dg.get_src_node_features_in_partition(src_feat, scatter_features=scatter_features)
            global_agg = _scatter_reduce(ref_src, offsets, indices)
            local = dg.get_src_node_features_in_local_graph(local)
            local_agg = _scatter_reduce(local, dg.graph_partition.local_offsets, dg.graph_partition.local_indices)
            local_agg = dg.get_global_dst_node_features(local_agg, get_on_all_ranks=get_on_all_ranks)
            if get_on_all_ranks or rank == 0:
                assert torch.allclose(local_agg, global_agg)
            (local_agg.sum() / (P if get_on_all_ranks else 1)).backward()
            global_agg.sum().backward()
            if scatter_features:
                if rank == 0:
                    assert torch.allclose(src_feat.grad, ref_src.grad)
            else:
                with torch.no_grad():
                    a = dg.get_src_node_features_in_partition(src_feat.grad, scatter_features=False)
                    b = dg.get_src_node_features_in_partition(ref_src.grad, scatter_features=False)
                    assert torch.allclose(a, b)
    halo = dg.graph_partition.num_local_src_nodes - dg.graph_partition.sizes[rank][rank]
    return {"halo_rows": int(halo), "local_src": int(dg.graph_partition.num_local_src_nodes)}


# ------------------------------------------------------------------------------------------------------------------
# partitioned model == single-device model (test_meshgraphnet_snmg.py)
# ------------------------------------------------------------------------------------------------------------------
def model_parity(dm, cases):
    """cases: list of dicts(name, use_bf16, shuffle, layers, partition).  Returns per case the relative deviations."""
    from modulus_b200 import fused
    from modulus_b200.distributed import mark_module_as_shared, unmark_module_as_shared
    from modulus_b200.mesh import _csc_from_pairs, triangle_grid_mesh
    from modulus_b200.models.gnn_layers import CuGraphCSC, partition_graph_by_coordinate_bbox
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    dev, world, rank = dm.device, dm.world_size, dm.rank
    if "graph_partition" not in dm.group_names:
        dm.create_process_subgroup("graph_partition", world)
    results = {}
    for case in cases:
        mesh = triangle_grid_mesh(48, 37)
        n = mesh["num_nodes"]
        offsets, indices, coords = mesh["offsets"], mesh["indices"], mesh["coords"]
        if case.get("shuffle"):  # random node numbering: almost every source row of a rank is a halo row
            perm = torch.randperm(n, generator=torch.Generator().manual_seed(5))
            deg = offsets[1:] - offsets[:-1]
            dst = torch.repeat_interleave(torch.arange(n), deg)
            offsets, indices = _csc_from_pairs(perm[indices], perm[dst], n, n)
            new_coords = torch.empty_like(coords)
            new_coords[perm] = coords
            coords = new_coords
        deg = offsets[1:] - offsets[:-1]
        dst = torch.repeat_interleave(torch.arange(n), deg)
        disp = coords[indices] - coords[dst]
        ef = torch.cat([disp, disp.norm(dim=1, keepdim=True)], dim=1)
        g = torch.Generator().manual_seed(3)
        nf, tgt = torch.randn(n, 6, generator=g), torch.randn(n, 3, generator=g)
        torch.manual_seed(11)
        model = MeshGraphNet(6, 3, 3, processor_size=case.get("layers", 3)).to(dev)
        use_bf16 = case["use_bf16"]

        def step(m, graph, nf_l, ef_l, tgt_l):
            m.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_bf16):
                out = m(nf_l.to(dev), ef_l.to(dev), graph)
            loss = ((out.float() - tgt_l.to(dev)) ** 2).sum() / n  # sum-reduced: per-rank losses add up to the global one
            loss.backward()
            return out.detach().float(), {k: v.grad.detach().clone() for k, v in m.named_parameters()}

        g_single = CuGraphCSC(offsets.to(dev), indices.to(dev), n, n)
        out_s, grads_s = step(model, g_single, nf, ef, tgt)
        # "lean": the memory-lean mode of the fused path (no stored h1, projections recomputed, full-table exchange with
        # the exchanged rows kept for the backward pass) on the partitioned side
        fused.KEEP_H1 = not case.get("lean", False)
        gp = None
        if case.get("partition") == "bbox":  # vertical strips of the unit square by x coordinate
            cuts = [i / world for i in range(world + 1)]
            lo = [[cuts[i] if i > 0 else None, None] for i in range(world)]
            hi = [[cuts[i + 1] if i + 1 < world else None, None] for i in range(world)]
            gp = partition_graph_by_coordinate_bbox(offsets.to(dev), indices.to(dev), coords.to(dev), coords.to(dev), lo, hi,
                                                    world, rank, dev)
        g_dist = CuGraphCSC(offsets.to(dev), indices.to(dev), n, n, partition_size=world,
                            partition_group_name="graph_partition", graph_partition=gp)
        mark_module_as_shared(model, "graph_partition")
        nf_l = g_dist.get_src_node_features_in_partition(nf.to(dev))
        ef_l = g_dist.get_edge_features_in_partition(ef.to(dev))
        tgt_l = g_dist.get_dst_node_features_in_partition(tgt.to(dev))
        out_l, grads_d = step(model, g_dist, nf_l, ef_l, tgt_l)
        out_d = g_dist.get_global_dst_node_features(out_l)
        fused.KEEP_H1 = True
        unmark_module_as_shared(model)
        torch.cuda.synchronize()

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

        res = {"out": rel(out_d, out_s), "grads": {k: rel(grads_d[k], grads_s[k]) for k in grads_s}, "fused": None}
        h = g_dist.b200_plan().extra.get("halo")
        if h is not None:
            res["fused"] = dict(e0=h.e0, e1=h.e1, n_edges=g_dist.b200_plan().n_edges, halo_rows=h.halo_rows,
                                n_part=h.n_part, remote_only=h.remote_only, peer=getattr(h, "peer", None) is not None)
        results[case["name"]] = res
    return results
