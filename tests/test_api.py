"""Host-side contract of the drop-in: library symbols, constructor/state_dict parity, init RNG
stream, error behaviour.  No GPU needed."""
import ctypes
import re

import pytest
import torch

from conftest import load_golden
from modulus_b200 import _lib
from modulus_b200.models.gnn_layers import (CuGraphCSC, MeshEdgeBlock, MeshGraphMLP, MeshNodeBlock,
                                            aggregate_and_concat)
from modulus_b200.models.meshgraphnet import MeshGraphNet


def test_library_loads_and_exports_every_declared_symbol():
    header = _lib.HEADER.read_text()
    declared = set(re.findall(r"\b(mgn_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    assert len(declared) >= 15
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()  # loading must work without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mgn_version() >= 100
    assert lib.mgn_error_string(-1) == b"invalid argument"


def test_null_arguments_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    assert lib.mgn_segment_sum(0, None, 4, 0, 4, None, None, 3, None, 4, 0, 0, 0, None) == _lib.MGN_EINVAL
    assert lib.mgn_linear_fwd(7, None, 0, 0, 0, None, None, 0, 0, None, None, 0, None) == _lib.MGN_OK  # M == 0
    with pytest.raises(ValueError):
        _lib.check(_lib.MGN_EINVAL, "x")
    with pytest.raises(_lib.MGNError):
        _lib.check(_lib.MGN_EUNSUPPORTED, "x")


@pytest.mark.parametrize("case", ["relu_sum", "concat_trick"])
def test_state_dict_layout_and_init_match_reference(case):
    """Same seed -> same parameter names, shapes AND values as the reference constructor."""
    g = load_golden(f"ref_mgn_{case}.pt")
    seed = {"relu_sum": 11, "concat_trick": 14}[case]
    torch.manual_seed(seed)
    model = MeshGraphNet(**g["kwargs"])
    sd, ref = model.state_dict(), g["state_dict"]
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape, k
        assert torch.equal(sd[k], ref[k]), k
    model.load_state_dict(ref)  # interchangeable checkpoints


def test_default_model_has_reference_parameter_count():
    m = MeshGraphNet(4, 3, 2)
    assert sum(p.numel() for p in m.parameters()) == 2332034
    assert len(m.state_dict()) == 263


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError):
        MeshGraphNet(4, 3, 2, norm_type="BatchNorm")
    with pytest.raises(ValueError):
        MeshGraphMLP(4, 8, 8, 2, norm_type="Foo")
    with pytest.raises(ValueError):  # recompute_activation needs SiLU (mesh_graph_mlp.py:172-177)
        MeshGraphMLP(4, 8, 8, 2, activation_fn=torch.nn.ReLU(), recompute_activation=True)
    with pytest.raises(KeyError):
        MeshGraphNet(4, 3, 2, mlp_activation_fn="nope")
    with pytest.raises(ValueError):
        MeshGraphNet(4, 3, 2, processor_size=3, num_processor_checkpoint_segments=4)
    assert isinstance(MeshGraphMLP(4, hidden_layers=None).model, torch.nn.Identity)


def test_cugraphcsc_contract():
    off = torch.tensor([0, 2, 4], dtype=torch.int64)
    idx = torch.tensor([0, 1, 1, 2], dtype=torch.int64)
    g = CuGraphCSC(off, idx, 3, 2)
    assert not g.is_distributed and g.dgl_graph is None
    assert g.get_src_node_features_in_local_graph("x") == "x"
    with pytest.raises(TypeError):
        g.to(dtype=torch.float32)
    g.to(dtype=torch.int32)
    assert g.offsets.dtype == torch.int32
    with pytest.raises(AssertionError):
        CuGraphCSC(off, idx, 3, 2, ef_indices=torch.arange(4), partition_size=2)
    with pytest.raises(RuntimeError):
        g.to_static_csc()


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors."""
    m = MeshGraphNet(4, 3, 2, processor_size=1)
    off = torch.tensor([0, 1, 2], dtype=torch.int64)
    idx = torch.tensor([1, 0], dtype=torch.int64)
    g = CuGraphCSC(off, idx, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4), torch.randn(2, 3), g)
    with pytest.raises(RuntimeError):
        aggregate_and_concat(torch.randn(2, 4), torch.randn(2, 4), g, "max")
    blk = MeshEdgeBlock(8, 8, 8, 8, 2, torch.nn.ReLU())
    assert list(blk.state_dict().keys())[0] == "edge_mlp.model.0.weight"
    nb = MeshNodeBlock("sum", 8, 8, 8, 8, 2, torch.nn.ReLU())
    assert nb.node_mlp.model[0].in_features == 16


def test_product_never_imports_the_oracle():
    import pathlib

    root = pathlib.Path(_lib.ROOT)
    for p in root.rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p


def test_library_carries_the_digest_of_its_sources_and_stale_builds_are_refused(monkeypatch):
    """The built library embeds the sha256 of the sources + flags it was made from; the loader recomputes it from the
    sources next to it and refuses anything else (a stale .so bound to a newer header is undefined behaviour)."""
    from modulus_b200 import build

    lib = _lib.load()
    assert lib.mgn_build_digest().decode() == build._digest() == build.embedded_digest()
    monkeypatch.setattr(build, "_digest", lambda: "0" * 64)
    with pytest.raises(_lib.MGNError, match="rebuild"):
        _lib._verify_digest(lib)


def test_product_library_has_no_process_global_debug_switches():
    """Profiling hooks (include/mgn_b200_debug.h) exist only in -DMGN_DEBUG_HOOKS builds."""
    import os

    if "MGN_DEBUG_HOOKS" in os.environ.get("MGN_NVCC_EXTRA", ""):
        pytest.skip("debug build")
    lib = _lib.load()
    assert _lib.DEBUG_PROTOTYPES and not any(hasattr(lib, n) for n in _lib.DEBUG_PROTOTYPES)
    assert not any("debug" in n for n in _lib.PROTOTYPES)


def test_bench_config_identifies_the_workload_and_is_the_same_for_both_arms():
    """bench.py: `config` carries only what identifies the workload (nothing measured), built by one function for the B200 arm
    and the reference arm; the torus mesh size is known in closed form (E = 6 N) so the reference arm needs no GPU for it."""
    import argparse
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from modulus_b200 import mesh

    m = mesh.torus_surface_mesh(12, 10)
    assert bench.mesh_size("torus_surface_mesh", (12, 10)) == (m["num_nodes"], int(m["indices"].numel())) == (120, 720)
    t = mesh.triangle_grid_mesh(7, 9)
    assert bench.mesh_size("triangle_grid_mesh", (7, 9)) == (t["num_nodes"], int(t["indices"].numel()))
    args = argparse.Namespace(workload="c3", partition="nodewise", stripe_rows=25)
    one = bench.workload_config(args, 1, *bench.mesh_size("torus_surface_mesh", (1000, 1000)))
    assert one["nodes"] == 1_000_000 and one["edges"] == 6_000_000 and one["partition"] == "none"
    assert set(one) == {"workload", "nodes", "edges", "partition", "l2"} and "exceeds the 126 MB L2" in one["l2"]
    eight = bench.workload_config(args, 8, *bench.mesh_size("torus_surface_mesh", (8000, 1000)))
    assert eight["edges"] == 48_000_000 and "x8 ranks" in eight["workload"] and "nodewise" in eight["partition"]
    args.workload = "c1"
    assert "L2-warm" in bench.workload_config(args, 1, *bench.mesh_size("triangle_grid_mesh", (42, 45)))["l2"]
