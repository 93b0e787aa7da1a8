"""Multi-rank host logic on the CPU tier: two gloo ranks, host tensors (no GPU).  Replays
test/distributed/test_autograd.py:29-208 and test/models/test_distributed_graph.py:186-336 of the reference."""
import pytest

import dist_workers as W


def test_collectives_values_and_gradients_gloo_world2(tmp_path):
    assert W.launch(2, False, "collectives", tmp_path) == [True, True]


@pytest.mark.parametrize("scheme", ["nodewise", "lat_lon_bbox"])
def test_distributed_graph_all_combinations_gloo_world2(tmp_path, scheme):
    res = W.launch(2, False, "distributed_graph", tmp_path, partition_scheme=scheme)
    assert all(r["halo_rows"] > 0 for r in res)  # the exchange really moved remote rows
