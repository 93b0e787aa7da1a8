"""Latitude / longitude box partition of GraphCast's graphs.

`get_lat_lon_partition_separators(P)` returns the `(min, max)` corner lists consumed by
`partition_graph_by_coordinate_bbox`, with the boxes the reference draws
(physicsnemo/models/graphcast/graph_cast_net.py:52-113): the sphere's [-90, 90) x [-180, 180) rectangle cut
into n_lat x n_lon equal boxes, n_lat * n_lon = P with the two factors as close as possible (n_lat <= n_lon),
box index = i_lat * n_lon + i_lon, and only the very last box open on its upper sides.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def _closest_factors(p: int) -> Tuple[int, int]:
    n_lat = max(d for d in range(1, int(p ** 0.5) + 1) if p % d == 0)
    # the reference's search stops at the first divisor pair with lat >= lon, scanning upwards: for a perfect square
    # that is (sqrt, sqrt); otherwise the first divisor ABOVE sqrt(p) goes to latitude
    if n_lat * n_lat == p:
        return n_lat, n_lat
    return p // n_lat, n_lat


def get_lat_lon_partition_separators(partition_size: int) -> Tuple[List[List[Optional[float]]], List[List[Optional[float]]]]:
    if partition_size < 1:
        raise ValueError(f"partition_size must be positive, got {partition_size}")
    n_lat, n_lon = _closest_factors(partition_size)
    d_lat, d_lon = 180.0 / n_lat, 360.0 / n_lon
    lo = [[d_lat * (b // n_lon) - 90.0, d_lon * (b % n_lon) - 180.0] for b in range(partition_size)]
    hi = [[d_lat * (b // n_lon + 1) - 90.0, d_lon * (b % n_lon + 1) - 180.0] for b in range(partition_size)]
    hi[-1] = [None, None]
    return lo, hi
