"""Import stub so the reference's datapipes/gnn/vortex_shedding_dataset.py can be imported for its pure static
methods (cell_to_adj, create_graph, add_edge_features, normalize_edge).  No TFRecord reading."""
