"""Import stub (physicsnemo/datapipes/gnn/utils.py:23-26 imports vtk at module level for its file readers)."""
