"""Stub of `timm` (`physicsnemo/models/layers/transformer_layers.py:20-21`)."""
