class SwinTransformerStage:  # pragma: no cover
    pass
