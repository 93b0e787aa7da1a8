"""Stub of `s3fs` (`physicsnemo/utils/filesystem.py:28`).  Test infrastructure only."""


class S3FileSystem:  # pragma: no cover
    def __init__(self, *a, **k):
        pass
