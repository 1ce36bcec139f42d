"""Mesh / camera ingestion: the on-disk formats of the reference's callers."""
from .OBJReader import OBJReader  # noqa: F401
from .CameraReader import CameraReader  # noqa: F401
