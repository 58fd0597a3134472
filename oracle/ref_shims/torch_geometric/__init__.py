"""Test-scaffolding stand-in for `torch_geometric` (absent from this image).
ORACLE INFRASTRUCTURE ONLY -- never imported by the product path."""
__version__ = "2.3.0"
from . import data, nn, utils  # noqa: F401
