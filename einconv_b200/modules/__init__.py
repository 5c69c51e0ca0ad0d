"""Module interface: ``ConvNd`` and ``UnfoldNd``."""

from einconv_b200.modules.conv import ConvNd
from einconv_b200.modules.unfold import UnfoldNd

__all__ = ["ConvNd", "UnfoldNd"]
