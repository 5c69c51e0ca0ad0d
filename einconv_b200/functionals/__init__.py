"""Functional interface: ``convNd`` and ``unfoldNd``."""

from einconv_b200.functionals.conv import convNd
from einconv_b200.functionals.unfold import unfoldNd

__all__ = ["convNd", "unfoldNd"]
