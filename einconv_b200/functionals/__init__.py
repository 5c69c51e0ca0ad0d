"""Functional interface: ``convNd`` and ``unfoldNd`` (plus ``unfoldNd_host`` for host buffers)."""

from einconv_b200.functionals.conv import convNd
from einconv_b200.functionals.unfold import unfoldNd, unfoldNd_host

__all__ = ["convNd", "unfoldNd", "unfoldNd_host"]
