"""Functional interface: ``convNd``, ``unfoldNd``, ``unfoldNd_transpose`` (plus ``unfoldNd_host``)."""

from einconv_b200.functionals.conv import convNd
from einconv_b200.functionals.unfold import unfoldNd, unfoldNd_host
from einconv_b200.functionals.unfold_transpose import unfoldNd_transpose

__all__ = ["convNd", "unfoldNd", "unfoldNd_transpose", "unfoldNd_host"]
