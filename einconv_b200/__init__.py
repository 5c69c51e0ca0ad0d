"""einconv_b200: einconv's convolution tensor networks evaluated by sm_100a kernels.

Same public surface as the reference package (``einconv/__init__.py:3-9``): ``index_pattern`` and
``simplify`` at the top level, ``functionals``, ``modules`` and ``expressions`` below it.
"""

from einconv_b200.conv_index_pattern import IndexPattern, index_pattern, index_pattern_logical
from einconv_b200.simplifications import simplify

__all__ = ["index_pattern", "simplify", "IndexPattern", "index_pattern_logical"]


def __getattr__(name):
    # evaluator / functionals / modules import lazily so that `import einconv_b200` stays cheap
    if name in ("evaluator", "functionals", "modules", "expressions", "distributed"):
        import importlib

        return importlib.import_module(f"einconv_b200.{name}")
    raise AttributeError(name)
