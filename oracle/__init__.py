"""ORACLE — test infrastructure only.

CPU restatements of the reference's hot path (f-dangel/einconv):

* ``oracle.direct``      plain-C, double-precision evaluation of the six networks from their
                         definitions (``einconv_oracle.c``), called through ctypes/numpy;
* ``oracle.einsum_port`` the reference's own algorithm restated: dense one-hot index patterns +
                         ``torch.einsum`` on CPU (what ``einconv/functionals/conv.py:61`` runs).
                         This is what ``bench.py`` times as the CPU baseline.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  Nothing under ``einconv_b200/`` does.  Both restatements are pinned
against golden vectors produced by the reference itself (``tests/golden/``).
"""
