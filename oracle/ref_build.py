"""ORACLE (test infrastructure): build recipe for ``oracle/_ref`` — the UNMODIFIED reference.

The reference (f-dangel/einconv) is pure Python; ``pip install /root/reference`` fails offline
because its ``setup.cfg`` asks for ``setuptools_scm`` (absent from the wheelhouse).
Its "build" is therefore the byte-compilation of the package where it lies under
``/root/reference``: ``compileall``-style ``.pyc`` files, outputs ONLY into ``oracle/_ref/einconv``
(git-ignored, not gpurun-ignored, so it travels to the GPU box like the built ``.so`` files).  No
source file of the reference is copied or stored.

Users: ``tests/`` (cross-checks of the oracle port), ``bench.py``'s ``cpu_baseline`` leg and
``bench.py --impl reference``.  Nothing under ``einconv_b200/`` may import this.
"""

import importlib
import py_compile
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/einconv")
REF_OUT = HERE / "_ref"


def build(force: bool = False) -> bool:
    """Byte-compile the reference package into ``oracle/_ref/einconv``.  Returns True when the
    compiled reference is available afterwards (already built, or built now)."""
    marker = REF_OUT / "einconv" / "__init__.pyc"
    if marker.exists() and not force:
        return True
    if not REF_SRC.is_dir():
        return marker.exists()
    for src in sorted(REF_SRC.rglob("*.py")):
        dst = REF_OUT / "einconv" / src.relative_to(REF_SRC).with_suffix(".pyc")
        dst.parent.mkdir(parents=True, exist_ok=True)
        # dfile: the path recorded in tracebacks; unchecked hash -> no source needed at import time
        py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/einconv/{src.relative_to(REF_SRC)}",
                           doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return marker.exists()


def available() -> bool:
    return (REF_OUT / "einconv" / "__init__.pyc").exists()


def load():
    """Import and return the compiled reference package (``einconv``), or None if it was not built."""
    if not available():
        return None
    if str(REF_OUT) not in sys.path:
        sys.path.insert(0, str(REF_OUT))
    return importlib.import_module("einconv")


if __name__ == "__main__":
    print("reference built:", build(force="--force" in sys.argv))
