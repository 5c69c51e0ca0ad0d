"""ORACLE (test infrastructure): build recipe for ``oracle/_ref`` — the UNMODIFIED reference.

The reference (f-dangel/einconv) is pure Python; ``pip install /root/reference`` fails offline
because its ``setup.cfg`` asks for ``setuptools_scm`` (absent from the wheelhouse).
Its "build" is therefore the byte-compilation of the package where it lies under
``/root/reference``: sourceless ``.pyc`` files packed into ONE archive, ``oracle/_ref/einconv.zip``
(git-ignored, not gpurun-ignored, so it travels to the GPU box like the built ``.so`` files; loose
``*.pyc`` files are dropped by the snapshot).  Python imports it through ``zipimport``.  No source file
of the reference is copied or stored.

Users: ``tests/`` (cross-checks of the oracle port), ``bench.py``'s ``cpu_baseline`` leg and
``bench.py --impl reference``.  Nothing under ``einconv_b200/`` may import this.
"""

import importlib
import py_compile
import sys
import tempfile
import zipfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/einconv")
REF_OUT = HERE / "_ref"


ARCHIVE = REF_OUT / "einconv.zip"


def build(force: bool = False) -> bool:
    """Byte-compile the reference package into ``oracle/_ref/einconv.zip``.  Returns True when the compiled
    reference is available afterwards (already built, or built now)."""
    if ARCHIVE.exists() and not force:
        return True
    if not REF_SRC.is_dir():
        return ARCHIVE.exists()
    REF_OUT.mkdir(parents=True, exist_ok=True)
    tmp_archive = ARCHIVE.with_suffix(".zip.tmp")
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(tmp_archive, "w", zipfile.ZIP_DEFLATED) as zf:
        for src in sorted(REF_SRC.rglob("*.py")):
            rel = src.relative_to(REF_SRC).with_suffix(".pyc")
            dst = Path(tmp) / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            # dfile: the path recorded in tracebacks; unchecked hash -> no source needed at import time
            py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/einconv/{src.relative_to(REF_SRC)}",
                               doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            zf.write(dst, str(Path("einconv") / rel))
    tmp_archive.replace(ARCHIVE)
    return True


def available() -> bool:
    return ARCHIVE.exists()


def load():
    """Import and return the compiled reference package (``einconv``), or None if it was not built.
    The archive is tied to the interpreter that compiled it (same image here and on the GPU box)."""
    if not available():
        return None
    if str(ARCHIVE) not in sys.path:
        sys.path.insert(0, str(ARCHIVE))
    try:
        return importlib.import_module("einconv")
    except ImportError:  # e.g. bytecode of another Python version
        return None


if __name__ == "__main__":
    print("reference built:", build(force="--force" in sys.argv))
