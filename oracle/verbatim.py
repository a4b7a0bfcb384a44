"""Run the UNMODIFIED reference source on top of the restated dependency layer.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only where
``/root/reference`` exists (the build container): it is used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the
``-m "not gpu"`` tests (when the tree is present) to re-validate ``oracle.reference_path``.
Nothing here is copied from the reference; the reference package is imported from where
it lies.
"""

from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

REFERENCE_SRC = os.environ.get("TMC_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "torch_motion_correction"))


def load():
    """Import and return the reference package ``torch_motion_correction`` (verbatim)."""
    if not available():
        raise RuntimeError(f"reference source not present at {REFERENCE_SRC}")
    from oracle import deps

    deps.install_stand_ins()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    mod = sys.modules.get("torch_motion_correction")
    if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_SRC):
        # a different package (e.g. the repo's drop-in shim) owns the name: evict it
        for name in [n for n in sys.modules if n == "torch_motion_correction" or n.startswith("torch_motion_correction.")]:
            del sys.modules[name]
    return importlib.import_module("torch_motion_correction")


@contextlib.contextmanager
def quiet():
    """Silence the reference's print()/tqdm chatter."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield
