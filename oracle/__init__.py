"""CPU oracle for the torch-motion-correction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product package
``torch_motion_correction_b200`` never imports this package and has no CPU
fallback.

Contents
--------
``deps``            restatement of the published algorithms of the five un-vendored,
                    un-pinned third-party packages the reference imports
                    (torch-cubic-spline-grids, torch-image-interpolation,
                    torch-fourier-shift, torch-fourier-filter, torch-grid-utils;
                    ``pyproject.toml:36-46`` of the reference lists them as bare
                    names, no lock file exists).
``reference_path``  restatement (torch-CPU / numpy) of the reference's own functions on
                    the hot path, each citing the reference ``file:line`` it follows.
``verbatim``        loader that runs the UNMODIFIED reference source from
                    ``/root/reference/src`` on top of ``deps`` (only possible in the
                    build container; used by ``tests/golden/make_golden.py`` to pin
                    ``reference_path`` and to generate the committed golden vectors).

Parity status
-------------
The reference ships no golden vectors and asserts no numeric values beyond four
zero-field identity checks (reference ``tests/test_correct_motion.py:132-145,188-199,
241-252,482-499``).  ``reference_path`` is therefore pinned against outputs of the
reference's own source run here (``tests/golden/*.npz`` + generator script committed).
The five third-party packages could not be installed (no network), so their semantics
are restated from their published algorithms: **parity is pinned for the reference's
own code, and unpinned for the third-party dependency layer** ("parity unpinned" for
that layer).
"""
