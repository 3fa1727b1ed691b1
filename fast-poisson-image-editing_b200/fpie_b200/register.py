"""Registers ``b200`` as a ``-b`` backend of an installed ``fpie``.

``fpie`` builds its ``--backend`` choices from the live list
``fpie.process.ALL_BACKEND`` when ``get_args()`` runs (fpie/args.py:25-31) and
``fpie.cli`` / ``fpie.gui`` construct ``EquProcessor`` / ``GridProcessor`` by
name (fpie/cli.py:16-32, fpie/gui.py).  ``register()`` therefore

1. appends ``"b200"`` to ``ALL_BACKEND``;
2. rebinds ``EquProcessor`` / ``GridProcessor`` in ``fpie.process`` (and in
   ``fpie.cli`` / ``fpie.gui`` when already imported) to thin dispatching
   subclasses: every other backend goes to the stock class untouched,
   ``backend="b200"`` goes to this package.

Two wirings are offered for ``b200``:

* ``fused=True`` (default): this package's Processor -- preprocessing on the
  device from the uint8 images (``fpie_b200.process``);
* ``fused=False``: the *reference's own* Processor (host numpy preprocessing,
  fpie/process.py:192-271 / 321-386) with only the core solver swapped, i.e.
  exactly what ``core_cuda`` is to the reference.  The stock ``__init__``
  cannot be used for an unknown backend name (its error table raises
  ``KeyError``, process.py:94-103), so it is bypassed via ``BaseProcessor``.
"""

from __future__ import annotations

from . import process as b200_process
from .solver import EquSolver, GridSolver

BACKEND = b200_process.BACKEND


def _stage_io() -> None:
    """Rebind ``fpie.io.read_images`` / ``write_image`` (and the names ``fpie.cli`` / ``fpie.gui`` imported,
    cli.py:6) to this package's versions: concurrent decode into page-locked staging buffers, PNG encoding in
    a worker thread (``fpie_b200.io``).  Same arrays, same files."""
    import sys

    import fpie.io as fio

    from . import io as b200_io

    if getattr(fio, "_b200_staged", False):
        return
    fio.read_image, fio.read_images = b200_io.read_image, b200_io.read_images
    fio.write_image = b200_io.write_image_async
    for name in ("fpie.cli", "fpie.gui"):
        mod = sys.modules.get(name)
        if mod is not None:
            for attr in ("read_images", "read_image"):
                if hasattr(mod, attr):
                    setattr(mod, attr, getattr(fio, attr))
            if hasattr(mod, "write_image"):
                mod.write_image = fio.write_image
    fio._b200_staged = True


def register(fused: bool = True, make_default: bool = False, stage_io: bool = False):
    """Patch ``fpie`` in place; returns ``(EquProcessor, GridProcessor)`` dispatchers.  ``stage_io`` also
    swaps the image I/O helpers for the staged ones (callers must ``fpie_b200.io.flush_writes()`` before
    reading what was written; interpreter exit does it too)."""
    import sys

    import fpie.process as fp

    if stage_io:
        _stage_io()
    if getattr(fp, "_b200_registered", False):
        return fp.EquProcessor, fp.GridProcessor
    if BACKEND not in fp.ALL_BACKEND:
        fp.ALL_BACKEND.append(BACKEND)
    if make_default:
        fp.DEFAULT_BACKEND = BACKEND
    ref_equ, ref_grid, base = fp.EquProcessor, fp.GridProcessor, fp.BaseProcessor

    class B200CoreEquProcessor(ref_equ):
        """Reference EquProcessor (host preprocessing) over the b200 core."""

        def __init__(self, gradient="max", backend=BACKEND, n_cpu=0, min_interval=100, block_size=1024):
            base.__init__(self, gradient, 0, backend, EquSolver(block_size))

    class B200CoreGridProcessor(ref_grid):
        """Reference GridProcessor (host preprocessing) over the b200 core."""

        def __init__(self, gradient="max", backend=BACKEND, n_cpu=0, min_interval=100, block_size=1024, grid_x=8,
                     grid_y=8):
            base.__init__(self, gradient, 0, backend, GridSolver(grid_x, grid_y))

    def _pick(args, kwargs):
        return kwargs.get("backend", args[1] if len(args) > 1 else fp.DEFAULT_BACKEND)

    class EquProcessor(ref_equ):
        def __new__(cls, *args, **kwargs):
            if _pick(args, kwargs) == BACKEND:
                target = b200_process.EquProcessor if fused else B200CoreEquProcessor
                return target(*args, **kwargs)
            return object.__new__(cls)

    class GridProcessor(ref_grid):
        def __new__(cls, *args, **kwargs):
            if _pick(args, kwargs) == BACKEND:
                target = b200_process.GridProcessor if fused else B200CoreGridProcessor
                return target(*args, **kwargs)
            return object.__new__(cls)

    fp.EquProcessor, fp.GridProcessor = EquProcessor, GridProcessor
    fp.B200CoreEquProcessor, fp.B200CoreGridProcessor = B200CoreEquProcessor, B200CoreGridProcessor
    for name in ("fpie.cli", "fpie.gui"):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.EquProcessor, mod.GridProcessor = EquProcessor, GridProcessor
    fp._b200_registered = True
    return EquProcessor, GridProcessor


def main() -> None:
    """``python -m fpie_b200.register [fpie CLI flags]`` == ``fpie`` with ``-b b200`` available."""
    register(stage_io=True)
    from fpie.cli import main as fpie_main

    from .io import flush_writes

    try:
        fpie_main()
    finally:
        flush_writes()


def gui_main() -> None:
    """``fpie-gui`` with ``-b b200`` available (fpie/gui.py: reset + step per mouse click; a small edit is a
    few hundred microseconds of reset and a step on the device)."""
    register()
    from fpie.gui import main as fpie_gui_main

    fpie_gui_main()


if __name__ == "__main__":
    main()
