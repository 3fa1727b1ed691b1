"""Host logic of the Processor's private copy of the target (fpie_b200/process.py `_reset_with_canvas`; the
reference: fpie/process.py:268, 384 `self.tgt = tgt.copy()`): only what lies OUTSIDE the blend's bounding box is
copied from the caller's array -- by several threads for a large image, started the moment the core announces the
box -- and the inside is fetched from the device on first use.  A stand-in core plays the device."""

import numpy as np
import pytest

from fpie_b200 import process


class FakeCore:
    """Announces `box` through the hook (as the C ABI does once the mask's bounding box is known) or not at all."""

    def __init__(self, box, announce=True):
        self.box, self.announce, self.hook, self.fetches = box, announce, None, 0

    def on_box(self, fn):
        self.hook = fn

    def reset(self):
        if self.announce and self.hook is not None:
            self.hook(self.box)
        return 123, self.box


class Proc(process.BaseProcessor):
    def __init__(self, core):
        super().__init__("max", "b200", core)

    def reset(self, tgt):
        n, box = self._reset_with_canvas(tgt, self.core.reset)
        self.box = box
        return n

    def _fetch_into_canvas(self, iteration):
        x0, x1, y0, y1 = self.box
        self._tgt[x0:x1, y0:y1] = 200  # what the device would hand back for the box
        self.core.fetches += 1
        self._canvas_stale = False
        return np.zeros(3, np.float32)


@pytest.mark.parametrize("shape", [(40, 50, 3), (1800, 1700, 3)])  # (the second is copied by four threads)
@pytest.mark.parametrize("announce", [True, False])
@pytest.mark.parametrize("where", ["inside", "whole", "corner"])
def test_only_the_outside_of_the_box_comes_from_the_callers_array(shape, announce, where):
    rows, cols = shape[:2]
    box = {"inside": (rows // 4, rows // 2, cols // 3, cols - 5), "whole": (0, rows, 0, cols),
           "corner": (0, rows // 3, 0, cols // 2)}[where]
    rng = np.random.default_rng(0)
    tgt = rng.integers(0, 199, shape, dtype=np.uint8)
    proc = Proc(FakeCore(box, announce))
    for _ in range(2):  # (the second reset recycles the canvas of the first)
        mine = tgt.copy()
        assert proc.reset(mine) == 123
        mine[...] = 255  # every read of the caller's array has completed
        assert proc.core.hook is None  # the hook does not outlive the reset
        out = proc.tgt  # first use: the inside of the box arrives from the "device"
        x0, x1, y0, y1 = box
        want = tgt.copy()
        want[x0:x1, y0:y1] = 200
        np.testing.assert_array_equal(out, want)
        assert proc.tgt is out and proc.core.fetches >= 1


def test_a_failing_reset_leaves_no_canvas_and_no_hook():
    class Failing(FakeCore):
        def reset(self):
            raise RuntimeError("reset: the mask is empty")

    proc = Proc(Failing((0, 1, 0, 1)))
    with pytest.raises(RuntimeError, match="empty"):
        proc.reset(np.zeros((9, 9, 3), np.uint8))
    assert proc.core.hook is None and proc._tgt is None
    with pytest.raises(RuntimeError, match="before reset"):
        proc._require_reset()
