"""fpie_b200 -- B200-native Jacobi Poisson backend for Fast-Poisson-Image-Editing.

Public surface (mirrors the reference's backend modules and Processor layer):

* ``EquSolver`` / ``GridSolver`` -- core solvers (``partition / reset / sync / step``)
* ``EquProcessor`` / ``GridProcessor`` -- image-level front end with device-side preprocessing
* ``register()`` -- make ``-b b200`` selectable in an installed ``fpie``

Importing the package does not need a GPU; constructing a solver does, and
fails loudly (RuntimeError / ImportError) otherwise -- there is no CPU path.
"""

__version__ = "0.1.0"

from .solver import EquSolver, GridSolver, device_count, device_info  # noqa: F401
from .process import BatchGridProcessor, EquProcessor, GridProcessor  # noqa: F401
from .register import register  # noqa: F401
