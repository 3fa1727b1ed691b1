#!/bin/bash
# Installs the UNMODIFIED reference (fpie 0.3.2) into baseline/_ref (git-ignored, shipped to the GPU box by gpurun):
#   1. pip install from a scratch copy of /root/reference (the build writes into its source tree, which is
#      read-only).  The reference's setup.py drives CMake, whose configure step git-clones pybind11
#      (CMakeLists.txt:10-22) -- offline that fails, setup.py swallows the error (setup.py:84-86) and the wheel
#      carries the pure-Python package only (numpy / numba backends);
#   2. the native cores compiled from the same unmodified sources by oracle/Makefile (core_openmp, core_gcc,
#      core_cuda for sm_100a) are copied next to it, which is where fpie/process.py:52-85 imports them from.
# Result: `PYTHONPATH=baseline/_ref python -m fpie.cli -b {numpy,gcc,openmp,cuda} ...` is the stock reference.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
DEST="$HERE/_ref"
if [ ! -d "$REF/fpie" ]; then
  echo "install_ref: $REF not present; keeping whatever is in $DEST" >&2
  exit 0
fi
SCRATCH="$(mktemp -d)"
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REF" "$SCRATCH/src"
rm -rf "$DEST"
python -m pip install --quiet --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
  --target "$DEST" "$SCRATCH/src" >"$SCRATCH/pip.log" 2>&1 || { cat "$SCRATCH/pip.log" >&2; exit 1; }
make -s -C "$HERE/../oracle" ref ref-cuda REF="$REF" || echo "install_ref: native reference cores not rebuilt" >&2
cp -f "$HERE"/../oracle/_ref/core_*.so "$DEST/fpie/" 2>/dev/null || true
find "$DEST" -name __pycache__ -prune -exec rm -rf {} +
echo "installed reference into $DEST:" $(ls "$DEST/fpie")
