#!/usr/bin/env bash
# Builds the UNMODIFIED reference (NVlabs/WarpConvNet under /root/reference) with its own setup.py
# for sm_100a and stages it under baseline/_ref/ (git-ignored, travels to the GPU box):
#   baseline/_ref/warpconvnet/      python package + _C.*.so
#   baseline/_ref/torch_scatter/    segment_csr shim (torch_scatter is not in this image)
# The reference's setup.py patches its vendored CUTLASS in place, so the tree is copied to a
# scratch directory first (/root/reference is read-only). Summary log: baseline/build_ref.log.
# Usage: bash baseline/build_ref.sh [MAX_JOBS]     (skips the build when the .so already exists)
# Measured here: 85 ninja steps, ~15 min wall on 8 cores with MAX_JOBS=6, 87 MB extension.
set -eu
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${WCN_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
WORK="${WCN_REF_WORK:-/tmp/wcn_ref_build}"
JOBS="${1:-6}"
[ -d "$REF" ] || { echo "reference tree $REF absent: nothing to build"; exit 0; }
if ls "$OUT"/warpconvnet/_C*.so >/dev/null 2>&1; then echo "baseline/_ref already built"; exit 0; fi
mkdir -p "$WORK" "$OUT"
if [ ! -f "$WORK/setup.py" ]; then cp -r "$REF"/. "$WORK"/; fi
cd "$WORK"
if ! ls warpconvnet/_C*.so >/dev/null 2>&1; then
  export TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS="$JOBS" NVCC_THREADS=1 SETUPTOOLS_SCM_PRETEND_VERSION=0.0.0
  python setup.py build_ext --inplace > "$HERE/build_ref.full.log" 2>&1 || true
fi
if [ -f "$HERE/build_ref.full.log" ]; then
  {
    echo "# reference build for sm_100a (python setup.py build_ext --inplace, TORCH_CUDA_ARCH_LIST=10.0a)"
    echo "ninja steps completed: $(grep -c '^\[[0-9]*/[0-9]*\]' "$HERE/build_ref.full.log" || true)"
    echo "lines matching 'error': $(grep -ci ' error' "$HERE/build_ref.full.log" || true)"
    grep -E "^(Adding gencode|Adding feature macro|TORCH_CUDA_ARCH_LIST|Using CUDA path)" "$HERE/build_ref.full.log" || true
    tail -n 1 "$HERE/build_ref.full.log" | cut -c1-200
  } > "$HERE/build_ref.log"
fi
ls warpconvnet/_C*.so
rm -rf "$OUT/warpconvnet"
mkdir -p "$OUT/warpconvnet"
# python sources (csrc/ holds python modules too) + the built extension; no .cu/.h, no CUTLASS
find warpconvnet \( -name '*.py' -o -name '*.json' -o -name '*.msgpack' \) -print0 |
  while IFS= read -r -d '' f; do mkdir -p "$OUT/$(dirname "$f")"; cp "$f" "$OUT/$f"; done
cp warpconvnet/_C*.so "$OUT/warpconvnet/"
# torch_scatter is not in this image: segment_csr shim (the reference imports it for pools / norms)
mkdir -p "$OUT/torch_scatter"
cp "$HERE/torch_scatter_shim.py" "$OUT/torch_scatter/__init__.py"
echo "staged $(du -sh "$OUT" | cut -f1) under $OUT"
