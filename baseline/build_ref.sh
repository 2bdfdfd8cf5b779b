#!/usr/bin/env bash
# Builds the UNMODIFIED reference (NVlabs/WarpConvNet under /root/reference) with its own setup.py
# for sm_100a and stages it under baseline/_ref/ (git-ignored, travels to the GPU box):
#   baseline/_ref/warpconvnet/      python package + _C.*.so
#   baseline/_ref/torch_scatter/    segment_csr shim (torch_scatter is not in this image)
# The reference's setup.py patches its vendored CUTLASS in place, so the tree is copied to a
# scratch directory first (/root/reference is read-only). Log: baseline/build_ref.log.
# Usage: bash baseline/build_ref.sh [MAX_JOBS]     (skips the build when the .so already exists)
set -euo pipefail
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
export TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS="$JOBS" NVCC_THREADS=1 SETUPTOOLS_SCM_PRETEND_VERSION=0.0.0
python setup.py build_ext --inplace 2>&1 | tee "$HERE/build_ref.full.log" | grep -v "^\[\|warning\|Warning\|note:" | tail -200 > "$HERE/build_ref.log" || true
ls warpconvnet/_C*.so
rm -rf "$OUT/warpconvnet"
mkdir -p "$OUT/warpconvnet"
# python sources + the built extension only (no csrc, no CUTLASS)
(cd "$WORK" && find warpconvnet -name '*.py' -o -name '_C*.so' -o -name '*.json' -o -name '*.msgpack' | grep -v '/csrc/' | cpio -pdm "$OUT" 2>/dev/null)
cp "$WORK"/warpconvnet/_C*.so "$OUT/warpconvnet/"
echo "staged $(du -sh "$OUT" | cut -f1) under $OUT"
