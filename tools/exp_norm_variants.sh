# bring-up: builds libwcn_b200.so with different row-norm kernel parameters ON the GPU box and runs
# tools/exp_norm.py for each (the default build is restored by the caller)
set -e
for v in "-DWCN_RN_CLUSTER=1" "-DWCN_RN_SKIP_GLOBAL"; do
  WCN_EXTRA_FLAGS="$v" bash warpconvnet_b200/csrc/build.sh > /dev/null 2>&1
  echo "=== $v"
  python tools/exp_norm.py 2>&1 | grep -E "stats|reduce|one call"
done
