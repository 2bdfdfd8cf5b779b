#!/usr/bin/env bash
# SPDX-License-Identifier: Apache-2.0
# Next-round experiment (profiles/r1i): A/B of the forward gather-GEMM with zero rows written by
# st.shared in a second pass (WCN_ZERO_ROWS_SECOND_PASS) against the default zero-size cp.async.
# Run on the GPU box from the repo root:  gpurun --timeout 300 -- 'bash tools/exp_zero_rows.sh'
# The variant is built IN PLACE, parity-tested, timed, and the default build is restored at the end.
set -uo pipefail
mkdir -p gpurun_out
python tools/exp_fwd.py > gpurun_out/zero_rows_default.log 2>&1
WCN_ZERO_ROWS_SECOND_PASS=1 bash warpconvnet_b200/csrc/build.sh > gpurun_out/zero_rows_build.log 2>&1
python tools/exp_fwd.py > gpurun_out/zero_rows_second_pass.log 2>&1
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q \
  > gpurun_out/zero_rows_parity.log 2>&1
python bench.py --no-cpu-baseline --no-c4 --no-side --ref-gpu none --dist R > gpurun_out/zero_rows_bench_R.json 2>/dev/null
python bench.py --no-cpu-baseline --no-c4 --no-side --ref-gpu none > gpurun_out/zero_rows_bench_S.json 2>/dev/null
bash warpconvnet_b200/csrc/build.sh >> gpurun_out/zero_rows_build.log 2>&1   # restore the default
grep -h "plan=\|stages=" gpurun_out/zero_rows_default.log gpurun_out/zero_rows_second_pass.log
tail -2 gpurun_out/zero_rows_parity.log
