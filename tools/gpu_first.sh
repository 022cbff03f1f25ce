#!/bin/bash
# First GPU pass: parity suite without the tcgen05 kernels, then the tcgen05 tests on their own (so that a hang
# there cannot block the rest), then smoke + a short bench.  Everything is wrapped in `timeout`.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== suite with B2S_DISABLE_TC=1 (SIMT conv) ===" | tee gpurun_out/pytest_simt.log
B2S_DISABLE_TC=1 timeout -k 10 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_conv.py -p no:cacheprovider 2>&1 | tail -60 >> gpurun_out/pytest_simt.log
B2S_DISABLE_TC=1 timeout -k 10 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "not 2-" -p no:cacheprovider 2>&1 | tail -60 >> gpurun_out/pytest_simt.log
tail -25 gpurun_out/pytest_simt.log
echo "=== tcgen05 conv tests ===" | tee gpurun_out/pytest_tc.log
timeout -k 10 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider 2>&1 | tail -80 >> gpurun_out/pytest_tc.log
TC_RC=${PIPESTATUS[0]}
tail -40 gpurun_out/pytest_tc.log
if grep -q "failed" gpurun_out/pytest_tc.log || [ "$TC_RC" != "0" ]; then
  echo "tcgen05 tests not green -> bench with SIMT"; export B2S_DISABLE_TC=1
fi
echo "=== smoke ===" 
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench (short) ==="
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --plots-per-gpu ${PLOTS:-32} 2>&1 | tail -5 | tee gpurun_out/bench_first.log
