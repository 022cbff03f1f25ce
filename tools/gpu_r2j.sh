#!/bin/bash
# round 2: what the driver runs at round end (GPU tests, smoke, both bench arms) + the other BASELINE configs
mkdir -p gpurun_out
export B2S_PARITY_REPORT=gpurun_out/parity_report_r2.jsonl
rm -f $B2S_PARITY_REPORT
timeout 1800 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2j_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r2j_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_cfg2.json 2> gpurun_out/r2j_bench_cfg2.err; echo "bench cfg2 rc=$?"
python tools/show_bench.py gpurun_out/r2j_bench_cfg2.json 2>&1 | head -14
tail -3 gpurun_out/r2j_bench_cfg2.err
if [ "${ALL_CONFIGS:-0}" = "1" ]; then
for c in cfg1 cfg3 cfg5; do
  timeout 900 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_$c.json 2> gpurun_out/r2j_bench_$c.err; echo "bench $c rc=$?"
  python tools/show_bench.py gpurun_out/r2j_bench_$c.json 2>&1 | head -4; tail -3 gpurun_out/r2j_bench_$c.err
done
fi
