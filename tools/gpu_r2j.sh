#!/bin/bash
# round 2: full GPU tests, the bench line of every BASELINE config that fits one GPU, peak probes
mkdir -p gpurun_out
export B2S_PARITY_REPORT=gpurun_out/parity_report_r2.jsonl
rm -f $B2S_PARITY_REPORT
timeout 1800 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2j_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_cfg2.json 2> gpurun_out/r2j_bench_cfg2.err; echo "bench cfg2 rc=$?"
python tools/show_bench.py gpurun_out/r2j_bench_cfg2.json 2>&1 | head -40
tail -3 gpurun_out/r2j_bench_cfg2.err
for c in cfg1 cfg3 cfg5; do
  timeout 900 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_$c.json 2> gpurun_out/r2j_bench_$c.err; echo "bench $c rc=$?"
  python tools/show_bench.py gpurun_out/r2j_bench_$c.json 2>&1 | head -4; tail -3 gpurun_out/r2j_bench_$c.err
done
timeout 300 python tools/peak_probe.py 2>&1 | tail -3
