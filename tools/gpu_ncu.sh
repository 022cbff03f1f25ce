#!/bin/bash
# ncu evidence (1 GPU): launch list of the timed graph replays of the bench command + one --set full capture of one
# instance of every stage (tools/ncu_target.py).  Reports stay on the box (they exceed the copy-back limit); the CSV
# exports and summaries come back in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
OURS='regex:gather_gemm|wgrad|conv_lines|kernel_map|coordmap|quantize|colreduce|bn_|flat_kernel|maxpool|segment|se_gate|se_hidden|se_pooled|se_param|gated|bcast|adabelief|parity|pair_|scan_|popc|lines_|prep_weights|pad_rows|batch_counts|gather_rows|col_partials|plot_|grad_check|init_bounds'
if [ "${SKIP_LIST:-0}" != "1" ]; then
echo "=== launch list (the two timed graph replays of bench.py) ==="
B2S_NCU_RANGE=1 timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200; wc -l gpurun_out/${TAG}_launches.csv
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv 2 > gpurun_out/${TAG}_launches_summary.txt; head -40 gpurun_out/${TAG}_launches_summary.txt
fi
echo "=== full set, one instance per stage ==="
timeout -k 10 1200 ncu --set full --clock-control none -k "$OURS" -f -o /tmp/${TAG}_stages python tools/ncu_target.py > gpurun_out/${TAG}_ncu_stages.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_stages.log; ls -la /tmp/${TAG}_stages.ncu-rep
ncu -i /tmp/${TAG}_stages.ncu-rep --page raw --csv > gpurun_out/${TAG}_stages_raw.csv 2>/dev/null
python tools/ncu_instances.py gpurun_out/${TAG}_stages_raw.csv gpurun_out/ncu_manifest.json gpurun_out/${TAG}_ncu_instances.json MEASURED_PEAKS.json | tee gpurun_out/${TAG}_ncu_instances.txt
python tools/ncu_summary.py gpurun_out/${TAG}_stages_raw.csv > gpurun_out/${TAG}_stages_summary.txt 2>/dev/null
ls -la gpurun_out/${TAG}_*
