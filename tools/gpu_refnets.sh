#!/bin/bash
# Runs tests/test_gpu_reference_nets.py on the GPU box against a TRANSIENT copy of the reference's network files
# (made by the caller into .ref_tmp/, git-ignored, deleted after the call): the reference tree itself does not travel.
mkdir -p gpurun_out
export B2S_REFERENCE_TREE=$PWD/.ref_tmp/torch-points3d
export B2S_PARITY_REPORT=gpurun_out/r02_reference_nets.jsonl
rm -f $B2S_PARITY_REPORT
timeout 1200 python -m pytest tests/test_gpu_reference_nets.py -q -s --timeout 600 > gpurun_out/r02_reference_nets_gpu.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/r02_reference_nets_gpu.log; cat $B2S_PARITY_REPORT
