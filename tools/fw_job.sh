#!/bin/bash
# developer GPU job: fourwf parity tests + per-kernel timing of the fused path under several tunings
mkdir -p gpurun_out
TAG=${TAG:-r02a}
timeout 900 python -m pytest tests/test_fourwf_gpu.py -m gpu -x -q > gpurun_out/fw_tests_${TAG}.log 2>&1; tail -5 gpurun_out/fw_tests_${TAG}.log
timeout 600 python tools/tune_fourwf.py --workload si512 --ndat 128 --istwfk 2 --reps 5 ${TUNE_ARGS:---set half=0,1 --set half_cfg=0,1} > gpurun_out/fw_tune_${TAG}.jsonl 2>&1; cat gpurun_out/fw_tune_${TAG}.jsonl
