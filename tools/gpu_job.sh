#!/bin/bash
# One gpurun call: GPU parity tests, bench, ncu launch list, ncu full captures.  Usage (from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/gpu_job.sh tests bench launches ncu'
# Outputs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
TAG=${TAG:-r01}
for what in "$@"; do
  case $what in
    tests)
      timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log ;;
    bench)
      timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
      tail -c 3000 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err ;;
    benchref)
      timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
      tail -c 2000 gpurun_out/bench_ref_${TAG}.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} \
        > gpurun_out/launches_${TAG}.log 2>&1
      tail -2 gpurun_out/launches_${TAG}.log ;;
    ncu)
      for k in ${NCU_KERNELS:-k_dgemm_nn k_dgemm_tn k_fw_plane k_fw_x_forward k_fw_x_backward}; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f \
          -o gpurun_out/ncu_${k}_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} \
          > gpurun_out/ncu_${k}_${TAG}.log 2>&1
        tail -1 gpurun_out/ncu_${k}_${TAG}.log
      done ;;
    *) echo "unknown section $what" ;;
  esac
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader > gpurun_out/smi_after.txt 2>&1
