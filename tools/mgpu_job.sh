#!/bin/bash
# developer GPU job: every bench workload on N GPUs of one box (N = $1), outputs in gpurun_out/*_${N}gpu*.json
N=${1:-2}; TAG=${TAG:-r02}; PORT=${PORT:-29511}
mkdir -p gpurun_out
run() { if [ "$N" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@"; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@"; fi; }
for w in fe2 au108; do
  run --workload $w > gpurun_out/bench_${w}_${TAG}_${N}gpu.json 2> gpurun_out/bench_${w}_${TAG}_${N}gpu.err
  tail -c 300 gpurun_out/bench_${w}_${TAG}_${N}gpu.json | head -c 300; echo; PORT=$((PORT+1))
done
run --workload sweep ${SWEEP_ARGS:-} --out gpurun_out/sweep_${TAG}_${N}gpu.jsonl > gpurun_out/bench_sweep_${TAG}_${N}gpu.json 2> gpurun_out/bench_sweep_${TAG}_${N}gpu.err
cut -c1-200 gpurun_out/bench_sweep_${TAG}_${N}gpu.json; PORT=$((PORT+1))
run ${SI_ARGS:-} > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
cut -c1-300 gpurun_out/bench_${TAG}_${N}gpu.json
if [ "$N" != 1 ] && [ -z "${SKIP_EXTRA:-}" ]; then
  timeout 600 python -m pytest tests/test_chebfi_mgpu.py -m gpu -x -q > gpurun_out/mgpu_tests_${TAG}_${N}gpu.log 2>&1; tail -2 gpurun_out/mgpu_tests_${TAG}_${N}gpu.log
  run --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}_${N}gpu.json 2> gpurun_out/bench_ref_${TAG}_${N}gpu.err; cut -c1-200 gpurun_out/bench_ref_${TAG}_${N}gpu.json
fi
