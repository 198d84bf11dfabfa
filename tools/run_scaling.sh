#!/bin/bash
# North-star multi-GPU workloads (BASELINE configs 3 and 5) and the default config 2 at N GPUs of one box.
#   gpurun --gpus 8 -- 'bash tools/run_scaling.sh 8'      -> gpurun_out/scale_r2_n8.jsonl
#   gpurun          -- 'bash tools/run_scaling.sh 1'      -> gpurun_out/scale_r2_n1.jsonl
N=${1:-1}
OUT=gpurun_out/scale_${2:-r2}_n${N}.jsonl
: > $OUT
run() {
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N "$@" 2>/dev/null | grep '^{' >> $OUT
  else
    python bench.py --gpus 1 "$@" 2>/dev/null | grep '^{' >> $OUT
  fi
}
run --config 3 --steps 3 --warmup 1 --no-cpu-baseline
run --config 5 --steps 2 --warmup 1 --no-cpu-baseline
run --config 2 --steps 5 --warmup 3 --no-cpu-baseline
if [ "$N" -gt 1 ]; then
  python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -2 > gpurun_out/sharded_n${N}.log
fi
python - <<PY
import json
for ln in open("$OUT"):
    d = json.loads(ln)
    print(d["config"]["baseline_config"], d["n_gpus"], round(d["value"], 1), round(d["e2e"]["value"], 1), d["ms_per_step"], d["per_rank_ms"])
PY
