#!/bin/bash
# Round-2 final multi-GPU session (gpurun --gpus 8): the bench line of the final build at 8 GPUs, launched the way the driver
# launches it (torchrun, one rank per GPU, NCCL).
mkdir -p gpurun_out
for n in ${NLIST:-8}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 --no-eager > gpurun_out/r02_bench_${n}gpu_final3.json 2> gpurun_out/r02_bench_${n}gpu_final3.err
  tail -c 300 gpurun_out/r02_bench_${n}gpu_final3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_${n}gpu_final3.json").read().strip().splitlines()[-1])
    print($n, "gpus: value %.0f ms %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]),
          {k: (round(c.get("audio_s_per_s", 0)), round(c.get("ms_per_step", c.get("ms_per_step_back_to_back", 0)), 2)) for k, c in d["configs"].items() if "error" not in c},
          {k: c["error"] for k, c in d["configs"].items() if "error" in c})
except Exception as e:
    print($n, "gpus: no line", e)
PY
done
