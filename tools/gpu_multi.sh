#!/bin/bash
# Multi-GPU evidence on one box: usage  gpurun --gpus N --timeout 2400 -- 'bash tools/gpu_multi.sh N TAG [full]'
#   bench.py under torchrun at N ranks in both sharding modes, the one-process device group (tools/groupbench.py), and
#   (full) the other BASELINE configs at N ranks.
N=${1:-2}
TAG=${2:-r02m}
FULL=${3:-}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 256 --warmup 32 > $O/bench_n${N}_frames_$TAG.log 2> $O/bench_n${N}_frames_$TAG.err; echo "bench frames N=$N rc=$?"
$TR bench.py --gpus $N --steps 256 --warmup 32 --mode templates > $O/bench_n${N}_templates_$TAG.log 2> $O/bench_n${N}_templates_$TAG.err; echo "bench templates N=$N rc=$?"
python tools/groupbench.py --frames 2048 --reps 4 > $O/groupbench_n${N}_$TAG.log 2> $O/groupbench_n${N}_$TAG.err; echo "groupbench rc=$?"; cat $O/groupbench_n${N}_$TAG.log
if [ -n "$FULL" ]; then
$TR bench.py --gpus $N --steps 128 --warmup 32 --config 4 --mode templates > $O/bench_n${N}_config4_templates_$TAG.log 2> $O/bench_n${N}_config4_templates_$TAG.err; echo "config 4 templates N=$N rc=$?"
$TR bench.py --gpus $N --steps 128 --warmup 32 --config 4 > $O/bench_n${N}_config4_frames_$TAG.log 2> $O/bench_n${N}_config4_frames_$TAG.err; echo "config 4 frames N=$N rc=$?"
$TR bench.py --gpus $N --steps 128 --warmup 32 --config 5 > $O/bench_n${N}_config5_$TAG.log 2> $O/bench_n${N}_config5_$TAG.err; echo "config 5 N=$N rc=$?"
fi
for f in $O/bench_n${N}_*_$TAG.log; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "fps", "ms_per_step", "timed_repeats")}, "e2e", {k: d["e2e"][k] for k in ("fps", "ms_per_step")}, d["config"]["mode"], d["config"]["templates_total"], d.get("parity"))
except Exception as e:
    print("unreadable:", e)
PY
done
for f in $O/bench_n${N}_*_$TAG.err; do tail -n 2 $f; done
