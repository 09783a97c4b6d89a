#!/bin/bash
# One GPU box, one pass: parity tests, smoke, both bench arms, then the ncu evidence (each capture only after its command
# exited 0 without ncu).  Outputs under gpurun_out/ with the tag given as $1 (default r01n); summarise here with
# profiles/summarize.py.   usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r01n'
TAG=${1:-r01n}
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
python bench.py > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference > $O/bench_ref_$TAG.log 2> $O/bench_ref_$TAG.err; echo "reference arm rc=$?"
python tools/trainbench.py > $O/trainbench_$TAG.log 2>&1; echo "trainbench rc=$?"; cat $O/trainbench_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 64 --warmup 4 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_similarity_coarse_rec -s 20 -c 3 -f -o $O/coarse_$TAG \
    python bench.py --steps 64 --warmup 4 --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
cat > /tmp/train64.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import test_oracle_render as golden
from linemod_pose_estimation_b200 import Detector, Mesh, training
views, idx = golden._oracle_views()
T = np.array([views[i][0] for i in idx[:64]]); up = np.array([views[i][1] for i in idx[:64]])
print(Detector().trainViews(Mesh(golden.G["triangles"]), golden._golden_camera(training), T, up, "obj")[0])
PY
python /tmp/train64.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file $O/launches_train_$TAG.csv python /tmp/train64.py > $O/ncu_train_$TAG.log 2>&1; echo "ncu trainer launch list rc=$?"
tail -c 600 $O/bench_$TAG.log
