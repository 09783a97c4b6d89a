#!/bin/bash
# One GPU box, one pass: parity tests, smoke, both bench arms (driver's arguments and defaults), then the ncu evidence (each
# capture only after its command exited 0 without ncu).  Outputs under gpurun_out/ with the tag given as $1; summarise here
# with profiles/summarize.py.   usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02a [quick]'
TAG=${1:-r02a}
QUICK=${2:-}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu_$TAG.txt 2>&1; nproc >> $O/gpu_$TAG.txt
python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_$TAG.log 2> $O/bench_driver_$TAG.err; echo "bench (driver args) rc=$?"
python bench.py > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"
python tools/kbench.py --configs "batch_frames=16,streams=4;batch_frames=8,streams=4;batch_frames=16,streams=4,refine_tiled=0;batch_frames=16,streams=4,coarse_narrow=0;batch_frames=16,streams=4,coarse_share=0;batch_frames=16,streams=4,prune=0;batch_frames=1,streams=8" > $O/kbench_$TAG.log 2>&1; echo "kbench rc=$?"; cat $O/kbench_$TAG.log
python tools/kbench.py --workload stress --configs "batch_frames=16,streams=4;batch_frames=8,streams=4" > $O/kbench_stress_$TAG.log 2>&1; echo "kbench stress rc=$?"; cat $O/kbench_stress_$TAG.log
if [ -z "$QUICK" ]; then
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_$TAG.log 2> $O/bench_ref_$TAG.err; echo "reference arm rc=$?"
# the profiled runs load the trained template set from a cache written by a plain run: no trainer launches in the lists
export LM_BENCH_TEMPLATE_CACHE=/tmp/lm_bench_templates.lmb2
python bench.py --steps 64 --warmup 8 --no-cpu-baseline > $O/bench_cache_$TAG.log 2>&1; echo "cache-writing run rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 500 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 64 --warmup 8 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launch list rc=$?"
for K in k_similarity_coarse_rec63 k_cg_fused k_dn_fused k_spread_all k_pyrdown_fast k_refine_nib; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 -f -o $O/${K}_$TAG \
    python bench.py --steps 64 --warmup 8 --no-cpu-baseline > $O/ncu_full_${K}_$TAG.log 2>&1; echo "ncu full $K rc=$?"
done
unset LM_BENCH_TEMPLATE_CACHE
for C in 3 4 5; do
python bench.py --config $C --steps 64 --warmup 8 --no-cpu-baseline > $O/bench_c${C}_$TAG.log 2> $O/bench_c${C}_$TAG.err; echo "config $C rc=$?"; tail -c 600 $O/bench_c${C}_$TAG.log; echo
done
python tools/trainbench.py > $O/trainbench_$TAG.log 2>&1; echo "trainbench rc=$?"; cat $O/trainbench_$TAG.log
fi
tail -c 1500 $O/bench_driver_$TAG.log; echo; tail -c 3000 $O/bench_$TAG.log
tail -5 $O/bench_$TAG.err
