#!/usr/bin/env bash
# One gpurun call that re-establishes the measured numbers of DESIGN.md §4 / §7 on a fresh 1-GPU box (≈ 9 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_checkpoint.sh r03'
# and, on N GPUs of one box (charged N-fold; ≈ 1 minute of box time):
#   /usr/local/graft/bin/gpurun --gpus N --timeout 300 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
#       --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus N --steps 3 --warmup 3 > gpurun_out/rXX_bench_nN.json'
# Outputs go to gpurun_out/ with the given round tag; copy what should be judged into profiles/ (scripts/ncu_summarise.py
# <tag> turns the two ncu outputs into the tracked summaries).
set -uo pipefail
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; ( "$@" ) > $OUT/${TAG}_$name.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_$name.log; tail -2 $OUT/${TAG}_$name.log; }

run gpu_tests      timeout 500 python -m pytest tests -m gpu -x -q
run smoke          timeout 60  python -c "import __graft_entry__ as g; g.smoke()"
( timeout 330 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err ); tail -c 400 $OUT/${TAG}_bench_n1.json; echo
( timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err )
run reference_benchmarks timeout 150 python scripts/reference_benchmarks.py
run config_bench   timeout 200 python scripts/config_bench.py 8
run trace_config1  timeout 60  python scripts/trace_config1.py
run trace_config2  timeout 120 python scripts/trace_config2.py
run trace_config3  timeout 120 python scripts/trace_config3.py
run variant_bench  timeout 200 python scripts/variant_bench.py
# launch list of the bench itself (per-launch times are cold-cache and serialised: only the kernel's SHARE is comparable)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --no-tc > $OUT/${TAG}_ncu_bench.log 2>&1
# one full capture of the dominant kernel on one pullback pair (398 units)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -f -o $OUT/prof_sweep_u398 \
    python scripts/ncu_target.py 398 > $OUT/${TAG}_ncu_full.log 2>&1
echo "done: $(ls $OUT | grep -c "^${TAG}_") files"
