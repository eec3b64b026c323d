#!/usr/bin/env bash
# One gpurun call that re-establishes every measured number of DESIGN.md §7 on a fresh box (≈ 6 GPU-minutes at N = 1):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_checkpoint.sh r02'
# Outputs go to gpurun_out/ with the given round tag; copy what should be judged into profiles/.
set -uo pipefail
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; ( "$@" ) > $OUT/${TAG}_$name.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_$name.log; tail -2 $OUT/${TAG}_$name.log; }

run gpu_tests      timeout 420 python -m pytest tests -m gpu -x -q
run smoke          timeout 60  python -c "import __graft_entry__ as g; g.smoke()"
( timeout 300 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err ); tail -c 400 $OUT/${TAG}_bench_n1.json; echo
( timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err )
run reference_benchmarks timeout 120 python scripts/reference_benchmarks.py
run trace_config1  timeout 60  python scripts/trace_config1.py
run config_bench   timeout 200 python scripts/config_bench.py 8
# launch list of the bench itself (per-launch times are cold-cache and serialised: only the kernel's SHARE is comparable)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --no-tc > $OUT/${TAG}_ncu_bench.log 2>&1
echo "done: $(ls $OUT | grep -c "^${TAG}_") files"
