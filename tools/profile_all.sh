#!/bin/bash
# Round-end profiling recipe (B200_PROFILING.md): plain bench first, then the ncu launch list, then one
# --set full capture per hot kernel.  Run on the GPU box:  bash tools/profile_all.sh <tag> [all]
# Numbers printed under ncu are never bench values.  Each capture is summarised on the box (tools/ncu_summary.py ->
# <tag>_<name>_details.txt / _raw_metrics.txt) and the .ncu-rep deleted: gpurun brings back at most 64 MiB.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || echo "bench failed"
tail -c 600 $OUT/bench_${TAG}.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 > $OUT/ncu_${TAG}_list.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {   # name, evals per launch, kernel regex, skip, command...
    local name=$1 evals=$2 rx=$3 skip=$4; shift 4
    timeout 240 $NCU -k regex:$rx -s $skip -c 1 -o $OUT/prof_${TAG}_${name} "$@" > $OUT/ncu_${TAG}_${name}.log 2>&1
    python tools/ncu_summary.py $OUT/prof_${TAG}_${name}.ncu-rep $OUT/${TAG}_${name} $evals > /dev/null 2>&1 || echo "summary of $name failed"
    rm -f $OUT/prof_${TAG}_${name}.ncu-rep
}
# the headline kernel at the bench's own launch size (1e6 samples: 17.6 waves; at 2e5 the last partial wave costs 10 %)
cap ts_cfg2 4096000000 qo_mc_ts 3 python bench.py --steps 2 --warmup 3 --no-extras --samples 1000000
cap ts_cfg5 4096000000 qo_mc_ts 3 python bench.py --steps 2 --warmup 3 --no-extras --samples 1000000 --workload cfg5
if [ "${2:-}" = "all" ]; then
export QO100NET_KERNEL=tf
cap tf_cfg2 819200000 qo_mc_tf 3 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000
cap tf_cfg5 819200000 qo_mc_tf 3 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000 --workload cfg5
unset QO100NET_KERNEL
export QO100NET_KERNEL=ladder
cap ladder_cfg2 819200000 ladder 3 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000
cap ladder_cfg5 819200000 ladder 3 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000 --workload cfg5
unset QO100NET_KERNEL
cap fulls_tf 268435456 qo_fs_tf 2 python bench.py --steps 1 --warmup 3 --samples 200000 --north-star-samples 2000000
cap nodal 4000000 qo_nodal_kernel 1 python tools/nodal_bench.py --samples 4000
cap nodal_jit 100000000 qo_nodal_jit 1 python tools/nodal_bench.py --samples 100000
cap board_cfg3 6000000 qo_mc_board 2 python tools/cfg3_run.py
cap generic_cfg3 6000000 generic 2 python tools/cfg3_run.py
fi
ls -la $OUT/*${TAG}*
