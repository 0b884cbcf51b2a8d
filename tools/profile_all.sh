#!/bin/bash
# Round-end profiling recipe (B200_PROFILING.md): plain bench first, then the ncu launch list, then one
# --set full capture per hot kernel.  Run on the GPU box:  bash tools/profile_all.sh <tag>
# Numbers printed under ncu are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || echo "bench failed"
tail -c 600 $OUT/bench_${TAG}.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 > $OUT/ncu_${TAG}_list.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:qo_mc_tf -s 3 -c 1 -o $OUT/prof_${TAG}_tf_cfg2 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000 > $OUT/ncu_${TAG}_a.log 2>&1
timeout 200 $NCU -k regex:qo_mc_tf -s 3 -c 1 -o $OUT/prof_${TAG}_tf_cfg5 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000 --workload cfg5 > $OUT/ncu_${TAG}_b.log 2>&1
if [ "${2:-}" = "all" ]; then
QO100NET_KERNEL=ladder timeout 200 $NCU -k regex:ladder -s 3 -c 1 -o $OUT/prof_${TAG}_ladder_cfg2 python bench.py --steps 2 --warmup 3 --no-extras --samples 200000 > $OUT/ncu_${TAG}_c.log 2>&1
timeout 200 $NCU -k regex:lumped -s 2 -c 1 -o $OUT/prof_${TAG}_fulls python bench.py --steps 1 --warmup 3 --samples 20000 > $OUT/ncu_${TAG}_d.log 2>&1
timeout 200 $NCU -k regex:nodal -s 1 -c 1 -o $OUT/prof_${TAG}_nodal python tools/nodal_bench.py --samples 4000 > $OUT/ncu_${TAG}_e.log 2>&1
fi
ls -la $OUT/*${TAG}*
