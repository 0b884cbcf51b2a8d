#!/bin/bash
# One GPU call on the code as it stands: the GPU test suite, the bench line, the compiled-chain speed table and one ncu capture
# of the compiled chain kernel (config 2 kept off the polynomial kernels).  bash tools/final_check.sh <tag>
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests_${TAG}.log 2>&1; echo "tests rc=$?" | tee -a $OUT/gpu_tests_${TAG}.log
tail -3 $OUT/gpu_tests_${TAG}.log
timeout 300 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
tail -c 300 $OUT/bench_${TAG}.json
timeout 120 python tools/chain_jit_speed.py --out $OUT/chain_jit_${TAG}.json > $OUT/cj_speed_${TAG}.log 2>&1; echo "speed rc=$?"
export QO100NET_KERNEL=interp QO100NET_CHAIN=jit
timeout 150 ncu --set full --clock-control none --import-source on -f -k regex:qo_mc_chain_jit -s 3 -c 1 -o $OUT/prof_chain \
    python bench.py --steps 2 --warmup 3 --no-extras --samples 200000 > $OUT/ncu_${TAG}_chain.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $OUT/prof_chain.ncu-rep $OUT/${TAG}_chain_jit_cfg2 819200000 > /dev/null 2>&1 || echo "summary failed"
rm -f $OUT/prof_chain.ncu-rep
ls $OUT | tail -20
