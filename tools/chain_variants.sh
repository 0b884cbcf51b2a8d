#!/bin/bash
# the compiled chain kernel built for 2 / 3 resident blocks per SM, per-sample coefficients hoisted into registers or read in the loop
mkdir -p gpurun_out
for v in "2 regs" "3 loop" "3 regs" "2 loop"; do
  set -- $v
  QO100NET_CHAIN_MINB=$1 QO100NET_CHAIN_COEF=$2 timeout 100 python tools/chain_jit_speed.py --jit-only --fs-samples 8192 --out gpurun_out/cjv_$1_$2.json > gpurun_out/cjv_$1_$2.log 2>&1
  echo "== MINB=$1 COEF=$2 rc=$?"; grep -h '"evals_per_s"\|"n_pass"' gpurun_out/cjv_$1_$2.json | tr -d ' \n'; echo
done
