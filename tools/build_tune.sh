#!/bin/bash
# builds tools/_tune/scan_tune_<threads>_<u>_<ctas> variants
set -e
cd "$(dirname "$0")"
mkdir -p _tune
gcc -O2 -ffp-contract=off -c ../dtv-utils_b200/csrc/papr_host.c -o _tune/papr_host.o
for cfg in "1024 4 1" "768 4 1" "1024 3 1"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ftz=false -DPAPR_THREADS=$1 -DPAPR_U=$2 -DPAPR_CTAS_PER_SM=$3 \
       scan_tune.cu _tune/papr_host.o -o _tune/scan_tune_$1_$2_$3 2>&1 | grep -E "error" || true
done
ls _tune
