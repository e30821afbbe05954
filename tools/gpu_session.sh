#!/bin/bash
# what the last 1-GPU gpurun call of the development loop ran; outputs land in gpurun_out/
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in - tools/_tune/libpapr_early.so tools/_tune/libpapr_promo256.so tools/_tune/libpapr_earlypromo.so -; do
  timeout 90 python tools/ab_probe.py $v 2>&1 | tail -2 | tee -a $O/ab_variants.txt
done
