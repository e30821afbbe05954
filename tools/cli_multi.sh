#!/bin/bash
# `PAPR_B200_DEVICES=N papr <file>` on a 16 GiB capture in tmpfs, N = 1, 2, 4, 8 (as many as the box has):
# analysis wall time of the single-process multi-GPU file driver (papr_multi_*), stdout compared across N.
cd "$(dirname "$0")/.."
F=/dev/shm/papr_b200_multi.cfile
python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import torch, dtv_utils_b200 as pb
n = 1 << 31
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
torch.cuda.synchronize()
with open("$F", "wb") as f:
    step = 1 << 27
    for k in range(0, 2 * n, step):
        f.write(d[k:k + step].cpu().numpy().tobytes())
print("wrote", os.path.getsize("$F") >> 30, "GiB")
PY
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  for rep in 1 2; do
    /usr/bin/env PAPR_B200_DEVICES=$N PAPR_B200_STATS=1 bash -c "time dtv-utils_b200/bin/papr $F > /tmp/out_$N.txt" 2>&1 | grep -E "papr_b200|real" | tr '\n' ' '
    echo " [N=$N rep=$rep md5=$(md5sum < /tmp/out_$N.txt | cut -c1-8)]"
  done
done
rm -f $F
