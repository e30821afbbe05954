cd /root/repo
python - <<'PY'
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import torch
import dtv_utils_b200 as pb
n = 1 << 31
path = "/dev/shm/papr_cli_q.cfile"
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
with open(path, "wb") as f:
    step = 1 << 27
    for k in range(0, 2 * n, step):
        f.write(d[k:k + step].cpu().numpy().tobytes())
del d; eng.close(); torch.cuda.empty_cache()
try:
    for i in range(4):
        env = dict(os.environ, PAPR_B200_STATS="1")
        t0 = time.perf_counter()
        r = subprocess.run([pb.cli_path(), path], capture_output=True, env=env)
        dt = time.perf_counter() - t0
        print(f"wall {dt*1e3:.0f} ms | {r.stderr.decode().strip()}", flush=True)
    t0 = time.perf_counter(); subprocess.run(["/bin/true"]); print("spawn overhead ms", (time.perf_counter()-t0)*1e3)
finally:
    os.unlink(path)
PY
