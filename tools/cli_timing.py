"""Developer probe: wall-clock of the drop-in CLI on a capture file in tmpfs (page cache), against the
reference binary on a head of the same file.  Not the contract bench."""
import os, shutil, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dtv_utils_b200 as pb

free = shutil.disk_usage("/dev/shm").free
log2n = 31 if free > (20 << 30) else 30 if free > (10 << 30) else 29
n = 1 << log2n
path = "/dev/shm/papr_cli_timing.cfile"
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
t0 = time.perf_counter()
with open(path, "wb") as f:
    step = 1 << 27
    for k in range(0, 2 * n, step):
        f.write(d[k:k + step].cpu().numpy().tobytes())
print(f"wrote {n*8>>30} GiB to {path} in {time.perf_counter()-t0:.1f} s", flush=True)
del d
eng.close()
torch.cuda.empty_cache()
cli = pb.cli_path()

def run(env_extra, args, label):
    env = dict(os.environ, PAPR_B200_STATS="1", **env_extra)
    t0 = time.perf_counter()
    r = subprocess.run([cli] + args + [path], capture_output=True, env=env)
    dt = time.perf_counter() - t0
    print(f"[{label}] wall {dt*1e3:8.1f} ms  -> {n*8/dt/1e9:6.2f} GB/s of file, rc={r.returncode}  | {r.stderr.decode().strip().splitlines()[-1] if r.stderr else ''}", flush=True)
    return r

try:
    r0 = run({}, [], "1dB cold")
    r1 = run({"PAPR_B200_TRACE": "1"}, [], "1dB trace")
    print(r1.stderr.decode(), flush=True)
    run({}, [], "1dB warm")
    run({"PAPR_B200_EXACT_SUM": "0"}, [], "1dB exact_sum=0")
    run({"PAPR_B200_MAX_RESIDENT_MB": "1024"}, [], "1dB re-streamed (1 GiB budget)")
    rg = run({}, ["-g"], "-g")
    # the reference on the first 1 GiB (2^27 samples)
    head = "/dev/shm/papr_cli_timing_head.cfile"
    with open(path, "rb") as f, open(head, "wb") as g:
        g.write(f.read(1 << 30))
    ref = os.path.join(ROOT, "oracle", "_ref", "papr")
    for args in ([], ["-g"]):
        t0 = time.perf_counter(); want = subprocess.run([ref] + args + [head], capture_output=True).stdout; tr = time.perf_counter() - t0
        t0 = time.perf_counter(); got = subprocess.run([cli] + args + [head], capture_output=True).stdout; tg = time.perf_counter() - t0
        print(f"1 GiB head {args}: reference {tr:.2f} s, drop-in {tg:.2f} s, identical={want == got}", flush=True)
    os.unlink(head)
finally:
    if os.path.exists(path):
        os.unlink(path)
