"""Environment + cuBLAS DGEMM probe for the GPU box (run under gpurun). Writes gpurun_out/probe_box.json."""
import json, os, shutil, subprocess, time
import torch

info = {"nproc": os.cpu_count()}
with open("/proc/meminfo") as f:
    for line in f:
        k, v = line.split(":")
        if k in ("MemTotal", "MemAvailable"):
            info[k] = int(v.split()[0]) // 1024  # MiB
info["Rscript"] = shutil.which("Rscript")
info["gpu"] = torch.cuda.get_device_name(0)
info["n_gpus"] = torch.cuda.device_count()
try:
    info["lscpu"] = subprocess.run("lscpu | grep -E 'Model name|Socket|Core|Thread'", shell=True, capture_output=True, text=True).stdout
except Exception as e:  # noqa
    info["lscpu"] = str(e)

def dgemm(n, reps=10):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained 4 s
    t0 = time.time(); n_it = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(5):
            torch.matmul(a, b, out=c); n_it += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n_it
    return {"n": n, "burst_tflops": 2 * n**3 / best * 1e-9, "sustained_tflops": 2 * n**3 / sus * 1e-9}

info["dgemm_8192"] = dgemm(8192)
info["dgemm_4096"] = dgemm(4096, reps=10)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe_box.json", "w") as f:
    json.dump(info, f, indent=1)
print(json.dumps(info, indent=1))
