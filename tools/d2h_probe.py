"""Host-sink probe for a multi-GPU box: what can the HOST take from n GPUs copying device -> host at once?
(python tools/d2h_probe.py  -> gpurun_out/d2h_probe.json)

Round 1 measured 8 x 29.2 GB in 2.57 s = 91 GB/s through eight processes against 57 GB/s for one GPU alone;
this tells whether that is the box's ceiling.  For n in {1, 2, 4, 8} GPUs: every GPU streams 256 MiB chunks
from HBM into its own pinned host buffer (one process, one stream per device, all in flight together), GB/s
from the wall clock around a synchronize of all devices.  Variants: default allocation, an allocating thread
bound to the device's local CPUs (first touch on its node), and interleaved over the NUMA nodes.  Also
recorded: the sysfs topology (PCI bus id -> numa_node / local_cpulist), lscpu, and the host's own memcpy
bandwidth with 1..32 threads (the pageable path's second hop)."""
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GiB = 1 << 30
CHUNK = 256 << 20
PER_GPU = 4 * GiB


def sysfs(dev):
    p = torch.cuda.get_device_properties(dev)
    bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    out = {"bdf": bdf}
    for k in ("numa_node", "local_cpulist", "current_link_speed", "current_link_width", "max_link_speed"):
        try:
            out[k] = open(f"/sys/bus/pci/devices/{bdf}/{k}").read().strip()
        except OSError as e:
            out[k] = f"? ({e.__class__.__name__})"
    return out


def parse_cpulist(s):
    cpus = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            cpus.append(int(part))
    return cpus


def set_mempolicy_interleave(on):
    libc = ctypes.CDLL(None, use_errno=True)
    SYS_set_mempolicy = 238  # x86_64
    if on:
        n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        mask = ctypes.c_ulong((1 << n_nodes) - 1)
        return libc.syscall(SYS_set_mempolicy, 3, ctypes.byref(mask), 65)  # MPOL_INTERLEAVE
    return libc.syscall(SYS_set_mempolicy, 0, None, 0)


def alloc_pinned(nbytes, cpus=None, interleave=False):
    res = {}

    def work():
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)
            except OSError:
                pass
        if interleave:
            set_mempolicy_interleave(True)
        res["t"] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        if interleave:
            set_mempolicy_interleave(False)

    th = threading.Thread(target=work)
    th.start()
    th.join()
    return res["t"]


def run(n, variant, topo):
    devs = list(range(n))
    src, dst, streams = [], [], []
    for d in devs:
        torch.cuda.set_device(d)
        src.append(torch.empty(CHUNK, dtype=torch.uint8, device=f"cuda:{d}"))
        cpus = parse_cpulist(topo[d]["local_cpulist"]) if variant == "local" and not topo[d]["local_cpulist"].startswith("?") else None
        dst.append(alloc_pinned(PER_GPU, cpus=cpus, interleave=(variant == "interleave")))
        streams.append(torch.cuda.Stream(device=d))
    best = 0.0
    for rep in range(3):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for off in range(0, PER_GPU, CHUNK):
            for d in devs:
                with torch.cuda.stream(streams[d]):
                    dst[d][off:off + CHUNK].copy_(src[d], non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        best = max(best, n * PER_GPU / dt * 1e-9)
    del src, dst
    return best


def memcpy_bw(threads, nbytes=2 * GiB):
    a = np.ones(nbytes // 8)
    b = np.empty_like(a)
    b[:] = 0
    per = a.size // threads

    def work(i):
        np.copyto(b[i * per:(i + 1) * per], a[i * per:(i + 1) * per])

    best = 0.0
    for _ in range(3):
        th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        best = max(best, nbytes / (time.perf_counter() - t0) * 1e-9)
    return best


def main():
    n_gpu = torch.cuda.device_count()
    info = {"n_gpus": n_gpu, "nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    for cmd, key in (("lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA'", "lscpu"), ("nvidia-smi topo -m", "topo"),
                     ("cat /sys/devices/system/node/online", "nodes_online"),
                     ("grep -E 'MemTotal|MemAvailable|Hugepagesize' /proc/meminfo", "meminfo")):
        try:
            info[key] = subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
        except Exception as e:  # noqa: BLE001
            info[key] = str(e)
    topo = [sysfs(d) for d in range(n_gpu)]
    info["devices"] = topo
    info["d2h_gbs"] = {}
    ns = [n for n in (1, 2, 4, 8) if n <= n_gpu]
    multi_node = len(set(t["numa_node"] for t in topo)) > 1 or "-" in info.get("nodes_online", "") or "," in info.get("nodes_online", "")
    variants = ["default"] + (["local", "interleave"] if multi_node else [])
    for v in variants:
        info["d2h_gbs"][v] = {}
        for n in ns:
            info["d2h_gbs"][v][str(n)] = run(n, v, topo)
            print(v, n, "GPUs: %.1f GB/s aggregate" % info["d2h_gbs"][v][str(n)], flush=True)
    # each GPU alone (are all links equal?)
    info["d2h_gbs_each_alone"] = []
    for d in range(n_gpu):
        torch.cuda.set_device(d)
        s = torch.empty(CHUNK, dtype=torch.uint8, device=f"cuda:{d}")
        h = torch.empty(GiB, dtype=torch.uint8, pin_memory=True)
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for off in range(0, GiB, CHUNK):
                h[off:off + CHUNK].copy_(s, non_blocking=True)
            torch.cuda.synchronize(d)
            best = max(best, GiB / (time.perf_counter() - t0) * 1e-9)
        info["d2h_gbs_each_alone"].append(best)
        del s, h
    info["host_memcpy_gbs"] = {str(t): memcpy_bw(t) for t in (1, 2, 4, 8, 16, 32) if t <= (os.cpu_count() or 1)}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/d2h_probe.json", "w") as f:
        json.dump(info, f, indent=1)
    print(json.dumps({k: v for k, v in info.items() if k not in ("topo", "lscpu")}, indent=1))


if __name__ == "__main__":
    main()
