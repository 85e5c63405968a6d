"""Summarise ncu outputs into small text files for profiles/:
   python tools/ncu_summary.py launches gpurun_out/launches.csv
   python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep
   python tools/ncu_summary.py source gpurun_out/prof.ncu-rep"""
import collections, csv, io, re, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'launches':>8} {'total ms':>12} {'share':>8}  kernel   (ncu gpu__time_duration.sum, --clock-control none; cold-cache/serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:8d} {v[1]:12.3f} {100*v[1]/tot:7.2f}%  {k}")

KEYS = r"(^Kernel Name|gpu__time_duration.sum$|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|launch__registers_per_thread$|launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|launch__occupancy_limit|sm__warps_active.avg.pct|subpipe_dmma.*pct_of_peak_sustained_active|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|bank_conflicts_pipe_lsu_mem_shared_op_ld.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum|smsp__average_warps_issue_stalled.*_per_issue_active|smsp__issue_active.avg.pct|smsp__inst_executed.sum$|sm__cycles_elapsed.avg$|sm__cycles_elapsed.avg.per_second|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct|lts__t_bytes.sum$|l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum$)"

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for vals in r[2:]:
        for h, u, v in zip(hdr, units, vals):
            if re.search(KEYS, h):
                print(f"{h:95s} {v} {u}")
        print()

def source(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    def g(r, k):
        try: return float(r[idx[k]] or 0)
        except Exception: return 0.0
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(g(r, "# Samples") for r in data)
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in data:
        m = re.match(r"(@!?U?P\d\s+)?([A-Z0-9_]+)", r[idx["Source"]].strip())
        op = m.group(2) if m else "?"
        agg[op]["n"] += g(r, "# Samples")
        for k in reasons: agg[op][k] += g(r, k)
    print(rows[0][1]); print(f"total warp samples {tot:.0f}; share by opcode, with stall reasons > 0.3 % of all samples")
    for op, d in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:16]:
        rs = " ".join(f"{k[6:]}={d[k]/tot*100:.1f}" for k in reasons if d[k]/tot > 0.003)
        print(f"  {op:10s} {d['n']/tot*100:5.1f}%  {rs}")
    print(f"top {top} instructions by samples")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:top]:
        rs = " ".join(f"{k[6:]}={int(g(r,k))}" for k in reasons if g(r, k) > 0)
        print(f"  {int(g(r,'# Samples')):7d} {r[idx['Source']].strip()[:58]:58s} {rs}")

if __name__ == "__main__":
    {"launches": launches, "raw": raw, "source": source}[sys.argv[1]](sys.argv[2])
