import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r, e = d["roofline"], d.get("e2e") or {}
    print(f"{f}: value {d['value']:.3e} ({d['ms_per_step']:.1f} ms) frac {r['frac']:.3f} executed {r.get('executed_frac_of_peak', 0):.3f} "
          f"| e2e {e.get('value', 0):.3e} ({e.get('ms_per_step', 0):.1f} ms, min {e.get('ms_per_step_min_rank0', 0):.1f}, "
          f"pcie floor {e.get('pcie_floor_ms') or 0:.1f}) | cpu {(d.get('cpu_baseline') or {}).get('value')} | launches {d['gpu_launches']} | clocks {d['clocks']}")
