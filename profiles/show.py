"""Print the key fields of bench JSON lines (helper for reading gpurun_out/*.json)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    r = d.get("roofline", {})
    print(f"{f}: {d['config']['workload']} N={d['n_gpus']} value={d['value'] / 1e9:.3f} G{d['unit']} "
          f"ms/step={d['ms_per_step']:.1f} e2e={d['e2e']['value'] / 1e9:.3f}G ({d['e2e'].get('ms_per_step', 0):.1f} ms) "
          f"roofline={r.get('achieved') and round(r['achieved'])} GB/s frac={r.get('frac') and round(r['frac'], 3)}")
    print("   phases:", {k: round(v, 1) for k, v in d.get("phase_ms", {}).items()}, "launches", d.get("gpu_launches"),
          "clocks", d.get("clocks"))
    if "cpu_baseline" in d:
        print("   cpu:", d["cpu_baseline"])
