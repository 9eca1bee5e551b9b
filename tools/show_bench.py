#!/usr/bin/env python3
"""Print the headline numbers of bench.py JSON lines (one file per argument)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    r = d.get("roofline") or {}
    st = r.get("stage_ms_per_step", {})
    print(f"{f}: value={d.get('value'):.4g} ms/step={d.get('ms_per_step'):.4g} e2e={((d.get('e2e') or {}).get('value') or 0):.4g} frac={r.get('frac')}")
    print("   ", {k[3:]: round(v, 4) for k, v in st.items()}, "noi", d.get("config", {}).get("mean_interactions"))
