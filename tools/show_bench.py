#!/usr/bin/env python
"""One-line digest of a bench.py JSON line: tools/show_bench.py gpurun_out/bench.json"""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e); continue
    r = d.get("roofline") or {}
    print(f"{path}: value={d['value']:.2f} {d['unit']} ms/step={d['ms_per_step']:.4f} kernel_ms={r.get('kernel_ms', 0):.4f} "
          f"frac={r.get('frac', 0):.3f} e2e={(d['e2e'].get('value') or 0):.2f} kernel={str(r.get('kernel'))[:28]} loss={(d.get('run') or d['config']).get('loss')}")
