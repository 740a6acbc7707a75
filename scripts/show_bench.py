#!/usr/bin/env python
"""Print the per-regime numbers of bench.py JSON lines (one file per argument)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        txt = [l for l in open(f).read().splitlines() if l.startswith("{")]
        x = json.loads(txt[-1])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
        continue
    for r, res in x.get("regimes", {}).items():
        par = res.get("parity", {})
        print(f"{f.split('/')[-1]:34s} {r:8s} {res['pipeline']:12s} ms {res['ms_per_step']:.4f} it/s {res['value']:8.1f} "
              f"step-roof {res['roofline_step']['frac']:.3f} dom {res['roofline']['kernel']}:{res['roofline']['frac']:.3f} parity {par.get('ok')}")
        print("      ", {k: round(v, 4) for k, v in res["kernel_ms"].items()}, "keys", int(res.get("keys_emitted", 0)),
              "sumchk", (res.get("exchange_sum_check") or {}).get("ok"))
