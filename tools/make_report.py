#!/usr/bin/env python
"""Markdown tables for DESIGN.md from the bench / sweep JSON lines of a round.
    python tools/make_report.py profiles r02"""
import json
import os
import sys

d, tag = sys.argv[1], sys.argv[2]


def load(name):
    p = os.path.join(d, name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return [json.loads(l) for l in lines]


print("### bench.py (cfg2: 1M points per GPU, k = 16, 200*sqrt(N) slices)\n")
print("| GPUs | device ms/step | points/s | exchange ms | e2e ms/step | e2e points/s | CPU arm points/s (threads) | parity vs 1 GPU |")
print("|---|---|---|---|---|---|---|---|")
base = None
for n in (1, 2, 4, 8):
    b = load("%s_bench_%dgpu.json" % (tag, n))
    r = load("%s_bench_%dgpu_ref.json" % (tag, n))
    if not b:
        continue
    b = b[-1]
    ref = ("%.2f M (%d)" % (r[-1]["value"] / 1e6, r[-1]["cpu_baseline"]["cores"])) if r else "-"
    par = b.get("multi_gpu_parity")
    print("| %d | %.3f | %.2f G | %s | %.2f | %.2f G | %s | %s |" % (
        n, b["ms_per_step"], b["value"] / 1e9, "%.3f" % b["exchange_ms_per_step"] if "exchange_ms_per_step" in b else "-",
        b["e2e"]["ms_per_step"], b["e2e"]["value"] / 1e9, ref,
        ("%d rows / %d contours differ" % (par["normal_rows_differing"], par["contours_differ"])) if par else "-"))
    c3 = b.get("north_star_cfg3")
    if c3:
        print("|   | cfg3 on %d GPU(s): device %.2f ms, e2e %.2f ms (target 50 ms) | | %s | | | | |" % (
            n, c3["device_ms_per_step"], c3["e2e_ms_per_step"], ("%.3f" % c3["exchange_ms_per_step"]) if c3.get("exchange_ms_per_step") else "-"))
print()
for n in (1, 2, 4, 8):
    b = load("%s_bench_%dgpu.json" % (tag, n))
    if b and b[-1].get("roofline"):
        r = b[-1]["roofline"]
        print("N=%d kernels (ms/step, serialised): %s" % (n, ", ".join("%s %.3f" % kv for kv in list(r["kernel_ms_per_step"].items())[:14])))
        print("N=%d roofline: %s %.1f GB/s of %.0f = %.3f\n" % (n, r["kernel"], r["achieved"], r["peak"], r["frac"]))
print("\n### sweeps (device-resident ms/step, exchange included for N > 1)\n")
rows = {}
for n in (1, 2, 4, 8):
    for e in load("%s_sweep_%dgpu.jsonl" % (tag, n)) or []:
        rows.setdefault(e["name"], {})[n] = e
print("| point | " + " | ".join("%d GPU" % n for n in (1, 2, 4, 8)) + " |")
print("|---|---|---|---|---|")
for name, by in rows.items():
    print("| %s | " % name + " | ".join(("%.2f" % by[n]["ms_per_step"]) if n in by else "-" for n in (1, 2, 4, 8)) + " |")
