#!/usr/bin/env python
"""Aggregates an `ncu --set full --import-source on` report per CUDA source line.

    python tools/ncu_lines.py report.ncu-rep [top_n]

Prints, per source line: warp instructions executed, average active threads per instruction, stall samples
and the dominant stall reasons — the evidence DESIGN.md quotes for divergence and latency."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    agg = collections.OrderedDict()
    src = {}
    line_key = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {n: i for i, n in enumerate(hdr)}
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0] != "":                       # a CUDA source line (summary row)
            line_key = (cur_file, int(r[0]))
            src[line_key] = r[1]
            continue
        # SASS row belonging to line_key
        a = agg.setdefault(line_key, collections.Counter())
        def g(name):
            try:
                return float(r[ix[name]])
            except Exception:
                return 0.0
        a["inst"] += g("Instructions Executed")
        a["tinst"] += g("Thread Instructions Executed")
        a["samples"] += g("# Samples")
        for n in ix:
            if n.startswith("stall_") and "Not Issued" not in n:
                a[n] += g(n)
    tot_inst = sum(a["inst"] for a in agg.values())
    tot_samp = sum(a["samples"] for a in agg.values())
    tot_tinst = sum(a["tinst"] for a in agg.values())
    print(f"total warp inst {tot_inst:.3e}  thread inst {tot_tinst:.3e}  avg threads {tot_tinst / max(tot_inst, 1):.2f}  samples {tot_samp:.0f}")
    stall_tot = collections.Counter()
    for a in agg.values():
        for k, v in a.items():
            if k.startswith("stall_"):
                stall_tot[k] += v
    print("stalls:", ", ".join(f"{k[6:]} {v / max(tot_samp, 1) * 100:.1f}%" for k, v in stall_tot.most_common(8)))
    print(f"{'file:line':28s} {'inst%':>6s} {'thr':>5s} {'samp%':>6s}  top stalls | source")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_") and v > 0), key=lambda kv: -kv[1])[:2]
        sts = " ".join(f"{k}:{v / max(a['samples'], 1) * 100:.0f}%" for k, v in st)
        print(f"{key[0] + ':' + str(key[1]):28s} {a['inst'] / tot_inst * 100:6.2f} {a['tinst'] / max(a['inst'], 1):5.1f} "
              f"{a['samples'] / max(tot_samp, 1) * 100:6.2f}  {sts:28s} | {src.get(key, '').strip()[:90]}")


if __name__ == "__main__":
    main()
