"""Warp-stall samples of an .ncu-rep aggregated per CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_hotspots.py report.ncu-rep [top-n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    cur_file = cur_fn = None
    hdr = None
    acc = {}   # function -> {(file, line): [samples, instr, text]}
    line_key = None
    for r in rows:
        if r[0] == "File Path":
            cur_file = r[1]; continue
        if r[0] == "Function Name":
            cur_fn = r[1]; continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}; continue
        if hdr is None or cur_fn is None:
            continue
        si = [i for n, i in hdr.items() if n.startswith("Warp Stall Sampling (All")][0]
        ii = hdr.get("Instructions Executed")
        if r[0] not in ("", "..."):                     # a CUDA source line; its own counters are the sum over its SASS
            line_key = (cur_file.split("/")[-1], r[0])
            try:
                s = float(r[si] or 0); n = float(r[ii] or 0)
            except ValueError:
                s = n = 0.0
            d = acc.setdefault(cur_fn, {})
            e = d.setdefault(line_key, [0.0, 0.0, r[1].strip()])
            e[0] += s; e[1] += n
    for fn, d in acc.items():
        tot = sum(v[0] for v in d.values()) or 1.0
        toti = sum(v[1] for v in d.values()) or 1.0
        print("\n%s\n  total samples %d, total warp instructions %d" % (fn[:110], tot, toti))
        for (f, ln), (s, n, text) in sorted(d.items(), key=lambda kv: -kv[1][0])[:topn]:
            print("  %5.1f%% of samples %5.1f%% of instructions  %s:%s  %s" % (100 * s / tot, 100 * n / toti, f, ln, text[:110]))


if __name__ == "__main__":
    main()
