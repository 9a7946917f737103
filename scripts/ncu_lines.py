#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions of one kernel from an `ncu --set full --import-source on`
report.  ncu's CSV source page is SASS-level; the SASS -> line map comes from `nvdisasm -g` on the cubin extracted from
the shipped library (same build), matched by instruction order.
Usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin file> <function substring> [top_n]"""
import collections
import csv
import io
import re
import subprocess
import sys


def sass_lines(cubin, func):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur_fun, line, res = None, None, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fun = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if cur_fun and func in cur_fun and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            res.append(line)
    return res


def main():
    rep, kre, cubin, func = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first kernel instance only
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] == "Kernel Name":
            break
        body.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    lines = sass_lines(cubin, func)
    if len(lines) != len(body):
        print("warning: %d SASS instructions in the report vs %d in the cubin" % (len(body), len(lines)), file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = tot_i = 0
    for k, r in enumerate(body):
        ln = lines[k] if k < len(lines) else ("?", 0)
        s = int(r[ci["# Samples"]] or 0)
        n = int(r[ci["Instructions Executed"]] or 0)
        a = agg[ln]
        a[0] += s; a[1] += n
        for h in stall_cols:
            v = int(r[ci[h]] or 0)
            if v:
                a[2][h[6:]] += v
        tot_s += s; tot_i += n
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    print("%-22s %8s %6s %10s %6s  top stalls" % ("file:line", "samples", "%", "warp-inst", "%"))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ", ".join("%s %d" % (k, v) for k, v in a[2].most_common(3))
        print("%-22s %8d %6.2f %10d %6.2f  %s" % ("%s:%d" % ln if ln else "?", a[0], 100.0 * a[0] / max(1, tot_s), a[1], 100.0 * a[1] / max(1, tot_i), st))


if __name__ == "__main__":
    main()
