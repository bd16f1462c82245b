"""Per-SOURCE-LINE summary of an ncu report captured with --import-source (kernels built with -lineinfo): share of the
warp-stall samples and of the executed instructions per line, with the dominant stall reason.

    python tools/ncu_lines.py report.ncu-rep <kernel regex> [launch index] [top N] > profiles/xyz_lines.md
"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + pat, "--launch-skip", skip, "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
cur, hdr, agg, fn = None, None, {}, ""
STALLS = ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_math", "stall_mio", "stall_lg",
          "stall_not_selected")
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit() and r[2] == "-":     # the per-line summary rows
        a = agg.setdefault((cur, int(r[0])), [0.0, 0.0, r[1].strip(), {}])
        a[0] += float(r[6] or 0)
        a[1] += float(r[7] or 0)
        for name in STALLS:
            a[3][name] = a[3].get(name, 0.0) + float(r[hdr.index(name)] or 0)
tot = sum(v[0] for v in agg.values()) or 1.0
toti = sum(v[1] for v in agg.values()) or 1.0
print("# per-line profile of `%s`\n\n`%s`\n" % (rep, fn[:160]))
print("total stall samples %d, warp instructions %d\n" % (tot, toti))
print("| file:line | samples | instructions | top stall | source |\n|---|---|---|---|---|")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = max(v[3].items(), key=lambda kv: kv[1])[0]
    print("| %s:%d | %.1f %% | %.1f %% | %s | `%s` |" % (f, l, 100 * v[0] / tot, 100 * v[1] / toti, st, v[2][:100].replace("|", "\\|")))
