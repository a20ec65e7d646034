"""Warp instructions executed per CUDA source line of one ncu capture made with --import-source on and -lineinfo:
    python tools/ncu_lines.py file.ncu-rep [file-name-filter] [top]
(the interleaved source page: `ncu -i file --page source --csv --print-source cuda,sass`; rows with a line number carry the
totals of the SASS that line produced)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            rows.append((cur, int(r[0]), r[1].strip(), int(d["Instructions Executed"]), int(d["# Samples"] or 0)))
        except ValueError:
            pass
tot = sum(x[3] for x in rows)
print(f"# {rep}: {tot / 1e6:.1f} M warp instructions over {len(rows)} source lines")
by_file = {}
for f, _, _, n, _ in rows:
    by_file[f] = by_file.get(f, 0) + n
for f, n in sorted(by_file.items(), key=lambda kv: -kv[1]):
    print(f"{n / 1e6:9.1f} M  {100 * n / tot:5.1f} %  {f}")
print()
sel = [x for x in rows if flt in (x[0] or "")]
for f, ln, src, n, smp in sorted(sel, key=lambda x: -x[3])[:top]:
    print(f"{n / 1e6:8.2f} M {100 * n / tot:5.1f} %  samples {smp:6d}  {f.split('/')[-1]}:{ln:<5d} {src[:110]}")
