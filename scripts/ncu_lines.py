"""Per-source-line hot spots from an ncu report:
   python scripts/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, H, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        H = r
    elif H and r[0] not in ("", "Function Name") and len(r) > 10:
        try:
            lines.append((fname, int(r[0]), r[1].strip()[:90], int(r[H.index("# Samples")]),
                          int(r[H.index("Instructions Executed")])))
        except ValueError:
            pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s}, total warp-instructions {tot_i}")
for f, ln, src, s, i in sorted(lines, key=lambda l: -(l[4] if len(sys.argv) > 4 else l[3]))[:top]:
    print(f"{100 * s / tot_s:5.1f}% smp {100 * i / tot_i:5.1f}% ins  {f}:{ln:<4d} {src}")
