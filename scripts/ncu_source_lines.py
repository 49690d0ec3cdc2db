"""ncu report -> executed instructions and stall samples per CUDA source line and per file (needs -lineinfo and
--import-source on):  python scripts/ncu_source_lines.py in.ncu-rep out.csv [top]"""
import csv, subprocess, sys, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[3]) if len(sys.argv) > 3 else 80
cur = hdr = None
acc, tot, per_file = [], 0, collections.Counter()
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, iss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and len(r) == len(hdr) and r[0]:
        try:
            n, s = int(r[ie]), int(r[iss])
        except ValueError:
            continue
        acc.append((n, s, cur, r[0], " ".join(r[1].split())[:140]))
        tot += n
        per_file[cur] += n
ts = sum(a[1] for a in acc) or 1
acc.sort(reverse=True)
with open(sys.argv[2], "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["pct_of_executed_instructions", "pct_of_stall_samples", "file", "line", "source"])
    for f, n in per_file.most_common():
        w.writerow([f"{100 * n / tot:.2f}", "", f, "*", "(whole file)"])
    for n, s, f, l, src in acc[:top]:
        w.writerow([f"{100 * n / tot:.2f}", f"{100 * s / ts:.2f}", f, l, src])
