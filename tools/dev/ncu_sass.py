import csv, sys, subprocess, re
rep, which, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = 0; body = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1; continue
    if k == which and r and r[0].startswith('0x'): body.append(r)
for i, r in enumerate(body[lo:hi]):
    print(lo + i, r[0][-5:], r[5].rjust(7), r[2].rjust(4), r[1].strip()[:100])
