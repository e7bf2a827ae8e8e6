import csv, sys, subprocess, collections, re
rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = 0; body = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1; continue
    if k == which and r and r[0].startswith('0x'): body.append(r)
mix = collections.Counter()
for r in body:
    src = r[1].strip()
    src = re.sub(r'^@!?U?P\d+\s+', '', src)
    op = src.split()[0]
    mix[op] += int(r[5])
tot = sum(mix.values())
for op, n in mix.most_common(40):
    print(f"{op:40s} {n:10d} {100*n/tot:5.1f}%")
print('total', tot)
