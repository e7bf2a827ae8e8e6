import csv, sys, subprocess
rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = 0; hdr = None; body = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1; continue
    if k == which:
        if r and r[0] == 'Address': hdr = r; continue
        if r and r[0].startswith('0x'): body.append(r)
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[2]) for r in body)
print('total samples', tot, 'instructions', len(body), 'executed warp-instr', sum(int(r[5]) for r in body))
agg = {h: sum(int(r[i]) for r in body) for i, h in reasons}
print({h: v for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(body, key=lambda r: -int(r[2]))[:topn]:
    rs = sorted(((int(r[i]), h[6:]) for i, h in reasons if int(r[i])), reverse=True)[:3]
    print(r[0][-5:], r[2].rjust(5), r[5].rjust(7), r[1].strip()[:70].ljust(70), rs)
