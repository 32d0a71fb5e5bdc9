"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur_file = None; out = []
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; ismp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); continue
    if r[0] != '': out.append((cur_file, num(r[0]), r[1].strip()[:110], num(r[ismp]), num(r[iex])))
tot_s = sum(o[3] for o in out); tot_e = sum(o[4] for o in out)
print('total samples', tot_s, 'exec', tot_e)
for o in sorted(out, key=lambda x: -x[3])[:top]:
    print(f"{o[0]:16s} {o[1]:4d} smp={o[3]:6d} ({100*o[3]/tot_s:4.1f}%) exec={o[4]/1e6:7.2f}M  {o[2]}")
