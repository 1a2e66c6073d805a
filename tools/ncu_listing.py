"""SASS listing of one profiled launch with per-tile execution counts (development aid).
usage: ncu_listing.py <source-page csv> <tiles> [out]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
tiles = float(sys.argv[2])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ia, ie, iw, ix = col['Source'], col['Instructions Executed'], col['L1 Wavefronts Shared'], col['L1 Wavefronts Shared Excessive']
isamp = col['# Samples']
out = open(sys.argv[3], 'w') if len(sys.argv) > 3 else sys.stdout
tot = totw = 0
for n, r in enumerate(rows[2:]):
    if r and r[0] == 'Kernel Name':
        break
    if len(r) <= ie or not r[ie].isdigit():
        continue
    e = int(r[ie]) / tiles
    tot += e
    totw += int(r[iw]) / tiles
    out.write(f"{n:5d} {e:9.2f} wf {int(r[iw])/tiles:8.2f} ex {int(r[ix])/tiles:7.2f} smp {r[isamp]:>6} {r[ia].strip()[:90]}\n")
out.write(f"total exec/tile {tot:.1f} wavefronts/tile {totw:.1f}\n")
