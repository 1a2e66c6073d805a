import csv, sys, subprocess
rep=sys.argv[1]; tiles=400000
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); iw=hdr.index('L1 Wavefronts Shared'); ix=hdr.index('L1 Wavefronts Shared Excessive'); ii=hdr.index('L1 Wavefronts Shared Ideal')
data=[]
for r in rows[2:]:
    if r and r[0]=='Kernel Name': break
    if len(r)>ie and r[ie].isdigit(): data.append(r)
tot=sum(int(r[iw]) for r in data); print('total wavefronts/tile', tot/tiles, 'excessive', sum(int(r[ix]) for r in data)/tiles)
top=sorted(enumerate(data), key=lambda x:-int(x[1][iw]))[:45]
for i,r in sorted(top):
    print(i, r[ia].strip()[:60], 'exec/tile',round(int(r[ie])/tiles,2),'wf/tile',round(int(r[iw])/tiles,1),'ideal',round(int(r[ii])/tiles,1),'excess',round(int(r[ix])/tiles,1))
