import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
N=int(sys.argv[2]) if len(sys.argv)>2 else 60
cur=None; hdr=None; out=[]; seen_kernel=0
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1]; continue
    if len(r)>=2 and r[0]=='Function Name': continue
    if len(r)>=2 and r[0]=='Line No': hdr=r; continue
    if hdr and cur and r and r[0].isdigit() and len(r)>=12:
        ie=int(r[7]) if r[7].isdigit() else 0
        th=int(r[8]) if r[8].isdigit() else 0
        st=int(r[4]) if r[4].isdigit() else 0
        if ie>0: out.append((ie,cur.split('/')[-1],int(r[0]),r[1].strip()[:100],st,th))
tot=sum(o[0] for o in out); stt=sum(o[4] for o in out)
print('total warp inst',tot,'samples',stt)
byfile=collections.Counter(); sf=collections.Counter()
for o in out: byfile[o[1]]+=o[0]; sf[o[1]]+=o[4]
print({k:round(100*v/tot,1) for k,v in byfile.items()}, {k:round(100*v/stt,1) for k,v in sf.items()})
for o in sorted(out,reverse=True)[:N]:
    print('%5.2f%% st %5.2f%% %s:%d thr=%.1f | %s'%(100*o[0]/tot,100*o[4]/stt,o[1],o[2],o[5]/o[0],o[3]))
