import csv,collections,sys,subprocess,io
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
op=collections.Counter(); samp=collections.Counter(); tot=0; stot=0
stalls=collections.Counter()
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
per=[]
for r in data:
    if len(r)<len(hdr): continue
    s=r[ix['Source']].strip()
    parts=s.split()
    if parts[0].startswith('@'): parts=parts[1:]
    o=parts[0].split('.')[0]
    n=int(r[ix['Instructions Executed']]); op[o]+=n; tot+=n
    sm=int(r[ix['# Samples']]); samp[o]+=sm; stot+=sm
    for c in stall_cols: stalls[c]+=int(r[ix[c]])
    per.append((r[ix['Address']],s,n,sm,{c:int(r[ix[c]]) for c in stall_cols}))
print("total inst",tot, "samples",stot)
for o,n in op.most_common(28): print(f"{o:10s} {n:12d} {100*n/tot:5.1f}%  samples {100*samp[o]/stot:5.1f}%")
print()
for c,n in stalls.most_common(12): print(c,n, f"{100*n/stot:.1f}%")
if len(sys.argv)>2:
    # dump per-instruction table
    with open(sys.argv[2],'w') as f:
        for a,s,n,sm,st in per:
            top=sorted(st.items(),key=lambda x:-x[1])[:2]
            f.write(f"{a[-5:]} {n:10d} {sm:6d} {top[0][0][6:]}:{top[0][1]} {top[1][0][6:]}:{top[1][1]}  {s}\n")
