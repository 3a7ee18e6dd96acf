import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
ks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]}; ks.append(cur); continue
    if r and r[0]=='Address': cur['hdr']=r; continue
    if cur is not None and r: cur['rows'].append(r)
for k in ks:
    h=k['hdr']; iS=h.index('Source'); iE=h.index('Instructions Executed'); iSt=h.index('Warp Stall Sampling (All Samples)')
    tot=sum(int(r[iE]) for r in k['rows'])
    print(k['name'][:60], 'total warp instr', tot, 'n sass', len(k['rows']))
    op=collections.Counter(); st=collections.Counter()
    for r in k['rows']:
        t=r[iS].split()
        o=t[1] if t[0].startswith('@') else t[0]
        op[o]+=int(r[iE]); st[o]+=int(r[iSt])
    for o,c in op.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 14): print(f"   {o:20s} {c:9d} {c/tot:.3f}  stalls {st[o]}")
