import re, collections, csv, sys
srccsv, func = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)>10 and r[0].startswith('0x')]
ci=hdr.index('Instructions Executed'); cs=hdr.index('# Samples'); ct=hdr.index('Avg. Threads Executed')
a0=int(data[0][0],16)
prof=[(int(r[0],16)-a0,int(r[ci]),int(r[cs]),float(r[ct]),r[1].strip()) for r in data]
dis=open('/tmp/inflate2.disasm').read().splitlines()
start=[i for i,l in enumerate(dis) if l.startswith('_ZN') and func in l and l.rstrip().endswith(':')][0]
cur=None; off2line={}
for l in dis[start+1:]:
    if l.startswith('//-----') : break
    m=re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l)
    if m: off2line[int(m.group(1),16)]=cur
agg=collections.Counter(); samp=collections.Counter(); tot=0; tots=0
for o,n,s,t,txt in prof:
    k=off2line.get(o); agg[k]+=n; samp[k]+=s; tot+=n; tots+=s
print('total instr',tot,'samples',tots)
byfile=collections.Counter()
for k,v in agg.items(): byfile[k[0] if k else None]+=v
print({k:round(v/tot*100,1) for k,v in byfile.items()})
src={}
for fn in ('inflate2_core.cuh','inflate3_core.cuh','inflate_core.cuh','inflate2.cu'):
    src[fn]=open('/root/repo/wgbs_tools_b200/csrc/'+fn).read().splitlines()
for k,v in agg.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 40):
    if not k: continue
    f,ln=k
    text=src[f][ln-1].strip()[:90] if f in src and ln-1 < len(src[f]) else ''
    print(f"{v/tot*100:5.1f}% {samp[k]/tots*100:5.1f}% {f}:{ln}  {text}")
