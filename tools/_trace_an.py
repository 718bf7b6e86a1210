import gzip,json,collections,re,sys
d=json.load(gzip.open(sys.argv[1]))
ev=[e for e in d['traceEvents'] if e.get('cat') in ('kernel','gpu_memcpy','gpu_memset')]
ev.sort(key=lambda e:e['ts'])
t0=ev[0]['ts']
def nm(e):
    n=e['name'].replace('(anonymous namespace)::','')
    n=re.sub(r'\(.*','',n).replace('void ','').replace('ab200::','')
    return n[:44]
BIG=('double_layer','i8_gemm','encode_big','row_exp','col_exp')
def isbig(e):
    n=nm(e)
    return any(b in n for b in BIG) or ('dgemm' in n and e['dur']>1000)
pts=[]
for i,e in enumerate(ev):
    pts.append((e['ts']-t0,1,i)); pts.append((e['ts']-t0+e['dur'],0,i))
pts.sort()
act=set(); prev=0; small_only=collections.defaultdict(float); nobig=0; idle=0
segs=[]
for t,k,i in pts:
    dt=t-prev
    if dt>0:
        if not act: idle+=dt
        elif not any(isbig(ev[j]) for j in act):
            nobig+=dt; segs.append((prev,t))
            for j in act: small_only[nm(ev[j])]+=dt/len(act)
    prev=t
    if k: act.add(i)
    else: act.discard(i)
wall=max(e['ts']+e['dur'] for e in ev)-t0
print('wall %.1f idle %.1f no-big-kernel %.1f ms'%(wall/1e3,idle/1e3,nobig/1e3))
for k,v in sorted(small_only.items(), key=lambda x:-x[1])[:10]: print('   %-45s %.1f'%(k,v/1e3))
# histogram of no-big time over the sweep in 10ms bins for first 240 ms
bins=collections.defaultdict(float)
for a,b in segs:
    bins[int(a//10000)]+=b-a
print(' '.join('%d:%.1f'%(k*10,v/1e3) for k,v in sorted(bins.items()) if k<26))
if len(sys.argv)>2:
    strs=sorted(set(e['args'].get('stream') for e in ev))
    for s in strs:
        es=[e for e in ev if e['args'].get('stream')==s and e['ts']-t0<float(sys.argv[2])*1e3]
        out=[]
        for e in es:
            n=nm(e)
            if out and out[-1][0]==n and e['ts']-t0-out[-1][4]<500: out[-1][2]+=e['dur']; out[-1][3]+=1; out[-1][4]=e['ts']+e['dur']-t0
            else: out.append([n,e['ts']-t0,e['dur'],1,e['ts']+e['dur']-t0])
        print('stream',s)
        for o in out:
            if o[2]>500: print(f"  {o[1]/1e3:8.2f} -> {o[4]/1e3:8.2f}  busy {o[2]/1e3:7.2f} ms x{o[3]:4d} {o[0]}")
