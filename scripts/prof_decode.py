import re,collections,sys
names={0:'start',1:'A.resolve',2:'A.ln',3:'A.qkv',4:'A.attn',5:'A.wo',6:'A.store',7:'A.end',8:'A.bar',
11:'B.resolve',12:'B.ln',13:'B.qc',14:'B.attn',15:'B.woc',16:'B.store',17:'B.end',18:'B.bar',
21:'C.resolve',22:'C.ln',23:'C.fc1',25:'C.fc2',26:'C.store',27:'C.end',28:'C.bar',31:'F.end',32:'F.bar',33:'G.end',35:'G.x',40:'g.chunk',42:'g.acq',41:'g.sync1',36:'G.mma',37:'G.acc',38:'G.epi',34:'G.bar'}
for line in open(sys.argv[1]):
    if not line.startswith('PROF'): continue
    cta=line.split(':')[0]
    if cta not in ('PROF cta 0','PROF cta 127','PROF cta 147'): continue
    ev=[(int(a),int(b)) for a,b in re.findall(r' (\d+):(\d+)',line)]
    agg=collections.OrderedDict(); prev=0
    for tag,t in ev:
        if tag==0: prev=t; continue
        agg.setdefault(names.get(tag,tag),[]).append(t-prev); prev=t
    print(cta,'total',ev[-1][1]/1000,'us')
    print('   '+'  '.join(f"{k}={sum(v)/len(v)/1000:.1f}x{len(v)}" for k,v in agg.items()))
