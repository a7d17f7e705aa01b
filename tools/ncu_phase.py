import csv, sys, subprocess, collections, re
rep=sys.argv[1]
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
res=[]; cur=None; h=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No":
        h=r; ix={c:i for i,c in enumerate(h) if c not in ('Source',)}; continue
    if r[0]!="" and h:
        try: res.append((cur,int(r[0]),float(r[ix['# Samples']]),float(r[ix['Instructions Executed']])))
        except: pass
# map lines to function names by scanning source files
import os
def func_map(path):
    fm={}; name='?'
    for i,l in enumerate(open(path),1):
        m=re.match(r'^(static )?(__device__|__global__|cudaError_t|template).*?(\w+)\(',l)
        if m and not l.startswith(' '): name=m.group(3)
        fm[i]=name
    return fm
import os
base=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'decision-making-and-path-planning_b200','csrc')+'/'
fms={f:func_map(base+f) for f in ('dp_device.cuh','dp_cycle.cu','dp_fused.cuh')}
agg=collections.Counter(); smp=collections.Counter()
for f,l,s,i in res:
    fn=fms.get(f,{}).get(l,f)
    if f=='dp_cycle.cu' and fn=='dp_cycle_kernel':
        # phase split by line
        txt=open(base+f).read().split('\n')
        fn='kernel:'+('decision' if l< [k for k,t in enumerate(txt,1) if 'Planning thread iteration' in t][0] else 'planning')
    agg[fn]+=i; smp[fn]+=s
ti=sum(agg.values()); ts=sum(smp.values())
for k,v in agg.most_common(25): print("%-28s inst %5.1f%%  samples %5.1f%%"%(k,100*v/ti,100*smp[k]/ts))
