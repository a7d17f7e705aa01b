import csv, sys, subprocess, collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__cycles_elapsed.avg','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','launch__shared_mem_config_size','launch__shared_mem_per_block_static','sm__maximum_warps_per_active_cycle_pct']
for k in keys:
    idx=[i for i,h in enumerate(hdr) if h==k]
    if idx: print(k, [r[idx[0]] for r in rows[2:]], rows[1][idx[0]])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
res=[]; cur=None; h=None; nk=0
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No":
        h=r; ix={c:i for i,c in enumerate(h) if c not in ('Source',)}
        continue
    if r[0]!="" and h:
        try: res.append((cur,int(r[0]),r[1].strip()[:90],float(r[ix['# Samples']]),float(r[ix['Instructions Executed']]),{c:float(r[ix[c]]) for c in ('stall_wait','stall_no_inst','stall_long_sb','stall_short_sb','stall_branch_resolving','stall_selected','stall_math','stall_lg','stall_mio','stall_barrier','stall_dispatch','stall_not_selected')}))
        except Exception as e: pass
agg=collections.OrderedDict()
tot=collections.Counter()
for f,l,s,sm,ins,st in res:
    a=agg.setdefault((f,l),[s,0,0,collections.Counter()]); a[1]+=sm; a[2]+=ins; a[3].update(st); tot.update(st)
ts=sum(a[1] for a in agg.values()); ti=sum(a[2] for a in agg.values())
print("samples",ts,"inst",ti)
print({k:round(100*v/sum(tot.values()),1) for k,v in tot.most_common()})
BY = 2 if "--by-inst" in sys.argv else 1
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][BY])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    top=a[3].most_common(2)
    print("%-13s %4d smp %4.1f%% inst %4.1f%% %-28s| %s"%(k[0][:13],k[1],100*a[1]/ts,100*a[2]/ti,",".join("%s:%d"%(n[6:],v) for n,v in top),a[0]))
