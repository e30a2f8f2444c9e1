"""Digest of one `ncu --set full --import-source on` report (read on the CPU box): key raw metrics, stall-sample totals,
samples per 100-instruction window (= per warp role in the warp-specialised kernels), the hottest SASS instructions, and
optionally the stall mix and executed-instruction count of one instruction range.
    python tools/ncu_digest.py report.ncu-rep [n_top [first_instr last_instr]]"""
import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.per_cycle_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__cycles_elapsed.max','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__warps_eligible.avg.per_cycle_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.avg.per_cycle_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__shared_mem_per_block_dynamic']
for a,b in zip(h,v):
    if a in want: print(a,b)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1]; data=rows[2:]
isrc=h.index('Source'); isamp=h.index('# Samples'); iex=h.index('Instructions Executed')
stalls=[i for i,n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot=sum(int(r[isamp] or 0) for r in data)
print("total samples",tot, "instrs", len(data))
acc=0;start=0
for i,r in enumerate(data):
    if i%100==0 and i>0:
        print(f"{start}-{i}: {acc}",end='; '); acc=0; start=i
    acc+=int(r[isamp] or 0)
print(f"{start}-end: {acc}")
allst={}
for r in data:
    for j in stalls: allst[h[j]]=allst.get(h[j],0)+int(r[j] or 0)
print(sorted(((v,k) for k,v in allst.items() if v),reverse=True)[:10])
top=sorted(range(len(data)), key=lambda i:-int(data[i][isamp] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
for i in sorted(top):
    r=data[i]
    st=sorted(((int(r[j] or 0),h[j]) for j in stalls),reverse=True)[:2]
    print(i, r[isamp], r[iex], r[isrc][:80], st)
if len(sys.argv)>4:
    a,b=int(sys.argv[3]),int(sys.argv[4])
    t={}
    for r in data[a:b]:
        for j in stalls: t[h[j]]=t.get(h[j],0)+int(r[j] or 0)
    print("region",a,b,sum(t.values()),sorted(((v,k) for k,v in t.items() if v),reverse=True)[:10])
    ex=sum(int(r[iex] or 0) for r in data[a:b])
    print("warp-instructions executed in region", ex)
