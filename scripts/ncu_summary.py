"""Digest of an `ncu --set full --import-source on` report for one kernel (raw page + source page).
    python scripts/ncu_summary.py REPORT.ncu-rep KERNEL_REGEX [--json profiles/r02_<kernel>_counters.json --pairs P --npix N]
With --json the per-pair counters bench.py reports (roofline.traffic, issue_frac, thread_inst_per_px) are written together with a
hash of the kernel sources, so that bench.py can tell when they no longer describe the code it runs."""
import csv, sys, collections, subprocess, json, hashlib, os
rep, kern = sys.argv[1], sys.argv[2]
opt = dict(zip(sys.argv[3::2], sys.argv[4::2]))
vals = {}
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct']
for r in rows[2:]:
    if kern in r[hdr.index('Kernel Name')]:
        for w in want:
            if w in hdr:
                print('  ',w,'=',r[hdr.index(w)])
                vals[w] = (r[hdr.index(w)], rows[1][hdr.index(w)])
        break
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]
data=[]
for r in rows[hi+1:]:                       # one section only: a report with several kernels repeats the table
    if r and r[0]=='Address': break
    if len(r)==len(hdr): data.append(r)
ia=hdr.index('Instructions Executed'); ts=hdr.index('Thread Instructions Executed'); sc=hdr.index('Source'); ss=hdr.index('# Samples')
tot=sum(int(r[ia]) for r in data); tott=sum(int(r[ts]) for r in data); totsamp=sum(int(r[ss]) for r in data)
print('sass instrs',len(data),'warp-instr',tot,'thread-instr',tott,'samples',totsamp)
ops=collections.Counter(); samp=collections.Counter()
for r in data:
    t=r[sc].split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op]+=int(r[ia]); samp[op]+=int(r[ss])
for op,c in ops.most_common(18): print('   %-10s %6.2f%% inst   %6.2f%% samples'%(op,100*c/tot,100*samp[op]/totsamp))
# stall breakdown
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tots={h:sum(int(r[hdr.index(h)] or 0) for r in data) for h in stalls}
print({k:v for k,v in sorted(tots.items(), key=lambda kv:-kv[1])[:10]})
# hottest instructions by samples
top=sorted(data,key=lambda r:-int(r[ss]))[:25]
for r in top: print('   ',r[ss],r[ia],r[sc][:90])

if "--json" in opt:
    ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    def num(k):
        v, unit = vals[k]
        x = float(v.replace(",", ""))
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    pairs, npix = int(opt["--pairs"]), int(opt["--npix"])
    sources = ["new_bloom_filter_repo_b200/csrc/" + f for f in ("rbf_k3_query.cuh", "rbf_hash.cuh", "rbf_k2_insert.cuh", "rbf_kernels.cu")]
    h = hashlib.sha256()
    for name in sources:
        h.update(open(os.path.join(ROOT, name), "rb").read())
    out = {"kernel": kern, "report": os.path.basename(rep), "pairs_in_launch": pairs, "npix": npix,
           "gpu_time_us": num('gpu__time_duration.sum'),
           "dram_bytes_read": num('dram__bytes_read.sum'), "dram_bytes_write": num('dram__bytes_write.sum'),
           "dram_bytes_per_pair": (num('dram__bytes_read.sum') + num('dram__bytes_write.sum')) / pairs,
           "thread_inst_per_px": tott / (pairs * npix), "warp_inst": tot,
           "issue_active_frac": num('smsp__issue_active.avg.pct_of_peak_sustained_active') / 100.0,
           "alu_pipe_frac": num('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active') / 100.0,
           "fma_pipe_frac": num('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active') / 100.0,
           "sources": sources, "sources_sha16": h.hexdigest()[:16]}
    with open(opt["--json"], "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", opt["--json"])
