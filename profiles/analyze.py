"""Summarise an ncu report exported with
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:k_tile_pass > src.csv
   ncu -i X.ncu-rep --page raw --csv > raw.csv
usage: python profiles/analyze.py src.csv raw.csv [launch_index]"""
import csv, collections, sys
src, raw = sys.argv[1], sys.argv[2]
li = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rows = list(csv.reader(open(src)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_idx[li]; end = hdr_idx[li + 1] - 1 if len(hdr_idx) > li + 1 else len(rows)
hdr = rows[start]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[start + 1:end] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
byop = collections.Counter(); stalls = collections.Counter(); tot = 0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
blocks = []; cur = None
for k, r in enumerate(body):
    s = r[col['Source']].strip().split()
    o = (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
    n = int(r[col['Instructions Executed']]); byop[o] += n; tot += n
    for c in stall_cols: stalls[c] += int(r[col[c]] or 0)
    if cur is None or n != cur['n']:
        cur = {'n': n, 'ops': collections.Counter(), 'first': k}; blocks.append(cur)
    cur['ops'][o] += 1
R = list(csv.reader(open(raw))); H = R[0]; row = R[2 + li]
def g(name):
    return row[H.index(name)] if name in H else None
namps = None
print("launch", li, "duration ms", g('gpu__time_duration.sum'), "regs", g('launch__registers_per_thread'),
      "issue_active%", g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
      "dram%", g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))
print("dram read GB", g('dram__bytes_read.sum'), "write GB", g('dram__bytes_write.sum'),
      "smem bank conflicts", g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'))
print("total warp instr", tot)
for o, n in byop.most_common(14): print(f"  {o:10s} {n:12d} {100 * n / tot:5.1f}%")
ts = sum(stalls.values())
print("stalls:", ", ".join(f"{s[6:]} {100 * n / ts:.1f}%" for s, n in stalls.most_common(8)))
blocks.sort(key=lambda b: -b['n'] * sum(b['ops'].values()))
for b in blocks[:int(sys.argv[4]) if len(sys.argv) > 4 else 12]:
    w = b['n'] * sum(b['ops'].values())
    print(f"  {100 * w / tot:5.1f}%  exec={b['n']:9d} len={sum(b['ops'].values()):4d} idx={b['first']:5d} {dict(b['ops'].most_common(5))}")
