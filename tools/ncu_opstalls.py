"""Per-opcode stall-sample aggregation from `ncu --page source --csv` output (file given as argv[1])."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
keys = ['stall_long_sb', 'stall_wait', 'stall_math', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb',
        'stall_dispatch', 'stall_branch_resolving', 'stall_barrier', 'stall_membar', 'stall_sleep', 'stall_lg', 'stall_mio']
agg = {}
for r in data:
    parts = r[ix['Source']].split()
    op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    a = agg.setdefault(op, {'n': 0, 'exec': 0, **{k: 0 for k in keys}})
    a['n'] += f(r, '# Samples'); a['exec'] += f(r, 'Instructions Executed')
    for k in keys: a[k] += f(r, k)
print("total samples", tot)
for op, a in sorted(agg.items(), key=lambda kv: -kv[1]['n'])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    st = sorted(((k, a[k]) for k in keys), key=lambda t: -t[1])[:3]
    print(f"{op:10s} {a['n']/tot*100:6.1f}% exec {a['exec']/1e6:8.2f}M  " + ", ".join(f"{k[6:]}={v/tot*100:.1f}%" for k, v in st))
