"""Basic-block view of an `ncu --page source --csv --print-source sass` dump: contiguous runs of instructions with the
same executed count, with their share of executed warp instructions and of stall samples.
usage: python tools/ncu_blocks.py file.source.csv [kernel-index] [min-share-%]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
# split per kernel
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
k = kernels[kidx]
ix = {h: i for i, h in enumerate(k["hdr"])}
def f(r, key):
    try: return float(r[ix[key]])
    except Exception: return 0.0
data = k["data"]
totE = sum(f(r, "Instructions Executed") for r in data)
totS = sum(f(r, "# Samples") for r in data)
print(k["name"][:110], "instr", len(data), "exec %.2fM" % (totE / 1e6), "samples", int(totS))
stallkeys = [h for h in k["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
blocks, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or f(data[i], "Instructions Executed") != f(data[start], "Instructions Executed"):
        blocks.append((start, i)); start = i
for a, b in blocks:
    E = sum(f(r, "Instructions Executed") for r in data[a:b]); S = sum(f(r, "# Samples") for r in data[a:b])
    if E / totE * 100 < minshare and S / totS * 100 < minshare: continue
    ops = {}
    for r in data[a:b]:
        parts = r[ix["Source"]].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    st = sorted(((sk, sum(f(r, sk) for r in data[a:b])) for sk in stallkeys), key=lambda t: -t[1])[:4]
    top = ", ".join(f"{o}:{n}" for o, n in sorted(ops.items(), key=lambda t: -t[1])[:8])
    print(f"[{a:5d},{b:5d}) n={b-a:4d} exec/instr={f(data[a],'Instructions Executed')/1e3:9.1f}K  exec {E/totE*100:5.1f}%  samples {S/totS*100:5.1f}%  | {top} | " +
          ", ".join(f"{sk[6:]}={v/totS*100:.1f}%" for sk, v in st))
