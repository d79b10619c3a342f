import csv, sys
fn = sys.argv[1]
rows = list(csv.reader(open(fn)))
hdr = rows[1]; data = rows[2:]
ix = {h:i for i,h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instrs", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# overall by reason
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print({k: round(100*v/tot,1) for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v})
# segments: print running sample count per 64 instrs with top opcodes
lo = int(sys.argv[2]) if len(sys.argv)>2 else 0
hi = min(int(sys.argv[3]), len(data)) if len(sys.argv)>3 else len(data)
thr = float(sys.argv[4]) if len(sys.argv)>4 else 0.4
for i in range(lo, hi):
    r = data[i]; n = int(r[ix["# Samples"]])
    if 100*n/tot >= thr:
        top = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:3]
        print(i, f"{100*n/tot:5.2f}%", r[ix["Source"]].strip()[:70], "exec", r[ix["Instructions Executed"]], [(s[6:], v) for v, s in top if v])
