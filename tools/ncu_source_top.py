"""Digest of an `ncu --page source --csv` dump: stall-reason totals, executed warp-instructions by opcode, and the
hottest SASS instructions.  usage: python tools/ncu_source_top.py src.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(h)]
col = {n: i for i, n in enumerate(h)}
def num(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
tot_s = sum(num(r, "# Samples") for r in body); tot_i = sum(num(r, "Instructions Executed") for r in body)
print(rows[0][1][:110]); print("SASS instructions", len(body), "samples", int(tot_s), "warp-instructions executed", int(tot_i))
st = {n: sum(num(r, n) for r in body) for n in h if n.startswith("stall_") and not n.endswith("_not_issued")}
print("stalls:", ", ".join(f"{k[6:]} {100 * v / max(sum(st.values()), 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
ops = collections.Counter(); ops_s = collections.Counter()
for r in body:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"): op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] += num(r, "Instructions Executed"); ops_s[op] += num(r, "# Samples")
print("by opcode (share of executed | share of samples):", ", ".join(f"{k} {100 * v / tot_i:.0f}|{100 * ops_s[k] / max(tot_s, 1):.0f}" for k, v in ops.most_common(14)))
for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:top_n]:
    print(f"  {num(r, '# Samples') / max(tot_s, 1) * 100:5.1f}%  exec {int(num(r, 'Instructions Executed')):>9}  {r[col['Source']][:90]}")
