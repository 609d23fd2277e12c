"""Opcode mix and hottest SASS lines of one kernel from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
ix = {n: h.index(n) for n in ('Source', '# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_math', 'stall_wait',
                               'stall_short_sb', 'stall_not_selected', 'stall_dispatch', 'stall_branch_resolving')}
body = [r for r in rows[hi + 1:] if len(r) > ix['stall_wait']]
tot_s = sum(int(r[ix['# Samples']] or 0) for r in body)
tot_i = sum(int(r[ix['Instructions Executed']] or 0) for r in body)
ops = collections.Counter(); ops_s = collections.Counter()
for r in body:
    toks = r[ix['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    op = op.split('.')[0]
    ops[op] += int(r[ix['Instructions Executed']] or 0)
    ops_s[op] += int(r[ix['# Samples']] or 0)
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("opcode            inst%   samples%")
for op, n in ops.most_common(22):
    print(f"{op:16s} {100*n/tot_i:6.2f}  {100*ops_s[op]/tot_s:6.2f}")
print("\nhottest lines (by samples): idx samples% | long_sb math wait short_sb not_sel dispatch | source")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:top]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {100*int(r[ix['# Samples']])/tot_s:5.2f} | " + " ".join(f"{r[ix[k]]:>5s}" for k in
          ('stall_long_sb', 'stall_math', 'stall_wait', 'stall_short_sb', 'stall_not_selected', 'stall_dispatch')) + " | " + r[ix['Source']].strip())
