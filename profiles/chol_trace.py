"""Summarise a CMOE_CHOL_TRACE dump: per panel, the chain CTAs' step times (ns)."""
import sys
import re
import collections

rows = collections.defaultdict(dict)
for line in open(sys.argv[1]):
    m = re.match(r"p(\d+) c(\d+) last=(\d+) :(.*)", line)
    if not m:
        print(line.strip())
        continue
    p, c, last = int(m.group(1)), int(m.group(2)), int(m.group(3))
    st = {int(k): int(v.split("[")[0]) for k, v in (tok.split(":") for tok in m.group(4).split())}
    rows[p][c] = (last, st)
panels = [int(a) for a in sys.argv[2:]] or sorted(rows)[:3]
for p in panels:
    print(f"== panel {p}: CTAs 0..5 (slot:ns since the CTA started) ==")
    for c in range(6):
        if c in rows[p]:
            last, st = rows[p][c]
            print(f" c{c} last={last} " + " ".join(f"{k}:{v / 1e3:.1f}us" for k, v in sorted(st.items())))
    ends = [max(st.values()) for last, st in rows[p].values() if st]
    print(f" kernel span (max over CTAs): {max(ends) / 1e3:.1f} us; CTAs traced: {len(rows[p])}")
