"""Top stall-sample lines of `ncu --page source --csv` output (SASS or CUDA-C view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
for h in hi:
    hdr = rows[h]
    si, k = hdr.index('Source'), hdr.index('# Samples')
    data = []
    for r in rows[h + 1:]:
        if len(r) <= k or r == hdr or 'Source' in r:
            break
        try:
            data.append((float(r[k] or 0), r[si]))
        except ValueError:
            break
    tot = sum(d[0] for d in data) or 1
    print(f"== block at row {h}: {len(data)} lines, {tot:.0f} samples")
    for v, s in sorted(data, key=lambda d: -d[0])[:n]:
        print(f"{v / tot * 100:5.1f}%  {s[:120]}")
