"""Print one training step of a bench.py --profile *.timeline dump (start us, dur us, stream, kernel)."""
import sys
path, min_us = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
ev = []
for l in open(path):
    a, b, c, d = l.split(None, 3)
    ev.append((float(a), float(b), c, d.strip()))
starts = [e[0] for e in ev if 'lr_step' in e[3]]
s0, s1 = starts[2], starts[3]
streams = sorted(set(e[2] for e in ev))
print(f"step length {s1 - s0:.1f} us")
for e in ev:
    if s0 <= e[0] < s1 and (e[1] >= min_us or 'barrier' in e[3] or 'allreduce' in e[3]):
        print(f"{e[0]-s0:8.1f} {e[1]:7.1f} {e[0]-s0+e[1]:8.1f} s{streams.index(e[2])} {e[3][:64]}")
