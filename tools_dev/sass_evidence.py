"""cuobjdump -sass hugectr_b200/lib/libhctr_cuda.so | python tools_dev/sass_evidence.py > profiles/r2/sass_evidence_r2.txt"""
import collections
import re
import sys

cur = None
cnt = collections.defaultdict(collections.Counter)
pat = re.compile(r'\b(UTCQMMA|UTCHMMA|UTCCP|UTCBAR|LDTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|UTMAPF)[A-Z0-9_.]*')
for l in sys.stdin:
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for x in pat.findall(l):
            cnt[cur][x] += 1
print("cuobjdump -sass hugectr_b200/lib/libhctr_cuda.so : tensor-core / TMEM / TMA / bulk-copy mnemonic counts per kernel")
print("UTCQMMA = tcgen05.mma block-scaled (kind::mxf8f6f4.block_scale)   UTCHMMA = tcgen05.mma kind::f16")
print("UTCCP = tcgen05.cp smem->TMEM   LDTM = tcgen05.ld   UTCBAR = tcgen05.commit   UTMALDG / UTMASTG = TMA tensor load / store")
print("UBLKCP = cp.async.bulk (1-D)   SYNCS = mbarrier ops\n")
for k, v in sorted(cnt.items()):
    if v:
        print(k[:150])
        print("    " + ", ".join(f"{a}:{b}" for a, b in sorted(v.items())))
