#!/bin/bash
mkdir -p gpurun_out
HCTR_SYNTH_POOL=8 timeout 70 python bench.py --steps 5 --warmup 3 --no-secondary --sustained-sec 0.2 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,dict) else {a:(b if not isinstance(b,(dict,list)) else '..') for a,b in v.items()}) for k,v in d.items() if k!='config'})"
grep "bench \|Error\|Traceback" gpurun_out/j_bench.err | tail -14
