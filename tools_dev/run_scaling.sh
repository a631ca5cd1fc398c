mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_embedding_ops_gpu.py tests/test_model_gpu.py tests/test_dist.py -m gpu -x -q 2>&1 | tail -5
timeout -k 10 120 python tools_dev/bench_ew.py > gpurun_out/bench_ew.jsonl 2>&1; tail -3 gpurun_out/bench_ew.jsonl
timeout -k 10 200 python bench.py --steps 30 --warmup 5 --profile gpurun_out/prof1d.txt 2>&1 | grep -E "^\{|Error|error" | cut -c1-260
for n in 2 4; do timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 30 --warmup 5 --profile gpurun_out/prof${n}d.txt 2>&1 | grep -E "^\{|Error|error" | cut -c1-260; done
HCTR_PRIO_MAIN=-1 timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 30 --warmup 5 2>&1 | grep -E "^\{|Error|error" | cut -c1-260
HCTR_PRIO_EMB=0 timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 30 --warmup 5 2>&1 | grep -E "^\{|Error|error" | cut -c1-260
