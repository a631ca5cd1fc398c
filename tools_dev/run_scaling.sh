mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_embedding_ops_gpu.py tests/test_model_gpu.py tests/test_dist.py -m gpu -x -q 2>&1 | tail -5
timeout -k 10 200 python bench.py --steps 30 --warmup 5 --profile gpurun_out/prof1e.txt 2>&1 | grep -E "^\{|Error|error" | cut -c1-260
for n in 2 4; do timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 30 --warmup 5 --profile gpurun_out/prof${n}e.txt 2>&1 | grep -E "^\{|Error|error" | cut -c1-260; done
