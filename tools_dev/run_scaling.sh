mkdir -p gpurun_out
timeout -k 10 200 python bench.py --steps 30 --warmup 5 --profile gpurun_out/prof1i.txt 2>&1 | grep -E "^\{|Error|error" | cut -c1-260
for n in 2 4; do timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 30 --warmup 5 --profile gpurun_out/prof${n}i.txt 2>&1 | grep -E "^\{|Error|error" | cut -c1-260; done
timeout -k 10 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cd tools_dev; timeout -k 10 200 python bench_gemm.py > ../gpurun_out/bench_gemm_r3.jsonl 2>&1; timeout -k 10 100 python bench_gemm_epi.py > ../gpurun_out/bench_gemm_epi_r3.jsonl 2>&1; cd ..
