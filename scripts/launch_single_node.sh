#!/usr/bin/env bash
# One process per GPU on one NVSwitch box (the role of samples/dlrm/run_and_time.sh).
#   scripts/launch_single_node.sh 8 samples/dlrm_dcnv2_train.py --batchsize 55296
set -euo pipefail
NGPU=${1:?number of GPUs}; shift
export NCCL_NVLS_ENABLE=${NCCL_NVLS_ENABLE:-1}
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "${NGPU}" \
     --master-addr 127.0.0.1 --master-port "${MASTER_PORT:-29500}" "$@"
