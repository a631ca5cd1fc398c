"""MLPerf DLRM-DCNv2 sample (counterpart of samples/dlrm/train.py): N GPUs on one or several nodes
(scripts/launch_single_node.sh, scripts/slurm_multinode.sub).

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 samples/dlrm_dcnv2_train.py \
      --batchsize 55296 --max_iter 2000 [--source /data/train_data.bin --eval_source /data_val/val_data.bin]
Without --source the synthetic power-law reader is used.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hugectr  # noqa: E402
from hugectr_b200.models.dlrm import (CRITEO_TB_MULTI_HOT, CRITEO_TB_TABLE_SIZES,  # noqa: E402
                                      build_dlrm_dcnv2)
from hugectr_b200.tools.planner import generate_plan  # noqa: E402
from hugectr_b200.utils.mlperf import LoggingCallback  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--batchsize", type=int, default=55296)
p.add_argument("--batchsize_eval", type=int, default=55296)
p.add_argument("--lr", type=float, default=0.004)
p.add_argument("--max_iter", type=int, default=1000)
p.add_argument("--eval_interval", type=int, default=500)
p.add_argument("--display_interval", type=int, default=100)
p.add_argument("--sharding_plan", default="auto", choices=["round_robin", "uniform", "auto"])
p.add_argument("--auc_threshold", type=float, default=0.80275)
p.add_argument("--source", default=None)
p.add_argument("--eval_source", default=None)
p.add_argument("--optimizer", default="adagrad", choices=["adagrad", "sgd"])
p.add_argument("--num_nodes", type=int, default=1)
p.add_argument("--gpus_per_node", type=int, default=0, help="0: WORLD_SIZE / num_nodes")
p.add_argument("--cap_rows", type=int, default=0, help="cap table sizes (functional runs on small machines)")
args = p.parse_args()

num_gpus = int(os.environ.get("WORLD_SIZE", "1"))
gpn = args.gpus_per_node or max(1, num_gpus // args.num_nodes)
sizes = [min(s, args.cap_rows) if args.cap_rows else s for s in CRITEO_TB_TABLE_SIZES]
plan = generate_plan(sizes, CRITEO_TB_MULTI_HOT, num_gpus, plan=args.sharding_plan, num_nodes=args.num_nodes)
multi = dict(gpus_per_node=gpn, comm_strategy=hugectr.CommunicationStrategy.Hierarchical) if args.num_nodes > 1 else {}
cb = LoggingCallback(args.auc_threshold, 4195197692 / args.batchsize, args.batchsize)
model = build_dlrm_dcnv2(batchsize=args.batchsize, num_gpus=num_gpus, lr=args.lr, mixed=True,
                         shard_plan=plan, optimizer=args.optimizer, table_sizes=sizes, **multi,
                         source=[args.source] if args.source else None,
                         batchsize_eval=args.batchsize_eval, training_callbacks=[cb])
model.compile()
model.summary()
model.fit(max_iter=args.max_iter, display=args.display_interval, eval_interval=args.eval_interval,
          snapshot=2000000, snapshot_prefix="dlrm")
