# %% [markdown]
# # Embedding collections: tables, lookups, sharding plans, checkpoints
#
# The `EmbeddingCollectionConfig` API of HugeCTR (the path of the MLPerf DLRM-DCNv2 benchmark) on this framework:
# table configs, grouped lookups, a sharding plan from the planner, training, and a re-shardable checkpoint.
# Executed on CPU in one process; under `torchrun --nproc-per-node 8` the same script shards the tables over 8 GPUs
# and the exchanges run as peer-memory kernels over NVLink.

# %%
import os
import torch
import hugectr
from hugectr_b200.tools.planner import generate_plan
print(hugectr.__version__)

# %% [markdown]
# ## 1. A sharding plan for 8 GPUs
#
# `generate_plan(table_sizes, multi_hot_sizes, num_gpus)` returns `(shard_matrix, shard_strategy)`: small tables are
# data-parallel, the hottest table is row-split over several GPUs, the rest is placed table-wise by a bandwidth cost
# model.  (The same function produces the plan `bench.py` runs.)

# %%
sizes = [400000, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 130229, 3067956, 405282, 10, 2209, 11938, 155, 4, 976,
         14, 292775, 40790948, 187188, 590152, 12973, 108, 36]
hot = [3, 2, 1, 2, 6, 1, 1, 1, 1, 7, 3, 8, 1, 6, 9, 5, 1, 1, 1, 12, 100, 27, 10, 3, 1, 1]
matrix, strategy = generate_plan(sizes, hot, 8)
for kind, names in strategy:
    print(kind, names)
for g, row in enumerate(matrix):
    print(f"GPU {g}: tables", [i for i, x in enumerate(row) if x])

# %% [markdown]
# ## 2. A model with a collection (one process here, so the plan is the trivial one)

# %%
small = [min(s, 5000) for s in sizes[:8]]
solver = hugectr.CreateSolver(batchsize=256, batchsize_eval=256, lr=0.05, vvgpu=[[0]], repeat_dataset=True,
                              max_eval_batches=4)
reader = hugectr.DataReaderParams(hugectr.DataReaderType_t.RawAsync, source=["synthetic:1.1"], eval_source="synthetic:1.1",
                                  check_type=hugectr.Check_t.Non)
model = hugectr.Model(solver, reader, hugectr.CreateOptimizer(hugectr.Optimizer_t.AdaGrad, initial_accu_value=0.1))
model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=13, dense_name="dense",
                        data_reader_sparse_param_array=[hugectr.DataReaderSparseParam(f"data{i}", hot[i], False, 1)
                                                        for i in range(8)]))
tables = [hugectr.EmbeddingTableConfig(name=str(i), max_vocabulary_size=small[i], ev_size=16) for i in range(8)]
ebc = hugectr.EmbeddingCollectionConfig()
ebc.embedding_lookup(table_config=tables, bottom_name=[f"data{i}" for i in range(8)], top_name="emb",
                     combiner=["sum"] * 6 + ["mean", "sum"])
ebc.shard(shard_matrix=[[str(i) for i in range(8)]], shard_strategy=[("mp", [str(i) for i in range(8)])])
model.add(ebc)
model.add(hugectr.DenseLayer(hugectr.Layer_t.MLP, ["dense"], ["bottom"], num_outputs=[64, 16],
                             act_type=hugectr.Activation_t.Relu))
model.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["emb", "bottom"], ["concat"]))
model.add(hugectr.DenseLayer(hugectr.Layer_t.MLP, ["concat"], ["top"], num_outputs=[64, 1],
                             act_type=hugectr.Activation_t.Relu, use_bias=True,
                             activations=[hugectr.Activation_t.Relu, hugectr.Activation_t.Non]))
model.add(hugectr.DenseLayer(hugectr.Layer_t.BinaryCrossEntropyLoss, ["top", "label"], ["loss"]))
model.compile()
model.summary()

# %% [markdown]
# ## 3. Train, then dump the collection (`embedding_collection_0/{meta_data, key*, weight*}`)

# %%
model.fit(max_iter=200, display=100, eval_interval=100, snapshot=1000000)
model.embedding_dump("./ebc_ckpt", [str(i) for i in range(8)])
for root, _, files in os.walk("./ebc_ckpt"):
    print(root, sorted(files)[:6], "..." if len(files) > 6 else "")

# %% [markdown]
# ## 4. Load it back (any sharding, any world size) and check a table

# %%
e = model.ebcs_train[0]
before = {n: [x[1].clone() for x in e.dump_table_local(n)] for n in ("0", "6")}
for grp in e.groups:
    grp.table.zero_()
model.embedding_load("./ebc_ckpt", ["0", "6"])
for n in ("0", "6"):
    after = [x[1] for x in e.dump_table_local(n)]
    print("table", n, "max |diff| after reload:", max(float((a - b).abs().max()) for a, b in zip(after, before[n])))
