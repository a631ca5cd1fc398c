"""Further model families of the reference's sample directory, built through the public API:
NCF (GMF / MLP / NeuMF), MMoE (multi-task), DIN-style target attention and BST-style transformer
over a behaviour sequence.

Reference architectures: samples/ncf/{gmf,ncf,neumf}.py, samples/mmoe/{mmoe_parquet,shared_bottom}.py,
samples/din/din_parquet.py, samples/bst/bst_avg_pooling.py, samples/criteo/criteo_parquet.py (plain DNN),
samples/ftrl/dlrm_train_ftrl.py (DLRM-DCNv2 over an embedding collection trained with FTRL).  Every builder returns an un-compiled
``Model``; ``source="synthetic"`` trains on generated batches (no files needed).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import hugectr_b200 as hugectr

L = hugectr.Layer_t


def _solver(batchsize, lr, vvgpu, mixed, **kw):
    return hugectr.CreateSolver(max_eval_batches=kw.pop("max_eval_batches", 10),
                                batchsize_eval=kw.pop("batchsize_eval", batchsize),
                                batchsize=batchsize, lr=lr, vvgpu=vvgpu or [[0]], repeat_dataset=True,
                                use_mixed_precision=mixed, **kw)


def _reader(source, eval_source, fmt, slot_sizes):
    return hugectr.DataReaderParams(data_reader_type=fmt, source=[source], eval_source=eval_source,
                                    check_type=hugectr.Check_t.Non, slot_size_array=list(slot_sizes))


def _fc(model, bottom, top, n, act=True, dropout=0.0):
    model.add(hugectr.DenseLayer(L.InnerProduct, [bottom], [top + "_fc"], num_output=n))
    cur = top + "_fc"
    if act:
        model.add(hugectr.DenseLayer(L.ReLU, [cur], [top + "_relu"]))
        cur = top + "_relu"
    if dropout > 0:
        model.add(hugectr.DenseLayer(L.Dropout, [cur], [top + "_drop"], dropout_rate=dropout))
        cur = top + "_drop"
    return cur


# ----------------------------------------------------------------------------------------- NCF
def build_ncf(kind: str = "neumf", batchsize: int = 1024, num_users: int = 5000, num_items: int = 3000,
              gmf_dim: int = 16, mlp_dims: Sequence[int] = (64, 32, 16), source="synthetic",
              eval_source="synthetic", fmt=hugectr.DataReaderType_t.Parquet, lr: float = 0.0045,
              vvgpu=None, mixed: bool = False, comm=None, **solver_kw) -> "hugectr.Model":
    """kind: ``gmf`` (element-wise product of user/item vectors), ``mlp`` (MLP over the concatenated
    vectors) or ``neumf`` (both towers fused before the prediction layer).  Two slots: user, item."""
    assert kind in ("gmf", "mlp", "neumf")
    slots = [num_users, num_items]
    mlp_ev = mlp_dims[0] // 2
    ev = {"gmf": gmf_dim, "mlp": mlp_ev, "neumf": gmf_dim + mlp_ev}[kind]
    solver = _solver(batchsize, lr, vvgpu, mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global, beta1=0.25,
                                  beta2=0.5, epsilon=1e-8)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slots), opt, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=0, dense_name="dense",
                            data_reader_sparse_param_array=[hugectr.DataReaderSparseParam("data", 1, True, 2)]))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                      workspace_size_per_gpu_in_mb=max(1, (sum(slots) * ev * 4 * 3) >> 20),
                                      embedding_vec_size=ev, combiner="sum", sparse_embedding_name="emb",
                                      bottom_name="data", slot_size_array=slots, optimizer=opt))
    model.add(hugectr.DenseLayer(L.Reshape, ["emb"], ["emb_flat"], leading_dim=2 * ev))
    # [user(ev) | item(ev)] -> per tower slices
    tops = []
    if kind in ("gmf", "neumf"):
        model.add(hugectr.DenseLayer(L.Slice, ["emb_flat"], ["gmf_user", "gmf_item"],
                                     ranges=[(0, gmf_dim), (ev, ev + gmf_dim)]))
        model.add(hugectr.DenseLayer(L.ElementwiseMultiply, ["gmf_user", "gmf_item"], ["gmf_out"]))
        tops.append("gmf_out")
    if kind in ("mlp", "neumf"):
        off = gmf_dim if kind == "neumf" else 0
        model.add(hugectr.DenseLayer(L.Slice, ["emb_flat"], ["mlp_user", "mlp_item"],
                                     ranges=[(off, off + mlp_ev), (ev + off, ev + off + mlp_ev)]))
        model.add(hugectr.DenseLayer(L.Concat, ["mlp_user", "mlp_item"], ["mlp_in"]))
        cur = "mlp_in"
        for i, n in enumerate(mlp_dims[1:]):
            cur = _fc(model, cur, f"mlp{i}", n)
        tops.append(cur)
    if len(tops) == 2:
        model.add(hugectr.DenseLayer(L.Concat, tops, ["fused"]))
        last = "fused"
    else:
        last = tops[0]
    model.add(hugectr.DenseLayer(L.InnerProduct, [last], ["logit"], num_output=1))
    model.add(hugectr.DenseLayer(L.BinaryCrossEntropyLoss, ["logit", "label"], ["loss"]))
    return model


# ----------------------------------------------------------------------------------------- MMoE
def build_mmoe(batchsize: int = 1024, num_slots: int = 32, vocab: int = 2000, ev: int = 16,
               num_experts: int = 3, num_tasks: int = 2, expert_dims: Sequence[int] = (128, 64),
               tower_dim: int = 32, source="synthetic", eval_source="synthetic",
               fmt=hugectr.DataReaderType_t.Parquet, lr: float = 0.001, vvgpu=None, mixed: bool = False,
               comm=None, label_weights: Optional[List[float]] = None, **solver_kw) -> "hugectr.Model":
    """Multi-gate mixture of experts: shared embeddings -> E expert MLPs; per task a softmax gate
    over the experts, a weighted sum of the expert outputs, a tower and its own BCE loss."""
    slots = [vocab] * num_slots
    solver = _solver(batchsize, lr, vvgpu, mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slots), opt, comm=comm)
    names = [f"label{t}" for t in range(num_tasks)]
    model.add(hugectr.Input(label_dim=[1] * num_tasks, label_name=names, dense_dim=0, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("data", 1, True, num_slots)],
                            label_weights=label_weights or [1.0] * num_tasks))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                      workspace_size_per_gpu_in_mb=max(1, (sum(slots) * ev * 4 * 3) >> 20),
                                      embedding_vec_size=ev, combiner="sum", sparse_embedding_name="emb",
                                      bottom_name="data", slot_size_array=slots, optimizer=opt))
    model.add(hugectr.DenseLayer(L.Reshape, ["emb"], ["x"], leading_dim=num_slots * ev))
    experts = []
    for e in range(num_experts):
        cur = "x"
        for i, n in enumerate(expert_dims):
            cur = _fc(model, cur, f"e{e}_{i}", n, dropout=0.1)
        experts.append(cur)
    # experts stacked as [b, E, H]
    model.add(hugectr.DenseLayer(L.Concat, experts, ["experts_cat"]))
    model.add(hugectr.DenseLayer(L.Reshape, ["experts_cat"], ["experts"],
                                 shape=[-1, num_experts, expert_dims[-1]]))
    for t in range(num_tasks):
        model.add(hugectr.DenseLayer(L.InnerProduct, ["x"], [f"gate{t}_fc"], num_output=num_experts))
        model.add(hugectr.DenseLayer(L.Softmax, [f"gate{t}_fc"], [f"gate{t}"]))
        model.add(hugectr.DenseLayer(L.Reshape, [f"gate{t}"], [f"gate{t}_3d"], shape=[-1, 1, num_experts]))
        model.add(hugectr.DenseLayer(L.MatrixMultiply, [f"gate{t}_3d", "experts"], [f"mix{t}_3d"]))
        model.add(hugectr.DenseLayer(L.Reshape, [f"mix{t}_3d"], [f"mix{t}"], leading_dim=expert_dims[-1]))
        cur = _fc(model, f"mix{t}", f"tower{t}", tower_dim, dropout=0.1)
        model.add(hugectr.DenseLayer(L.InnerProduct, [cur], [f"logit{t}"], num_output=1))
        model.add(hugectr.DenseLayer(L.BinaryCrossEntropyLoss, [f"logit{t}", names[t]], [f"loss{t}"]))
    return model


def build_shared_bottom(batchsize: int = 1024, num_slots: int = 32, vocab: int = 2000, ev: int = 16,
                        shared_dims: Sequence[int] = (128, 256), tower_dim: int = 64, num_tasks: int = 2,
                        source="synthetic", eval_source="synthetic", fmt=hugectr.DataReaderType_t.Parquet,
                        lr: float = 0.001, vvgpu=None, mixed: bool = False, comm=None,
                        **solver_kw) -> "hugectr.Model":
    """Shared-bottom multi-task baseline of the MMoE sample: one shared MLP, fanned out (the graph
    compiler inserts the Slice) into one tower + BCE loss per task."""
    slots = [vocab] * num_slots
    solver = _solver(batchsize, lr, vvgpu, mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slots), opt, comm=comm)
    names = [f"label{t}" for t in range(num_tasks)]
    model.add(hugectr.Input(label_dim=[1] * num_tasks, label_name=names, dense_dim=0, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("data", 1, True, num_slots)]))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                      workspace_size_per_gpu_in_mb=max(1, (sum(slots) * ev * 4 * 3) >> 20),
                                      embedding_vec_size=ev, combiner="sum", sparse_embedding_name="emb",
                                      bottom_name="data", slot_size_array=slots, optimizer=opt))
    model.add(hugectr.DenseLayer(L.Reshape, ["emb"], ["x"], leading_dim=num_slots * ev))
    cur = "x"
    for i, n in enumerate(shared_dims):
        cur = _fc(model, cur, f"shared{i}", n, dropout=0.1)
    for t in range(num_tasks):
        tw = _fc(model, cur, f"tower{t}", tower_dim, dropout=0.1)
        model.add(hugectr.DenseLayer(L.InnerProduct, [tw], [f"logit{t}"], num_output=1))
        model.add(hugectr.DenseLayer(L.BinaryCrossEntropyLoss, [f"logit{t}", names[t]], [f"loss{t}"]))
    return model


def build_criteo_dnn(batchsize: int = 16384, num_slots: int = 26, vocab: int = 10000, ev: int = 64,
                     hidden: Sequence[int] = (200, 200, 200), source="synthetic", eval_source="synthetic",
                     fmt=hugectr.DataReaderType_t.Parquet, lr: float = 0.001, vvgpu=None, mixed: bool = False,
                     comm=None, **solver_kw) -> "hugectr.Model":
    """The plain Criteo DNN sample: categorical features only -> summed embeddings per slot -> MLP."""
    slots = [vocab] * num_slots
    solver = _solver(batchsize, lr, vvgpu, mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slots), opt, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=0, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("data1", 1, True, num_slots)]))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                      workspace_size_per_gpu_in_mb=max(1, (sum(slots) * ev * 4 * 3) >> 20),
                                      embedding_vec_size=ev, combiner="sum",
                                      sparse_embedding_name="sparse_embedding1", bottom_name="data1",
                                      slot_size_array=slots, optimizer=opt))
    model.add(hugectr.DenseLayer(L.Reshape, ["sparse_embedding1"], ["reshape1"], leading_dim=num_slots * ev))
    cur = "reshape1"
    for i, n in enumerate(hidden):
        cur = _fc(model, cur, f"fc{i + 1}", n)
    model.add(hugectr.DenseLayer(L.InnerProduct, [cur], ["logit"], num_output=1))
    model.add(hugectr.DenseLayer(L.BinaryCrossEntropyLoss, ["logit", "label"], ["loss"]))
    return model


def build_dlrm_ftrl(batchsize: int = 8192, num_gpus: int = 1, **kw) -> "hugectr.Model":
    """DLRM-DCNv2 over an embedding collection with the FTRL sparse / dense optimizer (samples/ftrl)."""
    from .dlrm import build_dlrm_dcnv2
    m = build_dlrm_dcnv2(batchsize=batchsize, num_gpus=num_gpus, **kw)
    ftrl = hugectr.CreateOptimizer(hugectr.Optimizer_t.Ftrl, hugectr.Update_t.Global, beta=0.9,
                                   lambda1=0.1, lambda2=0.1)
    m.opt_params = ftrl
    return m


# ----------------------------------------------------------------------------------------- DIN
def build_din(batchsize: int = 512, seq_len: int = 10, item_vocab: int = 4000, cate_vocab: int = 300,
              user_vocab: int = 1000, ev: int = 18, att_dims: Sequence[int] = (80, 40),
              mlp_dims: Sequence[int] = (200, 80), source="synthetic", eval_source="synthetic",
              fmt=hugectr.DataReaderType_t.Parquet, lr: float = 0.001, vvgpu=None, mixed: bool = False,
              comm=None, **solver_kw) -> "hugectr.Model":
    """Deep Interest Network: the candidate item attends over the user's behaviour history.
    Sparse inputs: user (1 slot), good = [candidate item + history items] (1 + T slots),
    cate = [candidate category + history categories] (1 + T slots)."""
    T = seq_len
    slots = [user_vocab] + [item_vocab] * (T + 1) + [cate_vocab] * (T + 1)
    solver = _solver(batchsize, lr, vvgpu, mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slots), opt, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=0, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("UserID", 1, True, 1),
                                hugectr.DataReaderSparseParam("GoodID", 1, True, T + 1),
                                hugectr.DataReaderSparseParam("CateID", 1, True, T + 1)]))
    for nm, bottom, sz in (("user_emb", "UserID", slots[:1]), ("good_emb", "GoodID", slots[1:T + 2]),
                           ("cate_emb", "CateID", slots[T + 2:])):
        model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                          workspace_size_per_gpu_in_mb=max(1, (sum(sz) * ev * 4 * 3) >> 20),
                                          embedding_vec_size=ev, combiner="sum", sparse_embedding_name=nm,
                                          bottom_name=bottom, slot_size_array=sz, optimizer=opt))
    # item_his [b*T, 2ev] (history), item [b, 2ev] (candidate)
    model.add(hugectr.DenseLayer(L.FusedReshapeConcat, ["good_emb", "cate_emb"], ["item_his", "item"]))
    model.add(hugectr.DenseLayer(L.Scale, ["item"], ["item_rep"], axis=1, factor=T))          # [b*T, 2ev]
    model.add(hugectr.DenseLayer(L.Sub, ["item_rep", "item_his"], ["att_sub"]))
    model.add(hugectr.DenseLayer(L.ElementwiseMultiply, ["item_rep", "item_his"], ["att_mul"]))
    model.add(hugectr.DenseLayer(L.Concat, ["item_rep", "item_his", "att_sub", "att_mul"], ["att_in"]))
    cur = "att_in"
    for i, n in enumerate(att_dims):
        model.add(hugectr.DenseLayer(L.InnerProduct, [cur], [f"att{i}_fc"], num_output=n))
        model.add(hugectr.DenseLayer(L.PReLU_Dice, [f"att{i}_fc"], [f"att{i}"], elu_alpha=0.2, eps=1e-8))
        cur = f"att{i}"
    model.add(hugectr.DenseLayer(L.InnerProduct, [cur], ["att_score"], num_output=1))          # [b*T, 1]
    model.add(hugectr.DenseLayer(L.Reshape, ["att_score"], ["att_score_bt"], leading_dim=T))      # [b, T]
    model.add(hugectr.DenseLayer(L.Softmax, ["att_score_bt"], ["att_w"]))
    model.add(hugectr.DenseLayer(L.Reshape, ["att_w"], ["att_w3"], shape=[-1, 1, T]))             # [b,1,T]
    model.add(hugectr.DenseLayer(L.Reshape, ["item_his"], ["item_his3"], shape=[-1, T, 2 * ev]))  # [b,T,2ev]
    model.add(hugectr.DenseLayer(L.MatrixMultiply, ["att_w3", "item_his3"], ["interest3"]))       # [b,1,2ev]
    model.add(hugectr.DenseLayer(L.Reshape, ["interest3"], ["interest"], leading_dim=2 * ev))
    model.add(hugectr.DenseLayer(L.Reshape, ["user_emb"], ["user"], leading_dim=ev))
    model.add(hugectr.DenseLayer(L.Concat, ["user", "interest", "item"], ["mlp_in"]))
    cur = "mlp_in"
    for i, n in enumerate(mlp_dims):
        model.add(hugectr.DenseLayer(L.InnerProduct, [cur], [f"fc{i}_fc"], num_output=n))
        model.add(hugectr.DenseLayer(L.PReLU_Dice, [f"fc{i}_fc"], [f"fc{i}"], elu_alpha=0.2, eps=1e-8))
        cur = f"fc{i}"
    model.add(hugectr.DenseLayer(L.InnerProduct, [cur], ["logit"], num_output=1))
    model.add(hugectr.DenseLayer(L.BinaryCrossEntropyLoss, ["logit", "label"], ["loss"]))
    return model


# ----------------------------------------------------------------------------------------- BST
def build_bst(batchsize: int = 512, seq_len: int = 8, item_vocab: int = 4000, user_vocab: int = 1000,
              ev: int = 32, heads: int = 4, ffn_dim: int = 64, mlp_dims: Sequence[int] = (128, 64),
              source="synthetic", eval_source="synthetic", fmt=hugectr.DataReaderType_t.Parquet,
              lr: float = 0.001, vvgpu=None, mixed: bool = False, comm=None, **solver_kw) -> "hugectr.Model":
    """Behaviour Sequence Transformer: one transformer block (multi-head self attention + residual
    LayerNorm + position-wise FFN) over [history items + candidate], average pooled, then an MLP."""
    S = seq_len + 1
    slots = [user_vocab] + [item_vocab] * S
    solver = _solver(batchsize, lr, vvgpu, mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slots), opt, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=0, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("UserID", 1, True, 1),
                                hugectr.DataReaderSparseParam("Seq", 1, True, S)]))
    for nm, bottom, sz in (("user_emb", "UserID", slots[:1]), ("seq_emb", "Seq", slots[1:])):
        model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                          workspace_size_per_gpu_in_mb=max(1, (sum(sz) * ev * 4 * 3) >> 20),
                                          embedding_vec_size=ev, combiner="sum", sparse_embedding_name=nm,
                                          bottom_name=bottom, slot_size_array=sz, optimizer=opt))
    # seq_emb: [b, S, ev]; projections act on the last dimension
    for nm in ("q", "k", "v"):
        model.add(hugectr.DenseLayer(L.InnerProduct, ["seq_emb"], [nm], num_output=ev))
    model.add(hugectr.DenseLayer(L.MultiHeadAttention, ["q", "k", "v"], ["attn"], num_attention_heads=heads))
    model.add(hugectr.DenseLayer(L.InnerProduct, ["attn"], ["attn_proj"], num_output=ev))
    model.add(hugectr.DenseLayer(L.Add, ["attn_proj", "seq_emb"], ["res1"]))
    model.add(hugectr.DenseLayer(L.LayerNorm, ["res1"], ["ln1"]))
    model.add(hugectr.DenseLayer(L.InnerProduct, ["ln1"], ["ffn1"], num_output=ffn_dim))
    model.add(hugectr.DenseLayer(L.ReLU, ["ffn1"], ["ffn1_relu"]))
    model.add(hugectr.DenseLayer(L.InnerProduct, ["ffn1_relu"], ["ffn2"], num_output=ev))
    model.add(hugectr.DenseLayer(L.Add, ["ffn2", "ln1"], ["res2"]))
    model.add(hugectr.DenseLayer(L.LayerNorm, ["res2"], ["ln2"]))
    model.add(hugectr.DenseLayer(L.ReduceMean, ["ln2"], ["pooled3"], axis=1))                     # [b,1,ev]
    model.add(hugectr.DenseLayer(L.Reshape, ["pooled3"], ["pooled"], leading_dim=ev))
    model.add(hugectr.DenseLayer(L.Reshape, ["user_emb"], ["user"], leading_dim=ev))
    model.add(hugectr.DenseLayer(L.Concat, ["user", "pooled"], ["mlp_in"]))
    cur = "mlp_in"
    for i, n in enumerate(mlp_dims):
        cur = _fc(model, cur, f"mlp{i}", n, dropout=0.1)
    model.add(hugectr.DenseLayer(L.InnerProduct, [cur], ["logit"], num_output=1))
    model.add(hugectr.DenseLayer(L.BinaryCrossEntropyLoss, ["logit", "label"], ["loss"]))
    return model
