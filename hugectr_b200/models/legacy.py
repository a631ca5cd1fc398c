"""Sample models of the reference built through the public API: DCN, DeepFM, Wide&Deep
(README quick-start, samples/deepfm/deepfm_parquet.py:77-279, samples/wdl/wdl_8gpu.py:79-283)."""
from __future__ import annotations

import hugectr_b200 as hugectr


def _reader(source, eval_source, fmt, slot_sizes):
    return hugectr.DataReaderParams(data_reader_type=fmt, source=[source], eval_source=eval_source,
                                    check_type=hugectr.Check_t.Non, slot_size_array=slot_sizes)


def build_dcn(batchsize=1024, vvgpu=None, source="synthetic", eval_source="synthetic",
              fmt=hugectr.DataReaderType_t.Parquet, slot_sizes=None, num_slots=26, vec=16,
              workspace_mb=75, lr=0.001, mixed=False, comm=None, **solver_kw):
    slot_sizes = list(slot_sizes or [1000] * num_slots)
    solver = hugectr.CreateSolver(max_eval_batches=solver_kw.pop("max_eval_batches", 10),
                                  batchsize_eval=solver_kw.pop("batchsize_eval", batchsize),
                                  batchsize=batchsize, lr=lr, vvgpu=vvgpu or [[0]],
                                  repeat_dataset=True, use_mixed_precision=mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slot_sizes), opt, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=13, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("data1", 1, True, num_slots)]))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                      workspace_size_per_gpu_in_mb=workspace_mb, embedding_vec_size=vec,
                                      combiner="sum", sparse_embedding_name="sparse_embedding1",
                                      bottom_name="data1", slot_size_array=slot_sizes, optimizer=opt))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["sparse_embedding1"], ["reshape1"],
                                 leading_dim=num_slots * vec))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["reshape1", "dense"], ["concat1"]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.MultiCross, ["concat1"], ["multicross1"], num_layers=6))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, ["concat1"], ["fc1"], num_output=1024))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.ReLU, ["fc1"], ["relu1"]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Dropout, ["relu1"], ["dropout1"], dropout_rate=0.5))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["dropout1", "multicross1"], ["concat2"]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, ["concat2"], ["fc2"], num_output=1))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.BinaryCrossEntropyLoss, ["fc2", "label"], ["loss"]))
    return model


def build_deepfm(batchsize=16384, vvgpu=None, source="synthetic", eval_source="synthetic",
                 fmt=hugectr.DataReaderType_t.Parquet, slot_sizes=None, num_slots=26, lr=0.001,
                 embedding_type=hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                 workspace_mb=61, mixed=False, comm=None, **solver_kw):
    slot_sizes = list(slot_sizes or [1000] * num_slots)
    vec = 11
    solver = hugectr.CreateSolver(max_eval_batches=solver_kw.pop("max_eval_batches", 10),
                                  batchsize_eval=solver_kw.pop("batchsize_eval", batchsize),
                                  batchsize=batchsize, lr=lr, vvgpu=vvgpu or [[0]],
                                  repeat_dataset=True, use_mixed_precision=mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, slot_sizes), opt, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=13, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("data1", 1, True, num_slots)]))
    model.add(hugectr.SparseEmbedding(embedding_type, workspace_mb, vec, "sum", "sparse_embedding1",
                                      "data1", slot_sizes, opt))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["sparse_embedding1"], ["reshape1"],
                                 leading_dim=vec))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Slice, ["reshape1"], ["slice11", "slice12"],
                                 ranges=[(0, 10), (10, 11)]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["slice11"], ["reshape2"],
                                 leading_dim=num_slots * 10))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["slice12"], ["reshape3"],
                                 leading_dim=num_slots))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.WeightMultiply, ["dense"], ["weight_multiply1"],
                                 weight_dims=[13, 10]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.WeightMultiply, ["dense"], ["weight_multiply2"],
                                 weight_dims=[13, 1]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["reshape2", "weight_multiply1"], ["concat1"]))
    prev = "concat1"
    for i in range(3):
        model.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, [prev], [f"fc{i + 1}"], num_output=400))
        model.add(hugectr.DenseLayer(hugectr.Layer_t.ReLU, [f"fc{i + 1}"], [f"relu{i + 1}"]))
        model.add(hugectr.DenseLayer(hugectr.Layer_t.Dropout, [f"relu{i + 1}"], [f"dropout{i + 1}"],
                                     dropout_rate=0.5))
        prev = f"dropout{i + 1}"
    model.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, [prev], ["fc4"], num_output=1))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.FmOrder2, ["concat1"], ["fmorder2"], out_dim=10))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.ReduceSum, ["fmorder2"], ["reducesum1"], axis=1))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["reshape3", "weight_multiply2"], ["concat2"]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.ReduceSum, ["concat2"], ["reducesum2"], axis=1))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Add, ["fc4", "reducesum1", "reducesum2"], ["add"]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.BinaryCrossEntropyLoss, ["add", "label"], ["loss"]))
    return model


def build_wdl(batchsize=16384, vvgpu=None, source="synthetic", eval_source="synthetic",
              fmt=hugectr.DataReaderType_t.Parquet, wide_slot_sizes=None, deep_slot_sizes=None,
              lr=0.001, workspace_mb=(23, 358), mixed=False, comm=None, etc=None, **solver_kw):
    """``etc=hugectr.CreateETC([TrainPSType_t.Cached, TrainPSType_t.Cached], ...)`` puts both tables on the
    host parameter server behind the HBM gpu_cache (BASELINE config 5)."""
    wide = list(wide_slot_sizes or [1000, 1000])
    deep = list(deep_slot_sizes or [1000] * 26)
    solver = hugectr.CreateSolver(max_eval_batches=solver_kw.pop("max_eval_batches", 10),
                                  batchsize_eval=solver_kw.pop("batchsize_eval", batchsize),
                                  batchsize=batchsize, lr=lr, vvgpu=vvgpu or [[0]],
                                  repeat_dataset=True, use_mixed_precision=mixed, **solver_kw)
    opt = hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam, hugectr.Update_t.Global)
    model = hugectr.Model(solver, _reader(source, eval_source, fmt, wide + deep), opt, etc=etc, comm=comm)
    model.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=13, dense_name="dense",
                            data_reader_sparse_param_array=[
                                hugectr.DataReaderSparseParam("wide_data", 1, True, len(wide)),
                                hugectr.DataReaderSparseParam("deep_data", 1, True, len(deep))]))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.DistributedSlotSparseEmbeddingHash,
                                      workspace_mb[0], 1, "sum", "sparse_embedding2", "wide_data",
                                      wide, opt))
    model.add(hugectr.SparseEmbedding(hugectr.Embedding_t.LocalizedSlotSparseEmbeddingHash,
                                      workspace_mb[1], 16, "sum", "sparse_embedding1", "deep_data",
                                      deep, opt))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["sparse_embedding1"], ["reshape1"],
                                 leading_dim=len(deep) * 16))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["sparse_embedding2"], ["reshape_wide"],
                                 leading_dim=len(wide)))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.ReduceSum, ["reshape_wide"], ["reshape2"], axis=1))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["reshape1", "dense"], ["concat1"]))
    prev = "concat1"
    for i in range(3):
        model.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, [prev], [f"fc{i + 1}"], num_output=1024))
        model.add(hugectr.DenseLayer(hugectr.Layer_t.ReLU, [f"fc{i + 1}"], [f"relu{i + 1}"]))
        model.add(hugectr.DenseLayer(hugectr.Layer_t.Dropout, [f"relu{i + 1}"], [f"dropout{i + 1}"],
                                     dropout_rate=0.5))
        prev = f"dropout{i + 1}"
    model.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, [prev], ["fc4"], num_output=1))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.Add, ["fc4", "reshape2"], ["add1"]))
    model.add(hugectr.DenseLayer(hugectr.Layer_t.BinaryCrossEntropyLoss, ["add1", "label"], ["loss"]))
    return model
