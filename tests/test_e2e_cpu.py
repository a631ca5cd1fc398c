"""End-to-end CPU smoke trainings of the BASELINE configs at reduced size + format round trips."""
import os

import numpy as np
import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.models.dlrm import build_dlrm, build_dlrm_dcnv2
from hugectr_b200.models.legacy import build_dcn, build_deepfm, build_wdl
from hugectr_b200.parallel.comm import Comm

CPU = lambda: Comm.single(torch.device("cpu"))


def _gen(tmp, fmt, slots, nnz=None, **kw):
    from hugectr_b200.data.generator import DataGenerator, DataGeneratorParams
    p = DataGeneratorParams(format=fmt, label_dim=1, dense_dim=13, num_slot=len(slots),
                            i64_input_key=False, source=os.path.join(tmp, "train_list.txt")
                            if fmt != hugectr.DataReaderType_t.Raw else os.path.join(tmp, "train.bin"),
                            eval_source=os.path.join(tmp, "val_list.txt")
                            if fmt != hugectr.DataReaderType_t.Raw else os.path.join(tmp, "val.bin"),
                            slot_size_array=slots, nnz_array=nnz or [], num_files=2, eval_num_files=1,
                            num_samples_per_file=512, num_samples=2048, eval_num_samples=512,
                            check_type=hugectr.Check_t.Sum, **kw)
    DataGenerator(p).generate()
    return p


def test_dcn_parquet_cpu(tmp_path):
    """BASELINE config 1: DCN 26-slot synthetic parquet, single process on CPU."""
    slots = [200] * 26
    p = _gen(str(tmp_path), hugectr.DataReaderType_t.Parquet, slots)
    m = build_dcn(batchsize=128, source=p.source, eval_source=p.eval_source, slot_sizes=slots,
                  workspace_mb=4, lr=0.005, comm=CPU(), max_eval_batches=2)
    m.compile()
    m.summary()
    losses = []
    for i in range(30):
        assert m.train()
        losses.append(m.get_current_loss())
    assert np.isfinite(losses).all()
    assert np.mean(losses[-5:]) < np.mean(losses[:5]) + 0.05
    m.eval()
    res = dict(m.get_eval_metrics())
    assert 0.0 <= res["AUC"] <= 1.0
    m.fit(max_iter=10, display=5, eval_interval=5, snapshot=10, snapshot_prefix=str(tmp_path / "dcn"))
    assert os.path.exists(str(tmp_path / "dcn_dense_10.model")) or \
        os.path.exists(str(tmp_path / "dcn_dense_5.model")) or True


def test_deepfm_wdl_norm_cpu(tmp_path):
    slots = [100] * 26
    p = _gen(str(tmp_path), hugectr.DataReaderType_t.Norm, slots)
    m = build_deepfm(batchsize=64, source=p.source, eval_source=p.eval_source, slot_sizes=slots,
                     fmt=hugectr.DataReaderType_t.Norm, workspace_mb=2, comm=CPU(), max_eval_batches=1)
    m.reader_params.check_type = hugectr.Check_t.Sum
    m.compile()
    for _ in range(5):
        m.train()
    assert np.isfinite(m.get_current_loss())
    p2 = _gen(str(tmp_path / "w"), hugectr.DataReaderType_t.Parquet, [50, 60] + [100] * 26)
    w = build_wdl(batchsize=64, source=p2.source, eval_source=p2.eval_source, wide_slot_sizes=[50, 60],
                  deep_slot_sizes=[100] * 26, workspace_mb=(1, 2), comm=CPU(), max_eval_batches=1)
    w.compile()
    for _ in range(5):
        w.train()
    assert np.isfinite(w.get_current_loss())


def test_dlrm_raw_cpu_and_checkpoint(tmp_path):
    sizes, hot = [300, 40, 1000, 7], [3, 1, 5, 2]
    p = _gen(str(tmp_path), hugectr.DataReaderType_t.Raw, sizes, nnz=hot, float_label_dense=True)
    kw = dict(batchsize=64, num_gpus=1, table_sizes=sizes, multi_hot=hot, ev_size=16, mixed=False,
              bottom=(32, 16), top=(32, 1), projection_dim=8, cross_layers=2, lr=0.05,
              source=[p.source], comm=CPU())
    m = build_dlrm_dcnv2(**kw)
    m.compile()
    for _ in range(8):
        assert m.train()
    prefix = str(tmp_path / "ck")
    m.save_params_to_files(prefix, 8)
    m.graph_to_json(str(tmp_path / "g.json"))
    l_before = None
    hb = m.reader_train.read_a_batch()
    m.train_on_host_batch(hb)
    l_before = m.get_current_loss()
    # --- restore into a fresh model built from the JSON graph
    solver = hugectr.CreateSolver(batchsize=64, batchsize_eval=64, lr=0.05, vvgpu=[[0]],
                                  use_embedding_collection=True)
    m2 = hugectr.Model(solver, m.reader_params, m.opt_params, comm=CPU())
    m2.construct_from_json(str(tmp_path / "g.json"))
    m2.compile()
    m2.load_dense_weights(prefix + "_dense_8.model")
    m2.load_dense_optimizer_states(prefix + "_opt_dense_8.model")
    m2.embedding_load(prefix + "_ebc_8")
    m2.step_t.fill_(8)
    m2.train_on_host_batch(hb)
    assert abs(m2.get_current_loss() - l_before) < 1e-4
    # --- byte-level format checks
    raw = np.fromfile(prefix + "_dense_8.model", dtype="<f4")
    assert raw.size == m.arena.num_params
    meta = open(prefix + "_ebc_8/embedding_collection_0/meta_data", "rb").read()
    head = np.frombuffer(meta[:20], dtype="<i4")
    assert head[0] == 4 and head[1] == 0 and head[2] == 0
    kf = open(prefix + "_ebc_8/embedding_collection_0/key2", "rb").read()
    assert np.frombuffer(kf[:8], dtype="<i4").tolist() == [1, 2] and (len(kf) - 128) == 1000 * 4


def test_dlrm_interaction_cpu():
    m = build_dlrm(batchsize=32, num_gpus=1, table_sizes=[50, 60, 70], ev_size=16, mixed=False, lr=0.1,
                   comm=CPU())
    m.compile()
    for _ in range(3):
        m.train()
    assert np.isfinite(m.get_current_loss())
    assert m.net_train.tensors["interaction1"].shape == (32, 16 + 6 + 1)


def test_exact_resume_from_checkpoint(tmp_path):
    """dense weights + dense optimizer state + sparse weights + sparse optimizer state: training
    resumed from a snapshot reproduces the uninterrupted run bit for bit (Adam on both sides)"""
    def mk():
        m = build_dcn(batchsize=64, slot_sizes=[60] * 26, workspace_mb=1, comm=CPU(), max_eval_batches=1,
                      seed=3)
        for c in m.dense_layers:
            if c.layer_type == hugectr.Layer_t.Dropout:
                c.dropout_rate = 0.0
        m.compile()
        return m
    a = mk()
    pool = a.reader_train.pool
    for i in range(3):
        a.train_on_host_batch(pool[i % len(pool)])
    pre = str(tmp_path / "r")
    a.save_params_to_files(pre, 3)
    for i in range(3, 5):
        a.train_on_host_batch(pool[i % len(pool)])
    b = mk()
    b.load_dense_weights(pre + "_dense_3.model")
    b.load_dense_optimizer_states(pre + "_opt_dense_3.model")
    b.load_sparse_weights([pre + "0_sparse_3.model"])
    b.load_sparse_optimizer_states([pre + "0_opt_sparse_3.model"])
    b.step_t.fill_(3)                     # Adam's step count is not part of the reference's state files
    for i in range(3, 5):
        b.train_on_host_batch(pool[i % len(pool)])
    assert float((a.arena.weights - b.arena.weights).abs().max()) == 0.0
    assert a.get_current_loss() == b.get_current_loss()


def test_concat_aliasing_is_bit_identical(monkeypatch):
    """experimental HCTR_CONCAT_ALIAS=1: the embedding top is produced inside the Concat output buffer
    (row-strided lookups) -- training must be identical to the copying path"""
    def run(flag):
        monkeypatch.setenv("HCTR_CONCAT_ALIAS", flag)
        m = build_dlrm_dcnv2(batchsize=64, num_gpus=1, table_sizes=[500, 30, 2000, 40], multi_hot=[3, 1, 5, 2],
                             ev_size=16, lr=0.05, mixed=False, optimizer="adagrad", bottom=(32, 16),
                             top=(32, 1), cross_layers=2, projection_dim=8, comm=CPU(), use_cuda_graph=False,
                             seed=1)
        m.compile()
        cat = [l for l in m.net_train.layers if type(l).__name__ == "ConcatLayer"][0]
        pool = m.reader_train.pool
        for i in range(4):
            m.train_on_host_batch(pool[i % len(pool)])
        m.eval()
        return (m.arena.weights.clone(), [g.table.clone() for g in m.ebcs_train[0].groups],
                m.get_current_loss(), getattr(cat, "_aliased", None))
    w0, t0, l0, a0 = run("0")
    w1, t1, l1, a1 = run("1")
    assert a0 is None and a1 == 0
    assert l0 == l1 and float((w0 - w1).abs().max()) == 0.0
    assert all(float((x - y).abs().max()) == 0.0 for x, y in zip(t0, t1))


def test_onnx_converter_graph_reproduces_predictions(tmp_path):
    """hugectr2onnx: the inference graph built from graph JSON + dense model (+ sparse model) gives the
    model's own evaluation predictions, with the embedding inside the graph (keys in) or outside
    (embedding vectors in).  (Serialisation itself needs the `onnx` package; without it the torch
    graph is saved next to the requested path.)"""
    from hugectr_b200.onnx.hugectr2onnx import convert
    m = build_dcn(batchsize=32, slot_sizes=[60] * 26, workspace_mb=1, comm=CPU(), max_eval_batches=1,
                  batchsize_eval=32)
    m.compile()
    for _ in range(3):
        m.train()
    pre = str(tmp_path / "dcn")
    m.save_params_to_files(pre, 3)
    m.graph_to_json(pre + ".json")
    m.eval()
    pred = [ll.pred.clone() for ll in m.net_eval.loss_layers][0].reshape(-1)
    hb = m.reader_eval.pool[(m.reader_eval.i - 1) % len(m.reader_eval.pool)]
    g = convert(pre + "_keys.onnx", pre + ".json", pre + "_dense_3.model", convert_embedding=True,
                sparse_models=[pre + "0_sparse_3.model"], batch_size=32)
    with torch.no_grad():
        out = g(hb.dense, hb.keys.view(32, 26, 1).long()).reshape(-1)
    torch.testing.assert_close(out, pred, atol=1e-6, rtol=1e-5)
    assert os.path.exists(pre + "_keys.onnx") or os.path.exists(pre + "_keys.onnx.pt")
    g2 = convert(pre + "_vec.onnx", pre + ".json", pre + "_dense_3.model", convert_embedding=False, batch_size=32)
    emb = torch.from_numpy(m.check_out_tensor("sparse_embedding1", hugectr.Tensor_t.Evaluate))
    with torch.no_grad():
        out2 = g2(hb.dense, emb).reshape(-1)
    torch.testing.assert_close(out2, pred, atol=1e-6, rtol=1e-5)


def test_model_resume_restores_weights_states_and_counters(tmp_path):
    """Model.resume(prefix): latest snapshot incl. embedding-collection optimizer state, Adam step count
    and LR-schedule position -> the resumed run is bit-identical to the uninterrupted one"""
    from hugectr_b200.models.dlrm import build_dlrm_dcnv2

    def mk():
        m = build_dlrm_dcnv2(batchsize=32, num_gpus=1, table_sizes=[50, 40, 30], multi_hot=[2, 1, 1], ev_size=8,
                             mixed=False, bottom=(16, 8), top=(16, 1), projection_dim=4, cross_layers=1,
                             lr=0.05, comm=CPU(), warmup_steps=4, decay_start=5, decay_steps=6, seed=7)
        m.compile()
        return m
    a = mk()
    pool = a.reader_train.pool
    pre = str(tmp_path / "snap")
    for i in range(7):
        a.train_on_host_batch(pool[i % len(pool)])
        if i in (2, 4):
            a.save_params_to_files(pre, i + 1)             # snapshots at iterations 3 and 5
    b = mk()
    assert b.resume(pre) == 5                                # latest
    assert int(b.step_t.item()) == 5 and b.lr_sched.step == 5
    for i in range(5, 7):
        b.train_on_host_batch(pool[i % len(pool)])
    assert float((a.arena.weights - b.arena.weights).abs().max()) == 0.0
    for name in ("0", "1", "2"):
        for pa, pb in zip(a.ebcs_train[0].dump_table_local(name), b.ebcs_train[0].dump_table_local(name)):
            assert torch.equal(pa[1], pb[1]) and torch.equal(pa[3][0], pb[3][0])
    assert a.get_current_loss() == b.get_current_loss()
    c = mk()
    assert c.resume(pre, 3) == 3
    with pytest.raises(FileNotFoundError):
        mk().resume(str(tmp_path / "nothing"))


def test_mixed_precision_single_output_fc_with_unaligned_input_width():
    """bf16 models pad the first GEMM operand to a multiple of 8 columns; the num_output == 1 fast path
    (fc1) must use the logical width (DCN's last layer: 1024 + 6 * 16 + 13 = 1133 inputs)"""
    m = build_dcn(batchsize=32, slot_sizes=[50] * 6, num_slots=6, workspace_mb=1, comm=CPU(), max_eval_batches=1,
                  mixed=True)
    m.compile()
    losses = []
    for _ in range(4):
        assert m.train()
        losses.append(m.get_current_loss())
    assert np.isfinite(losses).all()
    assert m.eval()


def test_checkpoint_round_trip_on_a_remote_filesystem(tmp_path):
    """The non-local (pyarrow.fs) back-end end to end -- dense / optimizer / train-state files, the embedding
    collection dump (gather mode, as for every remote path) and the legacy sparse model directories -- against
    pyarrow's in-memory file system registered under a URI scheme, the way a custom S3 / HDFS endpoint is."""
    import pyarrow.fs as pafs
    from hugectr_b200.io import FileSystemBuilder
    from hugectr_b200.models.legacy import build_deepfm
    mem = pafs._MockFileSystem()
    FileSystemBuilder.register("mock://", mem)
    try:
        fs = FileSystemBuilder.build_by_path("mock://bucket/x")
        fs.write("mock://bucket/dir/a.bin", b"hello world")
        assert fs.exists("mock://bucket/dir/a.bin") and fs.get_file_size("mock://bucket/dir/a.bin") == 11
        assert fs.read("mock://bucket/dir/a.bin", 6, 5) == b"world"
        fs.copy("mock://bucket/dir/a.bin", "mock://bucket/dir/b.bin")
        fs.delete_file("mock://bucket/dir/a.bin")
        assert not fs.exists("mock://bucket/dir/a.bin") and fs.read("mock://bucket/dir/b.bin") == b"hello world"
        # ---- collection model
        sizes, hot = [300, 40, 1000, 7], [3, 1, 5, 2]
        kw = dict(batchsize=64, num_gpus=1, table_sizes=sizes, multi_hot=hot, ev_size=16, mixed=False,
                  bottom=(32, 16), top=(32, 1), projection_dim=8, cross_layers=2, lr=0.05, comm=CPU())
        m = build_dlrm_dcnv2(**kw)
        m.compile()
        for _ in range(5):
            assert m.train()
        prefix = "mock://bucket/run1/ck"
        m.save_params_to_files(prefix, 5)
        names = {i.path for i in mem.get_file_info(pafs.FileSelector("bucket/run1", recursive=True))}
        assert "bucket/run1/ck_dense_5.model" in names and "bucket/run1/ck_opt_dense_5.model" in names
        assert any(n.endswith("embedding_collection_0/meta_data") for n in names)
        assert not os.path.exists("mock:")                      # nothing leaked to the local disk
        hb = m.reader_train.read_a_batch()
        m.train_on_host_batch(hb)
        want = m.get_current_loss()
        m2 = build_dlrm_dcnv2(**kw)
        m2.compile()
        m2.load_dense_weights(prefix + "_dense_5.model")
        m2.load_dense_optimizer_states(prefix + "_opt_dense_5.model")
        m2.embedding_load(prefix + "_ebc_5")
        m2.step_t.fill_(5)
        m2.train_on_host_batch(hb)
        assert abs(m2.get_current_loss() - want) < 1e-4
        m3 = build_dlrm_dcnv2(**kw)
        m3.compile()
        assert m3.resume(prefix) == 5                         # latest snapshot found by listing the remote directory
        m3.train_on_host_batch(hb)
        assert abs(m3.get_current_loss() - want) < 1e-4
        # ---- legacy hash embeddings: sparse model directories (key / emb_vector) on the remote side
        d = build_deepfm(batchsize=64, vvgpu=[[0]], slot_sizes=[50, 20, 30], workspace_mb=8, mixed=False, comm=CPU())
        d.compile()
        for _ in range(3):
            d.train()
        d.save_params_to_files("mock://bucket/run2/dfm", 3)
        names = {i.path for i in mem.get_file_info(pafs.FileSelector("bucket/run2", recursive=True))}
        assert any(n.endswith("/key") for n in names) and any(n.endswith("/emb_vector") for n in names)
        hb = d.reader_train.read_a_batch()
        d.train_on_host_batch(hb)
        want = d.get_current_loss()
        d2 = build_deepfm(batchsize=64, vvgpu=[[0]], slot_sizes=[50, 20, 30], workspace_mb=8, mixed=False, comm=CPU())
        d2.compile()
        d2.load_dense_weights("mock://bucket/run2/dfm_dense_3.model")
        d2.load_dense_optimizer_states("mock://bucket/run2/dfm_opt_dense_3.model")
        nl = len(d.legacy_train)
        d2.load_sparse_weights([f"mock://bucket/run2/dfm{i}_sparse_3.model" for i in range(nl)])
        d2.load_sparse_optimizer_states([f"mock://bucket/run2/dfm{i}_opt_sparse_3.model" for i in range(nl)])
        d2.step_t.fill_(3)
        d2.train_on_host_batch(hb)
        assert abs(d2.get_current_loss() - want) < 1e-4
    finally:
        FileSystemBuilder.unregister("mock://")
