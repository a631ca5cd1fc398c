"""GPU runs of what round 1 could only validate on CPU: file readers feeding CUDA models through pinned ring
slots + copy-complete events, the device-resident evaluation cache, parallel collection dumps from device
tables and ``Model.resume``.  Guarded until a GPU box has run them once (tools_dev/next_round.sh)."""
import os

import numpy as np
import pytest
import torch

import hugectr_b200 as hugectr

pytestmark = [pytest.mark.gpu]


def test_raw_reader_device_split_equals_host_split(tmp_path):
    """RawAsync device mode (O_DIRECT byte mover + csrc/reader_split.cu) produces exactly the tensors of the
    host-split mode: per-rank slices, feature-major keys, incomplete last batch, log1p of integer dense"""
    import types
    from hugectr_b200.data.raw_reader import RawAsyncReader
    N, hot = 1000, [3, 1, 7, 2]
    rng = np.random.default_rng(0)
    for dense_float in (True, False):
        path = str(tmp_path / f"t{int(dense_float)}.bin")
        lab = rng.integers(0, 2, (N, 1)).astype("<f4" if dense_float else "<u4")
        den = rng.random((N, 5)).astype("<f4") if dense_float else rng.integers(0, 1000, (N, 5)).astype("<u4")
        keys = rng.integers(0, 1 << 20, (N, sum(hot))).astype("<u4")
        np.concatenate([lab.view("<u4"), den.view("<u4"), keys], 1).astype("<u4").tofile(path)

        def stub(dev):
            m = types.SimpleNamespace()
            m.reader_params = types.SimpleNamespace(
                source=[path], eval_source=path, async_param=hugectr.AsyncParam(2, 3, is_dense_float=dense_float),
                float_label_dense=dense_float, num_samples=N, eval_num_samples=N)
            m.b_train = m.b_eval = 96
            m.comm = types.SimpleNamespace(rank=1)
            m.world = 3
            m.input = types.SimpleNamespace(label_dim=1, dense_dim=5)
            m.layout = types.SimpleNamespace(blocks=[(f"f{i}", 1, h, True) for i, h in enumerate(hot)])
            m.solver = types.SimpleNamespace(repeat_dataset=False, i64_input_key=False)
            m.key_dtype = torch.int32
            m.device = torch.device(dev)
            return m
        host = RawAsyncReader(stub("cpu"), True)
        devr = RawAsyncReader(stub("cuda"), True)
        assert devr.device_split and not host.device_split
        n = 0
        while True:
            a, b = host.read_a_batch(), devr.read_a_batch()
            assert (a is None) == (b is None)
            if a is None:
                break
            assert a.num_valid == b.num_valid
            sp = b.splitter
            lab_d = torch.empty(96, 1, device="cuda")
            den_d = torch.empty(96, 5, device="cuda")
            key_d = torch.empty(sp.total_keys, dtype=torch.int32, device="cuda")
            sp.run(b.raw, b.raw_skew, b.num_valid, lab_d, den_d, key_d)
            b.mark_copied()
            torch.cuda.synchronize()
            nv = a.num_valid
            torch.testing.assert_close(lab_d.cpu()[:nv], a.label[:nv])
            torch.testing.assert_close(den_d.cpu()[:nv], a.dense[:nv], atol=1e-6, rtol=1e-6)
            ka, kb = a.keys, key_d.cpu()
            off = 0
            for h in hot:
                assert torch.equal(ka[off:off + 96 * h].view(96, h)[:nv], kb[off:off + 96 * h].view(96, h)[:nv])
                off += 96 * h
            n += 1
        assert n == 4                      # ceil(1000 / 288)
        host.stop(); devr.stop()


@pytest.mark.parametrize("fmt", [hugectr.DataReaderType_t.Norm, hugectr.DataReaderType_t.Parquet,
                                 hugectr.DataReaderType_t.RawAsync])
def test_file_readers_feed_a_cuda_model(tmp_path, fmt):
    from hugectr_b200.data.generator import DataGenerator, DataGeneratorParams
    from hugectr_b200.models.legacy import build_dcn
    slots = [200] * 6
    raw = fmt == hugectr.DataReaderType_t.RawAsync
    p = DataGeneratorParams(fmt, 1, 13, 6, False, str(tmp_path / ("t.bin" if raw else "t.txt")),
                            str(tmp_path / ("v.bin" if raw else "v.txt")), slots, num_files=2, eval_num_files=1,
                            num_samples_per_file=2048, num_samples=4096, eval_num_samples=1024,
                            float_label_dense=True, check_type=hugectr.Check_t.Non)
    DataGenerator(p).generate()
    m = build_dcn(batchsize=256, source=p.source, eval_source=p.eval_source, slot_sizes=slots, num_slots=6,
                  fmt=fmt, workspace_mb=4, max_eval_batches=4, mixed=True)
    if raw:
        m.reader_params.num_samples, m.reader_params.eval_num_samples = 4096, 1024
        m.reader_params.async_param = hugectr.AsyncParam(2, 4, shuffle=False, multi_hot_reader=True,
                                                         is_dense_float=True)
    m.reader_params.cache_eval_data = 4
    m.compile()
    losses = []
    for _ in range(40):                     # several trips round the 4-slot staging ring
        assert m.train()
        losses.append(m.get_current_loss())
    assert np.isfinite(losses).all()
    aucs = []
    for _ in range(2):
        for _ in range(4):
            assert m.eval()
        aucs.append(m.get_eval_metrics()[0][1])
    assert aucs[0] == aucs[1]               # second round replayed from the device-resident cache


def test_parallel_dump_and_resume_on_cuda(tmp_path):
    from hugectr_b200.models.dlrm import build_dlrm_dcnv2

    def mk():
        m = build_dlrm_dcnv2(batchsize=256, num_gpus=1, table_sizes=[5000, 300, 50], multi_hot=[3, 1, 1],
                             ev_size=16, mixed=True, bottom=(32, 16), top=(32, 1), projection_dim=8,
                             cross_layers=1, lr=0.02, seed=11)
        m.compile()
        return m
    a = mk()
    pool = a.reader_train.pool
    pre = str(tmp_path / "s")
    for i in range(6):
        a.train_on_host_batch(pool[i % len(pool)])
        if i == 3:
            a.save_params_to_files(pre, 4)
    b = mk()
    assert b.resume(pre) == 4
    for i in range(4, 6):
        b.train_on_host_batch(pool[i % len(pool)])
    torch.cuda.synchronize()
    assert float((a.arena.weights - b.arena.weights).abs().max()) < 1e-6
    for name in ("0", "1", "2"):
        for pa, pb in zip(a.ebcs_train[0].dump_table_local(name), b.ebcs_train[0].dump_table_local(name)):
            assert float((pa[1] - pb[1]).abs().max()) < 1e-6
