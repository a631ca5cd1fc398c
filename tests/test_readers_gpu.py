"""GPU runs of what round 1 could only validate on CPU: file readers feeding CUDA models through pinned ring
slots + copy-complete events, the device-resident evaluation cache, parallel collection dumps from device
tables and ``Model.resume``.  Guarded until a GPU box has run them once (tools_dev/next_round.sh)."""
import os

import numpy as np
import pytest
import torch

import hugectr_b200 as hugectr

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("HCTR_TEST_EXPERIMENTAL"),
                                 reason="not yet validated on a GPU box (tools_dev/next_round.sh runs it)")]


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
