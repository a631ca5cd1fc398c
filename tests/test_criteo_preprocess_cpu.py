"""Native Criteo TSV preprocessing against the numpy implementation of the same categorification."""
import numpy as np
import pytest

from hugectr_b200.tools import criteo2raw


def _tsv(path, n, rng, trailing_newline=True):
    lines = []
    for _ in range(n):
        f = [str(int(rng.integers(0, 2))) if rng.random() > 0.02 else ""]
        for _ in range(13):
            r = rng.random()
            f.append("" if r < 0.1 else str(int(rng.integers(-3, 5000))))
        for j in range(26):
            r = rng.random()
            card = [3, 50, 2000][j % 3]
            f.append("" if r < 0.05 else format(int(rng.zipf(1.3)) % card * 2654435761 % (1 << 32), "08x"))
        lines.append("\t".join(f))
    open(path, "w").write("\n".join(lines) + ("\n" if trailing_newline else ""))


@pytest.mark.parametrize("threads,min_freq,rng_mod,nl", [(1, 1, 0, True), (5, 3, 0, False), (8, 2, 40, True)])
def test_native_matches_numpy(tmp_path, threads, min_freq, rng_mod, nl):
    rng = np.random.default_rng(11)
    a, b = str(tmp_path / "a.tsv"), str(tmp_path / "b.tsv")
    _tsv(a, 3000, rng, nl)
    _tsv(b, 500, rng)
    vocabs, sizes = criteo2raw.convert(a, str(tmp_path / "a_py.bin"), None, min_freq, rng_mod)
    criteo2raw.convert(b, str(tmp_path / "b_py.bin"), vocabs, min_freq, rng_mod)
    pre = criteo2raw.CriteoPreprocessor(num_threads=threads).fit(a)
    assert pre.num_lines == 3000
    got_sizes = pre.finalize(min_freq, rng_mod)
    assert got_sizes == sizes
    assert pre.transform(a, str(tmp_path / "a_nat.bin"), rng_mod) == 3000
    assert pre.transform(b, str(tmp_path / "b_nat.bin"), rng_mod) == 500
    for n in ("a", "b"):
        x = np.fromfile(tmp_path / f"{n}_py.bin", dtype="<u4").reshape(-1, 40)
        y = np.fromfile(tmp_path / f"{n}_nat.bin", dtype="<u4").reshape(-1, 40)
        assert x.shape == y.shape and (x == y).all(), np.argwhere(x != y)[:5]
    # vocabulary round trip through a fresh object, and appending behind existing records
    pre2 = criteo2raw.CriteoPreprocessor(num_threads=2)
    for j in range(26):
        pre2.load_vocabulary(j, pre.vocabulary(j))
    out = str(tmp_path / "both.bin")
    n0 = pre2.transform(a, out, rng_mod)
    pre2.transform(b, out, rng_mod, append_at=n0)
    both = np.fromfile(out, dtype="<u4").reshape(-1, 40)
    assert both.shape[0] == 3500 and (both[3000:] == y).all()


def test_cli(tmp_path, capsys):
    rng = np.random.default_rng(1)
    a = str(tmp_path / "day_0")
    _tsv(a, 200, rng)
    sizes = criteo2raw.main(["--train", a, "--out-dir", str(tmp_path / "o"), "--threads", "3"])
    assert len(sizes) == 26 and (tmp_path / "o" / "day_0.bin").stat().st_size == 200 * 160
    assert "slot_size_array" in capsys.readouterr().out


@pytest.mark.parametrize("compressed", [False, True])
def test_mlperf_numpy_days_to_raw(tmp_path, compressed):
    """tools/convert_to_raw.py: memory-mapped day files -> train / val / test raw records"""
    from hugectr_b200.tools import convert_to_raw as C
    rng = np.random.default_rng(5)
    hot = [3, 2, 1, 2, 6, 1, 1, 1, 1, 7, 3, 8, 1, 6, 9, 5, 1, 1, 1, 12, 100, 27, 10, 3, 1, 1]
    days, rows = 3, [50, 70, 90]
    data = []
    for d in range(days):
        n = rows[d]
        lab = rng.integers(0, 2, n).astype(np.int32)       # the dtype of the MLPerf dumps
        den = rng.random((n, 13), dtype=np.float32)
        sp = {str(i): rng.integers(0, 1 << 20, (n, hot[i])).astype(np.int32) for i in range(26)}
        np.save(tmp_path / f"day_{d}_labels.npy", lab)
        np.save(tmp_path / f"day_{d}_dense.npy", den)
        (np.savez_compressed if compressed else np.savez)(tmp_path / f"day_{d}_sparse_multi_hot.npz", **sp)
        data.append((lab, den, sp))
    if not compressed:
        assert isinstance(C.npz_member(str(tmp_path / "day_0_sparse_multi_hot.npz"), "20"), np.memmap)
    out = tmp_path / "out"
    counts = C.main(["--input_dir_labels_and_dense", str(tmp_path), "--input_dir_sparse_multihot", str(tmp_path),
                     "--output_dir", str(out), "--num_days", "3", "--split_point", "60", "--chunk_size", "16"])
    assert counts == {"train": 120, "val": 60, "test": 30}

    def rows_of(lab, den, sp, lo, hi):        # the straightforward per-row writer as the oracle
        return b"".join(lab[i].tobytes() + den[i].tobytes() + b"".join(sp[str(f)][i].tobytes() for f in range(26))
                        for i in range(lo, hi))
    assert (out / "train_data.bin").read_bytes() == rows_of(*data[0], 0, 50) + rows_of(*data[1], 0, 70)
    assert (out / "val_data.bin").read_bytes() == rows_of(*data[2], 0, 60)
    assert (out / "test_data.bin").read_bytes() == rows_of(*data[2], 60, 90)
    assert (out / "val_data.bin").stat().st_size == 60 * 4 * (1 + 13 + sum(hot))


def test_converted_raw_data_trains_dlrm_dcnv2(tmp_path):
    """NumPy days -> convert_to_raw -> RawAsync reader -> DLRM-DCNv2 fit with evaluation (the MLPerf data
    path end to end on a toy data set)"""
    import torch
    from hugectr_b200.models.dlrm import build_dlrm_dcnv2
    from hugectr_b200.parallel.comm import Comm
    from hugectr_b200.tools import convert_to_raw as C
    rng = np.random.default_rng(9)
    sizes = [50, 30, 20, 40] + [10] * 22
    hot = [3, 2, 1, 2] + [1] * 22
    for d in range(2):
        n = 256
        sp = {str(i): rng.integers(0, sizes[i], (n, hot[i])).astype(np.int32) for i in range(26)}
        lab = (sp["2"][:, 0] % 2).astype(np.int32)              # learnable from the one-hot feature 2
        np.save(tmp_path / f"day_{d}_labels.npy", lab)
        np.save(tmp_path / f"day_{d}_dense.npy", rng.random((n, 13), dtype=np.float32))
        np.savez(tmp_path / f"day_{d}_sparse_multi_hot.npz", **sp)
    C.convert(str(tmp_path), str(tmp_path), str(tmp_path / "raw"), num_days=2, split_point=128, log=lambda *a: None)
    m = build_dlrm_dcnv2(batchsize=64, batchsize_eval=64, num_gpus=1, table_sizes=sizes, multi_hot=hot,
                         ev_size=8, mixed=False, bottom=(16, 8), top=(16, 1), projection_dim=4, cross_layers=1,
                         lr=0.05, source=[str(tmp_path / "raw" / "train_data.bin")],
                         comm=Comm.single(torch.device("cpu")), max_eval_batches=2)
    m.reader_params.eval_source = str(tmp_path / "raw" / "val_data.bin")
    m.reader_params.num_samples, m.reader_params.eval_num_samples = 256, 128
    m.compile()
    first = None
    for it in range(300):
        assert m.train()
        if it == 0:
            first = m.get_current_loss()
    assert m.get_current_loss() < first * 0.8
    for _ in range(2):
        assert m.eval()
    auc = dict(m.get_eval_metrics())["AUC"]
    assert auc > 0.85, auc


def test_criteo_tsv_to_parquet_and_train(tmp_path):
    """tools/criteo2parquet.py (the NVTabular preprocessing role): TSV -> categorified Parquet + metadata, ids equal
    the raw converter's, and the Parquet reader trains a model from the result"""
    import json
    import pyarrow.parquet as pq
    import torch
    import hugectr_b200 as hugectr
    from hugectr_b200.tools import criteo2parquet
    rng = np.random.default_rng(4)
    tr, va = str(tmp_path / "day_0"), str(tmp_path / "day_1")
    _tsv(tr, 700, rng)
    _tsv(va, 300, rng)
    sizes = criteo2parquet.convert([tr], [va], str(tmp_path / "pq"), freq_limit=2, normalize_dense=True,
                                   rows_per_file=256, threads=3, log=lambda *a: None)
    meta = json.load(open(tmp_path / "pq" / "train" / "_metadata.json"))
    assert sum(f["num_rows"] for f in meta["file_stats"]) == 700 and len(meta["file_stats"]) == 3
    assert [c["index"] for c in meta["cats"]] == list(range(14, 40))
    # same ids as the raw converter with the same vocabulary
    pre = criteo2raw.CriteoPreprocessor(num_threads=2).fit(tr)
    assert pre.finalize(2) == sizes
    pre.transform(tr, str(tmp_path / "t.bin"))
    raw = np.fromfile(tmp_path / "t.bin", dtype="<u4").reshape(700, 40)
    files = open(tmp_path / "pq" / "train" / "_file_list.txt").read().split()[1:]
    tbl = [pq.read_table(f).to_pandas() for f in files]
    import pandas as pd
    df = pd.concat(tbl, ignore_index=True)
    assert (df[[f"C{j + 1}" for j in range(26)]].to_numpy() == raw[:, 14:]).all()
    assert (df["label"].to_numpy() == raw[:, 0].view("<i4")).all()
    assert np.allclose(df["I3"].to_numpy(), np.log1p(np.maximum(raw[:, 3].view("<i4"), 0)), atol=1e-6)
    # train from it
    solver = hugectr.CreateSolver(batchsize=64, batchsize_eval=64, lr=0.01, vvgpu=[[0]], repeat_dataset=True,
                                  i64_input_key=True, use_cuda_graph=False, max_eval_batches=2)
    rp = hugectr.DataReaderParams(hugectr.DataReaderType_t.Parquet, source=[str(tmp_path / "pq" / "train" / "_file_list.txt")],
                                  eval_source=str(tmp_path / "pq" / "val" / "_file_list.txt"),
                                  check_type=hugectr.Check_t.Non, slot_size_array=sizes)
    m = hugectr.Model(solver, rp, hugectr.CreateOptimizer(hugectr.Optimizer_t.Adam))
    m.add(hugectr.Input(label_dim=1, label_name="label", dense_dim=13, dense_name="dense",
                        data_reader_sparse_param_array=[hugectr.DataReaderSparseParam("data1", 1, True, 26)]))
    m.add(hugectr.SparseEmbedding(hugectr.Embedding_t.LocalizedSlotSparseEmbeddingHash, 8, 4, "sum", "emb", "data1",
                                  slot_size_array=sizes))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.Reshape, ["emb"], ["r"], leading_dim=26 * 4))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.Concat, ["r", "dense"], ["c"]))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.InnerProduct, ["c"], ["fc"], num_output=1))
    m.add(hugectr.DenseLayer(hugectr.Layer_t.BinaryCrossEntropyLoss, ["fc", "label"], ["loss"]))
    m.compile()
    for _ in range(20):
        assert m.train()
    assert np.isfinite(m.get_current_loss())
    assert m.eval()


def test_criteo2predict_roundtrip(tmp_path):
    """tools/criteo2predict.py: the reference's four-line inference input format, written and read back"""
    import json
    from hugectr_b200.tools import criteo2predict as P
    rng = np.random.default_rng(2)
    b, D, slots = 5, 3, [1, 2, 1]
    rows = np.concatenate([rng.integers(0, 2, (b, 1)), rng.random((b, D)).round(4), rng.integers(0, 1000, (b, 4))], 1)
    np.savetxt(tmp_path / "test.txt", rows, fmt="%g")
    json.dump({"dense": D, "categorical": 4, "slot_size": slots}, open(tmp_path / "cfg.json", "w"))
    n = P.main(["--src_csv_path", str(tmp_path / "test.txt"), "--src_config_path", str(tmp_path / "cfg.json"),
                "--dst_path", str(tmp_path / "in.txt"), "--batch_size", "4"])
    assert n == 4
    lines = open(tmp_path / "in.txt").read().splitlines()
    assert len(lines) == 4 and len(lines[0].split()) == 4 and len(lines[2].split()) == 16 and len(lines[3].split()) == 13
    label, dense, keys, ptr = P.load(str(tmp_path / "in.txt"), D)
    assert np.allclose(label.numpy(), rows[:4, 0]) and np.allclose(dense.numpy(), rows[:4, 1:4], atol=1e-6)
    assert (keys.numpy() == rows[:4, 4:].astype(np.int64).reshape(-1)).all()
    assert ptr.tolist() == [0, 1, 3, 4, 5, 7, 8, 9, 11, 12, 13, 15, 16]
