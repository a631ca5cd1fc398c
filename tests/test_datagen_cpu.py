"""Native data set writer (csrc/host/data_generator.cpp): format, determinism, key distribution."""
import hashlib
import struct

import numpy as np
import pytest

import hugectr_b200 as hugectr
from hugectr_b200.data.generator import DataGenerator, DataGeneratorParams


def _md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


def _params(tmp, fmt, threads, **kw):
    slots = [1000, 7, 1, 50000]
    base = dict(num_files=3, eval_num_files=1, num_samples_per_file=500, num_samples=150000,
                eval_num_samples=1000, num_threads=threads)
    base.update(kw)
    return DataGeneratorParams(fmt, 2, 3, 4, kw.pop("i64", True), str(tmp / "train.txt" if fmt ==
                               hugectr.DataReaderType_t.Norm else tmp / "train.bin"),
                               str(tmp / "val.txt" if fmt == hugectr.DataReaderType_t.Norm else tmp / "val.bin"),
                               slots, nnz_array=[3, 1, 2, 4], **base)


def test_raw_bytes_do_not_depend_on_thread_count(tmp_path):
    sums = []
    for t in (1, 4):
        d = tmp_path / f"t{t}"
        d.mkdir()
        p = _params(d, hugectr.DataReaderType_t.Raw, t, float_label_dense=True)
        DataGenerator(p).generate()
        sums.append((_md5(p.source), _md5(p.eval_source)))
        rec = 4 * 5 + 10 * 8
        assert (d / "train.bin").stat().st_size == rec * 150000
    assert sums[0] == sums[1]
    assert sums[0][0] != sums[0][1]
    raw = np.fromfile(p.source, dtype=np.uint8).reshape(150000, -1)
    ld = raw[:, :20].copy().view("<f4")
    assert set(np.unique(ld[:, :2])) == {0.0, 1.0} and 0.45 < ld[:, :2].mean() < 0.55
    assert 0 <= ld[:, 2:].min() and ld[:, 2:].max() < 1 and abs(ld[:, 2:].mean() - 0.5) < 0.01
    keys = raw[:, 20:].copy().view("<i8")
    assert keys.shape == (150000, 10)
    for col, vocab in zip([0, 3, 4, 6], [1000, 7, 1, 50000]):
        assert keys[:, col].min() >= 0 and keys[:, col].max() < vocab
    # power law alpha 1.2 over [1, V+1): P(key = 0) = (1 - 2^(1-a)) / (1 - (V+1)^(1-a))
    a, V = 1.2, 50000
    p0 = (1 - 2 ** (1 - a)) / (1 - (V + 1) ** (1 - a))
    got = (keys[:, 6:10] == 0).mean()
    assert abs(got - p0) < 0.01, (got, p0)
    assert (keys[:, 6:10] < 10).mean() > (keys[:, 6:10] > 25000).mean() * 5


def test_raw_integer_records_and_uniform_keys(tmp_path):
    p = _params(tmp_path, hugectr.DataReaderType_t.Raw, 2, float_label_dense=False,
                dist_type=hugectr.Distribution_t.Uniform)
    p.i64_input_key = False
    DataGenerator(p).generate()
    raw = np.fromfile(p.source, dtype="<u4").reshape(150000, 5 + 10)
    assert set(np.unique(raw[:, :2])) == {0, 1}
    assert raw[:, 2:5].max() < 100
    k = raw[:, 11:15]
    assert k.max() < 50000 and abs(k.mean() - 25000) < 300          # uniform


@pytest.mark.parametrize("check", [hugectr.Check_t.Sum, hugectr.Check_t.Non])
def test_norm_files_are_well_formed(tmp_path, check):
    p = _params(tmp_path, hugectr.DataReaderType_t.Norm, 3, check_type=check)
    DataGenerator(p).generate()
    lines = open(p.source).read().split()
    assert int(lines[0]) == 3 and len(lines) == 4
    assert len({_md5(f) for f in lines[1:]}) == 3                    # every file has its own stream
    raw = open(lines[1], "rb").read()
    hdr = struct.unpack("<8q", raw[:64])
    assert hdr[:5] == (1 if check == hugectr.Check_t.Sum else 0, 500, 2, 3, 4)
    pos, cnts = 64, []
    for _ in range(500):
        if hdr[0]:
            nb = struct.unpack_from("<i", raw, pos)[0]
            body = raw[pos + 4:pos + 4 + nb]
            assert np.frombuffer(body, np.int8).sum(dtype=np.int8) == struct.unpack_from("<b", raw, pos + 4 + nb)[0]
            pos += 5 + nb
            q, buf = 0, body
        else:
            q, buf = pos, raw
        q += 20
        for s, (mx, vocab) in enumerate(zip([3, 1, 2, 4], [1000, 7, 1, 50000])):
            c = struct.unpack_from("<i", buf, q)[0]
            assert 1 <= c <= mx
            ks = np.frombuffer(buf, "<i8", c, q + 4)
            assert ks.min() >= 0 and ks.max() < vocab
            q += 4 + 8 * c
            cnts.append(c)
        if not hdr[0]:
            pos = q
        else:
            assert q == len(body)
    assert pos == len(raw)
    assert max(cnts) == 4
    # regenerating gives the same bytes (deterministic), the python writers remain selectable
    before = [_md5(f) for f in lines[1:]]
    DataGenerator(_params(tmp_path, hugectr.DataReaderType_t.Norm, 1, check_type=check)).generate()
    assert before == [_md5(f) for f in lines[1:]]


@pytest.mark.parametrize("odt", ["int32", "int64"])
def test_csr_to_padded_matches_python_loop(odt):
    from hugectr_b200.data.parquet_reader import _csr_to_padded
    rng = np.random.default_rng(3)
    n, S, H = 5000, 3, 4
    cnt = rng.integers(0, 7, n)
    offs = np.concatenate([[0], np.cumsum(cnt)]).astype(odt)
    vals = rng.integers(0, 1 << 40, int(cnt.sum()))
    blk = np.full((n + 5, S, H), -1, dtype="int64")
    nz = np.zeros((S, n + 5), dtype="int32")
    _csr_to_padded(offs, vals, n, H, 17, blk, 1, nz)
    exp = np.full_like(blk, -1)
    for i in range(n):
        c = min(cnt[i], H)
        exp[i, 1, :c] = vals[offs[i]:offs[i] + c] + 17
    assert (blk == exp).all()
    assert (nz[1, :n] == np.minimum(cnt, H)).all() and nz[0].sum() == 0 and nz[2].sum() == 0 and nz[1, n:].sum() == 0


def test_parquet_reader_stream_is_independent_of_the_worker_count(tmp_path):
    """row groups are decoded by `num_workers` threads but consumed in file order"""
    from types import SimpleNamespace as NS
    import torch
    from hugectr_b200.data.parquet_reader import ParquetReader
    from hugectr_b200.data.readers import SparseLayout
    p = DataGeneratorParams(hugectr.DataReaderType_t.Parquet, 1, 3, 4, True, str(tmp_path / "train.txt"),
                            str(tmp_path / "val.txt"), [1000, 7, 50, 50000], nnz_array=[3, 1, 2, 4],
                            num_files=3, eval_num_files=1, num_samples_per_file=70000)   # 2 row groups / file
    DataGenerator(p).generate()
    params = [NS(top_name=f"p{i}", slot_num=1, nnz_per_slot=[h], is_fixed_length=False)
              for i, h in enumerate([3, 1, 2, 4])]

    def stream(workers, rank):
        m = NS(reader_params=NS(source=[p.source], eval_source=p.eval_source, slot_size_array=None,
                                num_workers=workers),
               b_train=4096, b_eval=4096, comm=NS(rank=rank), world=2,
               solver=NS(repeat_dataset=False, drop_incomplete_batch=False),
               input=NS(label_dim=1, dense_dim=3), layout=SparseLayout(params), key_dtype=torch.int64,
               sparse_embeddings=[])
        r = ParquetReader(m, True)
        out = []
        while True:
            hb = r.read_a_batch()
            if hb is None:
                break
            out.append((hb.label.clone(), hb.keys.clone(), hb.nnz.clone(), hb.num_valid, r.get_current_batchsize()))
        r.stop()
        return out
    a, b = stream(1, 1), stream(6, 1)
    assert len(a) == len(b) == -(-210000 // 8192)
    for x, y in zip(a, b):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) and torch.equal(x[2], y[2]) and x[3:] == y[3:]
    assert a[-1][4] == 210000 - 8192 * (len(a) - 1)          # incomplete last batch is delivered
