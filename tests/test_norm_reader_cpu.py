"""Native Norm reader (csrc/host/norm_reader.cpp) against the pure-Python decoder of the same format."""
import struct
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import hugectr_b200 as hugectr
from hugectr_b200.data.norm_reader import DataCheckError, NormReader, PyNormReader
from hugectr_b200.data.readers import SparseLayout


def _write(path, n, L, D, S, maxnnz, vocab, check, i64, rng):
    recs = []
    for _ in range(n):
        body = rng.random(L + D).astype("<f4").tobytes()
        for s in range(S):
            c = int(rng.integers(0, maxnnz + 1))
            ks = rng.integers(0, vocab, c)
            body += struct.pack("<i", c) + ks.astype("<i8" if i64 else "<u4").tobytes()
        if check:
            cs = np.frombuffer(body, dtype=np.int8).sum(dtype=np.int8)
            body = struct.pack("<i", len(body)) + body + struct.pack("<b", int(cs))
        recs.append(body)
    with open(path, "wb") as f:
        f.write(struct.pack("<8q", int(check), n, L, D, S, 0, 0, 0))
        f.write(b"".join(recs))


def _model(flist, b, rank, world, L, D, params, check, i64, repeat, slot_sizes=None, key_dtype=torch.int64):
    layout = SparseLayout(params)
    return NS(reader_params=NS(source=[flist], eval_source=flist, check_type=check,
                               slot_size_array=slot_sizes),
              b_train=b, b_eval=b, comm=NS(rank=rank), world=world,
              solver=NS(repeat_dataset=repeat, i64_input_key=i64),
              input=NS(label_dim=L, dense_dim=D), layout=layout, key_dtype=key_dtype,
              sparse_embeddings=[1] if slot_sizes else [])


def _params(spec):
    return [NS(top_name=f"p{i}", slot_num=s, nnz_per_slot=[h] * s, is_fixed_length=False)
            for i, (s, h) in enumerate(spec)]


@pytest.mark.parametrize("check", [False, True])
@pytest.mark.parametrize("i64", [False, True])
@pytest.mark.parametrize("world", [1, 3])
def test_native_matches_python_decoder(tmp_path, check, i64, world):
    rng = np.random.default_rng(7)
    L, D = 2, 5
    spec = [(3, 4), (2, 2)]           # (slots, max nnz); the second block truncates longer bags
    S = sum(s for s, _ in spec)
    files = []
    for i, n in enumerate([37, 0, 21]):     # an empty file in the middle, ragged tail
        fp = str(tmp_path / f"f{i}.data")
        _write(fp, n, L, D, S, 4, 1000, check, i64, rng)
        files.append(fp)
    flist = str(tmp_path / "list.txt")
    open(flist, "w").write(f"{len(files)}\n" + "\n".join(files) + "\n")
    ck = hugectr.Check_t.Sum if check else hugectr.Check_t.Non
    slot_sizes = [1000] * S
    for rank in range(world):
        for kd in ((torch.int64, torch.int32) if rank == 0 else (torch.int64,)):
            args = (flist, 8, rank, world, L, D, _params(spec), ck, i64, False, slot_sizes, kd)
            nat, ref = NormReader(_model(*args), True), PyNormReader(_model(*args), True)
            assert type(nat) is NormReader
            nb = 0
            while True:
                a, b = nat.read_a_batch(), ref.read_a_batch()
                assert (a is None) == (b is None)
                if a is None:
                    break
                nb += 1
                assert nat.get_current_batchsize() == ref.get_current_batchsize()
                assert a.num_valid == b.num_valid
                assert torch.equal(a.label, b.label) and torch.equal(a.dense, b.dense)
                assert a.keys.dtype == kd and torch.equal(a.keys, b.keys)
                assert torch.equal(a.nnz, b.nnz)
            assert nb == -(-58 // (8 * world))
            assert nat.read_a_batch() is None          # end marker stays readable
            nat.stop()


def test_native_repeat_and_restart(tmp_path):
    rng = np.random.default_rng(1)
    fp = str(tmp_path / "f.data")
    _write(fp, 10, 1, 2, 2, 3, 50, True, True, rng)
    flist = str(tmp_path / "list.txt")
    open(flist, "w").write(f"1\n{fp}\n")
    m = _model(flist, 4, 0, 1, 1, 2, _params([(2, 3)]), hugectr.Check_t.Sum, True, True)
    r = NormReader(m, True)
    # 20 records = the file twice (batches straddle the wrap); slots are recycled -> clone at once
    a = torch.cat([r.read_a_batch().label.clone() for _ in range(5)])
    assert torch.equal(a[:10], a[10:20])
    r.set_source()                                   # restart from the top
    assert torch.equal(r.read_a_batch().label, a[:4])
    r.stop()


def test_native_errors(tmp_path):
    rng = np.random.default_rng(2)
    fp = str(tmp_path / "f.data")
    _write(fp, 16, 1, 2, 2, 3, 50, True, False, rng)
    flist = str(tmp_path / "list.txt")
    open(flist, "w").write(f"1\n{fp}\n")
    mk = lambda **kw: _model(flist, 4, 0, 1, kw.get("L", 1), 2, _params([(2, 3)]),
                             kw.get("ck", hugectr.Check_t.Sum), False, False)
    with pytest.raises(RuntimeError, match="does not match"):
        NormReader(mk(L=3), True).read_a_batch()
    with pytest.raises(RuntimeError, match="check_sum"):
        NormReader(mk(ck=hugectr.Check_t.Non), True).read_a_batch()
    raw = bytearray(open(fp, "rb").read())
    raw[len(raw) // 2] ^= 0x5A
    open(fp, "wb").write(raw)
    r = NormReader(mk(), True)
    with pytest.raises(DataCheckError):
        for _ in range(4):
            r.read_a_batch()
    open(fp, "wb").write(raw[:len(raw) - 9])
    r = NormReader(mk(), True)
    with pytest.raises(RuntimeError):
        for _ in range(4):
            r.read_a_batch()


def test_native_norm_and_raw_roundtrip_property_based(tmp_path_factory):
    """hypothesis: DataGenerator (native writers) -> native readers returns exactly the records written,
    for random label / dense / slot shapes, key widths, checksums, batch sizes and rank counts"""
    from hypothesis import given, settings, strategies as st
    from hugectr_b200.data.generator import DataGenerator, DataGeneratorParams
    from hugectr_b200.data.raw_reader import RawAsyncReader

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 3), st.integers(0, 5), st.lists(st.integers(1, 4), min_size=1, max_size=5),
           st.booleans(), st.booleans(), st.integers(1, 3), st.integers(3, 17))
    def check(L, D, nnz, i64, check, world, b):
        d = tmp_path_factory.mktemp("rt")
        S = len(nnz)
        sizes = [50 + 10 * i for i in range(S)]
        common = dict(label_dim=L, dense_dim=D, num_slot=S, i64_input_key=i64, slot_size_array=sizes,
                      nnz_array=nnz, num_files=2, eval_num_files=1, num_samples_per_file=23,
                      num_samples=46, eval_num_samples=5, float_label_dense=True,
                      check_type=hugectr.Check_t.Sum if check else hugectr.Check_t.Non)
        # ---- Norm: native reader == python decoder, all ranks together see every record once
        p = DataGeneratorParams(format=hugectr.DataReaderType_t.Norm, source=str(d / "n.txt"),
                                eval_source=str(d / "nv.txt"), **common)
        DataGenerator(p).generate()
        params = [NS(top_name=f"p{i}", slot_num=1, nnz_per_slot=[h], is_fixed_length=False) for i, h in enumerate(nnz)]
        seen = 0
        for rank in range(world):
            args = (p.source, b, rank, world, L, D, params, p.check_type, i64, False)
            nat, ref = NormReader(_model(*args), True), PyNormReader(_model(*args), True)
            while True:
                x, y = nat.read_a_batch(), ref.read_a_batch()
                assert (x is None) == (y is None)
                if x is None:
                    break
                assert torch.equal(x.label, y.label) and torch.equal(x.dense, y.dense)
                assert torch.equal(x.keys, y.keys) and torch.equal(x.nnz, y.nnz) and x.num_valid == y.num_valid
                seen += x.num_valid
            nat.stop()
        assert seen == 46
        # ---- Raw: fixed records, every rank reads only its slice
        pr = DataGeneratorParams(format=hugectr.DataReaderType_t.RawAsync, source=str(d / "r.bin"),
                                 eval_source=str(d / "rv.bin"), **common)
        DataGenerator(pr).generate()
        kt = np.dtype("<i8") if i64 else np.dtype("<u4")
        rec = 4 * (L + D) + sum(nnz) * kt.itemsize
        raw = np.fromfile(pr.source, dtype=np.uint8).reshape(46, rec)
        got = 0
        for rank in range(world):
            m = _model(pr.source, b, rank, world, L, D, params, hugectr.Check_t.Non, i64, False)
            m.reader_params.async_param = NS(is_dense_float=True, num_threads=2, num_batches_per_thread=2)
            m.reader_params.float_label_dense, m.reader_params.num_samples = True, 46
            m.reader_params.eval_num_samples = 5
            r = RawAsyncReader(m, True)
            it = 0
            while True:
                hb = r.read_a_batch()
                if hb is None:
                    break
                lo = it * b * world + rank * b
                n = hb.num_valid
                exp = raw[lo:lo + n]
                lab = exp[:, :4 * L].copy().view("<f4").reshape(n, L)
                assert np.array_equal(hb.label[:n].numpy(), lab)
                k0 = exp[:, 4 * (L + D):4 * (L + D) + nnz[0] * kt.itemsize].copy().view(kt).reshape(n, nnz[0])
                assert np.array_equal(hb.keys[:b * nnz[0]].view(b, nnz[0])[:n].numpy(), k0.astype("int64"))
                got += n
                it += 1
            r.stop()
        assert got == 46
    check()
