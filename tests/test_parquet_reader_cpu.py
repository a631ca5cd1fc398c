"""Parquet reader vs a brute-force oracle on randomised data sets: several files and row-group sizes, one-hot and
multi-hot (list) slots incl. empty and over-long bags, slot offsets, int32 / int64 key tensors, 1 - 3 ranks, an
incomplete last batch, 1 or 3 decode workers.  (A 300-seed sweep of the same check ran clean.)"""
import json
import os
import random
import tempfile
from types import SimpleNamespace as NS

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest
import torch

from hugectr_b200.data.parquet_reader import ParquetReader

def run(seed):
    rnd = random.Random(seed); rng = np.random.default_rng(seed)
    d = tempfile.mkdtemp()
    L, Dn = rnd.randint(1, 2), rnd.randint(0, 4)
    blocks = []   # (name, S, H, fixed)
    for bi in range(rnd.randint(1, 3)):
        S = rnd.randint(1, 4); H = rnd.choice([1, 1, 3])
        blocks.append((f"b{bi}", S, H, H == 1))
    nslots = sum(b[1] for b in blocks)
    files, rows_total = [], 0
    cols_all = {}
    nfiles = rnd.randint(1, 3)
    data = []
    for f in range(nfiles):
        n = rnd.randint(5, 40)
        cols, names = [], []
        lab = rng.integers(0, 2, (n, L)).astype("float32"); den = rng.random((n, Dn), dtype=np.float32)
        for i in range(L): cols.append(pa.array(lab[:, i])); names.append(f"l{i}")
        for i in range(Dn): cols.append(pa.array(den[:, i])); names.append(f"c{i}")
        cats = []
        for (nm, S, H, fx) in blocks:
            for s in range(S):
                if H == 1:
                    v = rng.integers(0, 1000, n).astype("int64"); cols.append(pa.array(v)); cats.append([[int(x)] for x in v])
                else:
                    cnt = rng.integers(0, H + 2, n)      # may exceed H (truncated) or be 0
                    lists = [rng.integers(0, 1000, c).astype("int64").tolist() for c in cnt]
                    cols.append(pa.array(lists, type=pa.list_(pa.int64()))); cats.append(lists)
                names.append(f"s{len(names)}")
        path = os.path.join(d, f"f{f}.parquet")
        pq.write_table(pa.Table.from_arrays(cols, names=names), path, row_group_size=rnd.choice([7, 16, 1000]))
        files.append(path); data.append((lab, den, cats, n)); rows_total += n
    json.dump({"file_stats": [], "labels": [{"index": i} for i in range(L)], "conts": [{"index": L + i} for i in range(Dn)],
               "cats": [{"index": L + Dn + i} for i in range(nslots)]}, open(os.path.join(d, "_metadata.json"), "w"))
    fl = os.path.join(d, "fl.txt"); open(fl, "w").write(f"{nfiles}\n" + "\n".join(files) + "\n")
    world = rnd.choice([1, 2, 3]); b = rnd.randint(2, 9); i64 = rnd.random() < 0.5
    use_off = rnd.random() < 0.5
    ssa = [1000] * nslots if use_off else []
    outs = []
    for rank in range(world):
        layout = NS(blocks=blocks, total_slots=nslots)
        model = NS(reader_params=NS(source=[fl], eval_source=fl, slot_size_array=ssa, num_workers=rnd.choice([1, 3])),
                   b_train=b, b_eval=b, comm=NS(rank=rank), world=world, input=NS(label_dim=L, dense_dim=Dn), layout=layout,
                   key_dtype=torch.int64 if i64 else torch.int32, solver=NS(repeat_dataset=False, drop_incomplete_batch=False),
                   sparse_embeddings=[1] if use_off else [], device=torch.device("cpu"))
        r = ParquetReader(model, True); r.start()
        got = []
        while True:
            hb = r.read_a_batch()
            if hb is None: break
            got.append((hb, r.current_batchsize))
        outs.append(got)
    # oracle
    lab = np.concatenate([x[0] for x in data]); den = np.concatenate([x[1] for x in data])
    cats = [sum((x[2][s] for x in data), []) for s in range(nslots)]
    gb = b * world
    nb = (rows_total + gb - 1) // gb
    for rank in range(world):
        assert len(outs[rank]) == nb, (seed, len(outs[rank]), nb)
        for bi, (hb, nvalid) in enumerate(outs[rank]):
            lo = bi * gb + rank * b; hi = min(lo + b, min((bi + 1) * gb, rows_total)); nloc = max(0, hi - lo)
            assert hb.num_valid == nloc, (seed, "nvalid")
            assert np.allclose(hb.label[:nloc].numpy(), lab[lo:hi]) and float(hb.label[nloc:].abs().sum()) == 0
            if Dn: assert np.allclose(hb.dense[:nloc].numpy(), den[lo:hi])
            ko = no = si = 0
            keys = hb.keys.numpy(); nnz = hb.nnz.numpy()
            for (nm, S, H, fx) in blocks:
                blk = keys[ko:ko + b * S * H].reshape(b, S, H); nz = nnz[no:no + S * b].reshape(S, b)
                for s in range(S):
                    off = (1000 * (si + s)) if use_off else 0
                    for i in range(b):
                        if i < nloc:
                            want = [k + off for k in cats[si + s][lo + i][:H]]
                        else:
                            want = []
                        assert blk[i, s, :len(want)].tolist() == want, (seed, "keys", nm, s, i)
                        assert (blk[i, s, len(want):] == -1).all(), (seed, "pad")
                        assert nz[s, i] == len(want), (seed, "nnz", nz[s, i], len(want))
                si += S; ko += b * S * H; no += S * b
    return True


@pytest.mark.parametrize("seed", list(range(24)))
def test_parquet_reader_matches_oracle(seed):
    assert run(seed)
