"""Compatibility check: exec the reference's sample training scripts UNMODIFIED against the ``hugectr``
drop-in module (synthetic data, 1 CPU/GPU device, a handful of iterations).  Development tool only --
it reads the scripts from /root/reference at run time and does not copy them.

    python tools_dev/run_reference_samples.py [pattern]
"""
import glob
import os
import runpy
import sys
import tempfile
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["HCTR_FORCE_SYNTHETIC"] = "1"
os.environ.setdefault("HUGECTR_LOG_LEVEL", "1")

import types  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Comm:
    def Get_rank(self): return 0
    def Get_size(self): return 1
    def Barrier(self): pass
    def bcast(self, x, root=0): return x
    def allreduce(self, x, op=None): return x


# the samples import mpi4py only to initialise MPI before hugectr; torch.distributed does that here
_stub("mpi4py", MPI=types.SimpleNamespace(COMM_WORLD=_Comm(), SUM=None))

class _Any:
    """accepts any attribute access / call (mlperf logging stand-in)"""
    def __init__(self, *a, **kw): pass
    def __getattr__(self, k): return _Any()
    def __call__(self, *a, **kw): return _Any()
    def __str__(self): return "stub"


_stub("mlperf_logging", mllog=_stub("mlperf_logging.mllog", constants=_stub("mlperf_logging.mllog.constants")))
sys.modules["mlperf_logging.mllog.constants"].__getattr__ = lambda k: k.lower()
sys.modules["mlperf_logging.mllog"].get_mllogger = lambda: _Any()
sys.modules["mlperf_logging.mllog"].config = lambda **kw: None
sys.modules["mlperf_logging"].__path__ = []
sys.modules["mlperf_logging.mllog"].__path__ = []
_stub("mlperf_common").__path__ = []
_stub("mlperf_common.frameworks").__path__ = []
_stub("mlperf_common.frameworks.hugectr", HCTRCommunicationHandler=_Any)
_stub("mlperf_common.logging", MLLoggerWrapper=_Any)

import hugectr  # noqa: E402
import hugectr_b200  # noqa: E402

_solver = hugectr_b200.CreateSolver


def _small_solver(*a, **kw):
    world = int(os.environ.get("WORLD_SIZE", "1"))      # under torchrun: one list entry per rank
    kw["vvgpu"] = [list(range(world))]
    kw["batchsize"] = min(int(kw.get("batchsize", 2048)), 256)
    kw["batchsize_eval"] = min(int(kw.get("batchsize_eval", 2048)), 256)
    kw["max_eval_batches"] = 2
    return _solver(*a, **kw)


CAP = 3000
_etc = hugectr_b200.EmbeddingTableConfig


def _small_table(name=None, max_vocabulary_size=None, *a, **kw):
    if max_vocabulary_size is not None and max_vocabulary_size > CAP:
        max_vocabulary_size = CAP
    return _etc(name, max_vocabulary_size, *a, **kw)


_drp = hugectr_b200.DataReaderParams


def _small_reader(*a, **kw):
    if kw.get("slot_size_array"):
        kw["slot_size_array"] = [min(int(v), CAP) for v in kw["slot_size_array"]]
    return _drp(*a, **kw)


for mod in (hugectr, hugectr_b200):
    mod.EmbeddingTableConfig = _small_table
    mod.DataReaderParams = _small_reader

_getsize = os.path.getsize
os.path.getsize = lambda p: (1 << 30) if str(p).startswith(("/data/", "/data_val/")) else _getsize(p)   # train.py sizes its dataset

_shard = hugectr_b200.EmbeddingCollectionConfig.shard


def _one_gpu_shard(self, shard_matrix, shard_strategy, *a, **kw):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if len(shard_matrix) != world and world > 1:      # planned for another GPU count: round-robin the tables
        names = sorted({t for row in shard_matrix for t in row}, key=int) if isinstance(shard_matrix[0][0], str) \
            else [str(i) for i in range(len(shard_matrix[0]))]
        shard_matrix = [[n for j, n in enumerate(names) if j % world == r] for r in range(world)]
        shard_strategy = [("mp", names)]
    elif len(shard_matrix) > 1 and world == 1:   # the script planned for its own GPU count: fold onto the one device
        if isinstance(shard_matrix[0][0], str):
            shard_matrix = [sorted({t for row in shard_matrix for t in row}, key=int)]
        else:
            shard_matrix = [[int(any(col)) for col in zip(*shard_matrix)]]
    return _shard(self, shard_matrix, shard_strategy, *a, **kw)


hugectr_b200.EmbeddingCollectionConfig.shard = _one_gpu_shard

def _remap(path):
    """absolute output paths of the test scripts (/onnx_converter/..., /dump_infer/...) go under the cwd"""
    path = str(path)
    if path.startswith("/") and not path.startswith(("/tmp", "/root/repo", "/dev/shm")):
        path = os.path.join(os.getcwd(), path.lstrip("/"))
        os.makedirs(os.path.dirname(path), exist_ok=True)
    return path


import numpy as _np  # noqa: E402
_npsave = _np.save
_np.save = lambda f, *a, **kw: _npsave(_remap(f) if isinstance(f, str) else f, *a, **kw)

_g2j = hugectr_b200.Model.graph_to_json
hugectr_b200.Model.graph_to_json = lambda self, graph_config_file, *a, **kw: _g2j(self, _remap(graph_config_file), *a, **kw)

_fit = hugectr_b200.Model.fit


def _short_fit(self, *a, **kw):
    kw = dict(kw)
    snap = 6 if kw.get("snapshot", 1000000) <= kw.get("max_iter", 0) else 1000000   # keep the final snapshot
    kw.update(max_iter=6, display=3, eval_interval=3, snapshot=snap, num_epochs=0)
    if "snapshot_prefix" in kw:
        kw["snapshot_prefix"] = _remap(kw["snapshot_prefix"])
    return _fit(self, **kw)


for mod in (hugectr, hugectr_b200):
    mod.CreateSolver = _small_solver
hugectr_b200.Model.fit = _short_fit


EXTRA_ARGS = {
    "train.py": ["--batchsize", "256", "--batchsize_eval", "256", "--max_iter", "6", "--eval_interval", "3",
                 "--max_eval_batches", "2", "--num_gpus_per_node", "1", "--ev_size", "16",
                 "--memory_cap_for_embedding", "1"],
    "dlrm_train_ftrl.py": ["--shard_plan", "hybrid", "--optimizer", "ftrl"],
    "din_fp32.py": ["--vvgpu", "0"],
    "hugectr_e2e_demo_with_nvtabular__18.py": ["--data_path", "./data", "--model_path", "./model"],
    "benchmarks/embedding_collection/hugectr/train.py": [
        "--batchsize", "256", "--batchsize_eval", "128", "--num_gpus_per_node", "1", "--max_iter", "6",
        "--eval_interval", "3", "--max_eval_batches", "2", "--ev_size_per_table", "16",
        "--vocabulary_size_per_table", ",".join(["3000"] * 26), "--memory_cap_for_embedding", "1"],
    "embedding_collection__dlrm_train.py": ["--shard_plan", "hybrid"],
    "din_matmul_fp32_1gpu.py": ["--vvgpu", "0"],
}


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    if "*" in pat or os.path.isfile(pat):                       # explicit glob, e.g. '/root/reference/test/pybind_test/*.py'
        scripts, pat = sorted(glob.glob(pat)), ""
    else:
        scripts = sorted(glob.glob("/root/reference/samples/*/*.py"))
    scripts = [s for s in scripts if pat in s and "preprocess" not in s and "/sharding/" not in s
               and not s.endswith("utils.py")]
    res = {}
    runs = []
    for s in scripts:
        if s.endswith("pybind_test/model_test.py"):
            for mt in ["CRITEO", "DCNV1", "DCNV2", "DEEPFM", "WDL", "BST"]:
                runs.append((s, ["--model_type", mt, "--vvgpu", "0", "--max_iter", "6", "--auc_threshold", "0", "--eval_interval", "2", "--display", "2"], f" [{mt}]"))
        else:
            extra = EXTRA_ARGS.get(os.path.basename(s), [])
            for k, v in EXTRA_ARGS.items():
                if "/" in k and s.endswith(k):
                    extra = v
            runs.append((s, extra, ""))
    for s, extra, tag in runs:
        cwd = tempfile.mkdtemp()
        os.chdir(cwd)
        sys.argv = [s] + extra
        sys.path.insert(0, os.path.dirname(s))
        try:
            runpy.run_path(s, run_name="__main__")
            res[s + tag] = "OK"
        except SystemExit as e:
            res[s + tag] = f"exit {e.code}"
        except BaseException as e:  # noqa: BLE001
            tb = traceback.extract_tb(e.__traceback__)
            res[s + tag] = f"{type(e).__name__}: {str(e)[:300]}  @ {tb[-1].filename.split('/')[-1]}:{tb[-1].lineno}"
    if int(os.environ.get("RANK", "0")) != 0:
        return
    print("\n==== summary ====")
    for s, r in res.items():
        print(f"{s.replace('/root/reference/', ''):60s} {r}")


if __name__ == "__main__":
    main()
