"""Inference session over a trained checkpoint (the API that survives as remnants in the reference:
InferenceParams / CreateInferenceSession / InferenceModel.predict, HugeCTR/include/inference/*;
HPS itself is removed upstream).  The session rebuilds the graph from its JSON, loads dense weights
and embedding tables and runs the eval network; embeddings may optionally be served through the
GPU hot-row cache (``use_gpu_embedding_cache`` + ``cache_size_percentage``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from .enums import Check_t, DataReaderType_t
from .model import Model
from .parallel.comm import Comm
from .solver import CreateSolver, DataReaderParams, OptParamsPy


@dataclass
class InferenceParams:
    model_name: str
    max_batchsize: int
    hit_rate_threshold: float = 1.0
    dense_model_file: str = ""
    sparse_model_files: List[str] = field(default_factory=list)
    device_id: int = 0
    use_gpu_embedding_cache: bool = False
    cache_size_percentage: float = 0.2
    i64_input_key: bool = False
    use_mixed_precision: bool = False
    embedding_collection_path: str = ""


class InferenceModel:
    def __init__(self, model_config_path: str, params: InferenceParams, device: Optional[str] = None):
        dev = torch.device(device) if device else (
            torch.device("cuda", params.device_id) if torch.cuda.is_available() else torch.device("cpu"))
        solver = CreateSolver(model_name=params.model_name, batchsize=params.max_batchsize,
                              batchsize_eval=params.max_batchsize, vvgpu=[[params.device_id]],
                              use_mixed_precision=params.use_mixed_precision,
                              i64_input_key=params.i64_input_key, use_cuda_graph=False)
        reader = DataReaderParams(DataReaderType_t.Parquet, source=["synthetic"], eval_source="synthetic",
                                  check_type=Check_t.Non)
        self.model = Model(solver, reader, OptParamsPy(), comm=Comm.single(dev))
        self.model.construct_from_json(model_config_path)
        self.model.compile()
        if params.dense_model_file:
            self.model.load_dense_weights(params.dense_model_file)
        self.hps = {}
        if params.sparse_model_files and params.use_gpu_embedding_cache and self.model.legacy_eval:
            # HPS-style serving: tables stay in the host parameter server, hot rows in the device cache
            from .cache import HpsEmbedding
            from .onnx.hugectr2onnx import load_sparse_model
            for rt, path in zip(self.model.legacy_eval, params.sparse_model_files):
                keys, emb = load_sparse_model(path, rt.vec)
                self.hps[rt.name] = HpsEmbedding(keys, emb, dev, params.cache_size_percentage,
                                                 "mean" if rt.combiner == 1 else "sum")
        elif params.sparse_model_files:
            self.model.load_sparse_weights(params.sparse_model_files)
        if params.embedding_collection_path:
            self.model.embedding_load(params.embedding_collection_path)
        self.params = params

    def predict(self, dense: np.ndarray, keys: np.ndarray) -> np.ndarray:
        """dense [n, dense_dim] float32, keys: feature-major key vector of the batch -> predictions"""
        m = self.model
        n = dense.shape[0]
        b = m.b_eval
        assert n <= b, "batch larger than max_batchsize"
        from .data.batch import HostBatch
        d = torch.zeros(b, m.input.dense_dim)
        d[:n] = torch.from_numpy(np.asarray(dense, dtype="float32"))
        k = torch.as_tensor(keys).to(m.key_dtype)
        hb = HostBatch(torch.zeros(b, m.input.label_dim), d, k, None, n)
        m._load_batch(hb, False)
        for e in m.ebcs_eval:
            e.forward(False)
        for rt in m.legacy_eval:
            if rt.name in self.hps:
                pooled = self.hps[rt.name].lookup(rt.keys_loc.view(rt.b, rt.S, rt.H))
                rt.top_data.copy_(pooled.to(rt.top_data.dtype))
            else:
                rt.forward(False)
        m.net_eval.fprop(False)
        ll = m.net_eval.loss_layers
        pred = ll[0].pred if len(ll) == 1 else torch.cat([l.pred.reshape(b, -1) for l in ll], 1)
        return pred[:n].float().cpu().numpy()


def CreateInferenceSession(model_config_path: str, inference_params: InferenceParams) -> InferenceModel:
    return InferenceModel(model_config_path, inference_params)
