"""hugectr2onnx: graph JSON + dense .model (+ sparse model dirs) -> inference graph -> ONNX."""
from .hugectr2onnx import InferenceGraph, convert, load_dense_weights, load_sparse_model  # noqa: F401
