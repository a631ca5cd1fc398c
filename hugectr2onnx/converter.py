"""``hugectr2onnx.converter.convert(onnx_model_path, graph_config, dense_model, convert_embedding=False,
sparse_models=None, ntp_file=None, graph_name="hugectr")`` (reference:
onnx_converter/hugectr2onnx/converter.py:22-47)."""
from hugectr_b200.onnx.hugectr2onnx import convert  # noqa: F401
