"""``import hugectr2onnx`` -- drop-in name of the reference's converter package
(onnx_converter/hugectr2onnx), backed by hugectr_b200.onnx."""
from . import converter  # noqa: F401
