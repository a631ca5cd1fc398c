"""Layer coverage tables of the converter, under the names the reference's coverage test imports
(onnx_converter/hugectr2onnx/hugectr_loader.py:22-56).  Computed from the layer dispatch of
``hugectr_b200.onnx.hugectr2onnx.InferenceGraph`` rather than maintained by hand."""
import inspect
import re

from hugectr_b200.enums import Layer_t
from hugectr_b200.onnx import hugectr2onnx as _impl

_handled = set(re.findall(r'"([A-Z][A-Za-z0-9_]+)"', inspect.getsource(_impl.InferenceGraph)))
ONNX_LAYER_TYPES = {n for n in Layer_t.__members__ if n in _handled} | {"Data", "DistributedSlotSparseEmbeddingHash",
                                                                        "LocalizedSlotSparseEmbeddingHash"}
# recurrent layer: exempted by the reference converter as well
EXEMPTION_LAYER_TYPES = {n for n in Layer_t.__members__ if n not in _handled}
