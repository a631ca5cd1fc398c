"""gpu_cache (HBM hot-row LRU cache, static table, UVM table) and the host parameter server /
embedding training cache tier."""
from .gpu_cache import GpuCache, StaticTable, UvmTable  # noqa: F401
from .hps import (EmbeddingTrainingCache, HMemCache, HostParameterServer, HpsEmbedding, OffloadedEmbedding,  # noqa: F401
                  SparseModelFile)
