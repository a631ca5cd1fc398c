"""``import hugectr`` -- drop-in module name of the reference Python API, backed by hugectr_b200."""
import sys as _sys

import hugectr_b200 as _impl
from hugectr_b200 import *  # noqa: F401,F403
from hugectr_b200 import (CreateOptimizer, CreateSolver, DataReaderParams, DataReaderSparseParam,  # noqa: F401
                          DenseLayer, DenseLayerComputeConfig, EmbeddingCollectionConfig,
                          EmbeddingTableConfig, Input, Model, SparseEmbedding, TrainingCallback,
                          AsyncParam, LearningRateScheduler)


def __getattr__(name):
    return getattr(_impl, name)
