"""Train one of the sample model families on synthetic data.

  python samples/model_zoo_train.py {ncf-gmf,ncf-mlp,ncf-neumf,mmoe,shared-bottom,din,bst,dcn,deepfm,wdl,criteo-dnn,dlrm-ftrl} [--iters N]

(reference: samples/ncf, samples/mmoe, samples/din, samples/bst, samples/dcn, samples/deepfm, samples/wdl)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hugectr_b200 import models  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("model")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--batchsize", type=int, default=1024)
a = ap.parse_args()
builders = {"ncf-gmf": lambda: models.build_ncf("gmf", batchsize=a.batchsize),
            "ncf-mlp": lambda: models.build_ncf("mlp", batchsize=a.batchsize),
            "ncf-neumf": lambda: models.build_ncf("neumf", batchsize=a.batchsize),
            "mmoe": lambda: models.build_mmoe(batchsize=a.batchsize),
            "din": lambda: models.build_din(batchsize=a.batchsize),
            "bst": lambda: models.build_bst(batchsize=a.batchsize),
            "dcn": lambda: models.build_dcn(batchsize=a.batchsize),
            "deepfm": lambda: models.build_deepfm(batchsize=a.batchsize),
            "wdl": lambda: models.build_wdl(batchsize=a.batchsize),
            "shared-bottom": lambda: models.build_shared_bottom(batchsize=a.batchsize),
            "criteo-dnn": lambda: models.build_criteo_dnn(batchsize=a.batchsize),
            "dlrm-ftrl": lambda: models.build_dlrm_ftrl(batchsize=a.batchsize, table_sizes=[100000] * 26)}
m = builders[a.model]()
m.compile()
m.summary()
m.fit(max_iter=a.iters, display=max(1, a.iters // 10), eval_interval=max(1, a.iters // 2), snapshot=0)
