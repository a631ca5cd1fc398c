"""MLPerf-style logging callback (samples/dlrm/mlperf_logger/callbacks.py): emits :::MLLOG lines for
eval accuracy / epoch / run stop through the logger, plugs into solver.training_callbacks."""
from __future__ import annotations

import time

from ..model import TrainingCallback
from . import logger


class LoggingCallback(TrainingCallback):
    def __init__(self, auc_threshold: float, iter_per_epoch: float, batchsize: int):
        self.auc_threshold, self.iter_per_epoch, self.batchsize = auc_threshold, iter_per_epoch, batchsize
        self.minimum_training_time = 0
        self._t0 = None
        self._success = False

    def on_training_start(self):
        self._t0 = time.time()
        logger.perf_log("init_stop")
        logger.perf_log("run_start")
        logger.perf_log("epoch_start", 0, epoch_num=0.0)

    def on_eval_start(self, current_iter):
        logger.perf_log("eval_start", None, epoch_num=current_iter / self.iter_per_epoch)
        return False

    def on_eval_end(self, current_iter, eval_results):
        auc = eval_results.get("AUC", 0.0)
        ep = current_iter / self.iter_per_epoch
        logger.perf_log("eval_accuracy", auc, epoch_num=ep)
        logger.perf_log("eval_stop", None, epoch_num=ep)
        if not self._success and auc >= self.auc_threshold:
            self._success = True
        elapsed_min = (time.time() - self._t0) / 60.0 if self._t0 else 0.0
        return self._success and elapsed_min >= self.minimum_training_time

    def on_training_end(self, current_iter):
        logger.perf_log("epoch_stop", None, epoch_num=current_iter / self.iter_per_epoch)
        logger.perf_log("run_stop", None, status="success" if self._success else "aborted")
        logger.perf_log("train_samples", current_iter * self.batchsize)
