"""Learning-rate schedule: warmup -> constant -> polynomial decay -> end_lr.

Parity: HugeCTR/include/learning_rate_scheduler.hpp:62-83; GPU twin in csrc/dense_ops.cu
(``lr_step_kernel``), reference HugeCTR/src/gpu_learning_rate_scheduler.cu:26.
"""
from __future__ import annotations


def lr_at(step: int, base_lr: float, warmup_steps: int = 1, decay_start: int = 0,
          decay_steps: int = 1, decay_power: float = 2.0, end_lr: float = 0.0) -> float:
    if step <= warmup_steps:
        return base_lr * step / max(1, warmup_steps)
    if decay_start == 0 or step <= decay_start:
        return base_lr
    if step <= decay_start + decay_steps:
        f = (decay_start + decay_steps - step) / float(decay_steps)
        return max(base_lr * (f ** decay_power), end_lr)
    return end_lr


class LearningRateScheduler:
    def __init__(self, base_lr: float, warmup_steps: int = 1, decay_start: int = 0,
                 decay_steps: int = 1, decay_power: float = 2.0, end_lr: float = 0.0):
        # same acceptance as the reference (learning_rate_scheduler.hpp:40-46): 0 warm-up / decay steps
        # are legal (samples/dlrm/train.py runs with warmup_steps=0)
        if base_lr < 0 or warmup_steps < 0 or decay_steps < 0 or decay_power < 1.0 or end_lr < 0:
            raise ValueError("base_lr < 0 || warmup_steps < 0 || decay_steps < 0 || decay_power < 1.0 || end_lr < 0")
        self.base_lr, self.warmup_steps = base_lr, warmup_steps
        self.decay_start, self.decay_steps = decay_start, decay_steps
        self.decay_power, self.end_lr = decay_power, end_lr
        self.step = 0

    def get_next(self) -> float:
        self.step += 1
        return lr_at(self.step, self.base_lr, self.warmup_steps, self.decay_start,
                     self.decay_steps, self.decay_power, self.end_lr)

    def get_lr(self) -> float:
        return lr_at(max(self.step, 1), self.base_lr, self.warmup_steps, self.decay_start,
                     self.decay_steps, self.decay_power, self.end_lr)

    def reset(self):
        self.step = 0

    def state_dict(self):
        return {"step": self.step}

    def load_state_dict(self, d):
        self.step = int(d["step"])
