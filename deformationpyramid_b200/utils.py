"""Host-side helpers the reference's scripts take from its utils/ package (kept by the reference callers as they are;
restated here so that the package is usable without a reference checkout): wall-clock Timers with the
tic / toc / tictoc / get_strings surface of utils/tiktok.py:10-76, AverageMeter and Logger of utils/utils.py:2-34,
setup_seed of utils/benchmark_utils.py:9-18."""
from __future__ import annotations

import os
import random
import time
from collections import defaultdict

import numpy as np
import torch


class Timer:
    def __init__(self):
        self.total_time, self.calls, self.start_time, self.diff = 0.0, 0, 0.0, 0.0

    def tic(self):
        self.start_time = time.time()

    def toc(self, average=True):
        self.tictoc(time.time() - self.start_time)

    def tictoc(self, diff):
        self.diff = diff
        self.total_time += diff
        self.calls += 1

    def total(self):
        return self.total_time

    def avg(self):
        return self.total_time / float(self.calls)


class Timers:
    def __init__(self):
        self.timers = defaultdict(Timer)

    def tic(self, key):
        self.timers[key].tic()

    def toc(self, key):
        self.timers[key].toc()

    def tictoc(self, key, diff):
        self.timers[key].tictoc(diff)

    def get_avg(self, key):
        return self.timers[key].avg()

    def get_strings(self):
        return ["{:}: \t  average {:.4f},  total {:.4f} ,\t calls {:}".format(k.ljust(30), v.avg(), v.total_time, v.calls)
                for k, v in self.timers.items()]


class AverageMeter:
    def __init__(self):
        self.val, self.avg, self.sum, self.sq_sum, self.count = 0, 0, 0.0, 0.0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
        self.sq_sum += val ** 2 * n


class Logger:
    def __init__(self, log_path):
        if os.path.exists(log_path):
            os.remove(log_path)
        self.fw = open(log_path, "a")

    def write(self, text):
        self.fw.write(text)
        self.fw.flush()

    def close(self):
        self.fw.close()


def setup_seed(seed):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
