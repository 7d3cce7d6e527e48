"""Step loop, timing and checkpointing (mirror of the reference's optimizer/optimizer.py:17-172).

``run`` keeps the reference's contract: iterate ``step()`` until the iteration budget, the time
budget or ``|x - x_prev| < tolerance`` stops it; ``trace.ts[k]`` is wall-clock since ``init_run`` at
the end of iteration k (optimizer.py:129,151), taken after the device work of the step has been
synchronised (every step ends with one device-to-host read of its scalars).
"""
from __future__ import annotations

import copy
import time

import numpy as np

from optimizer.opt_trace import Trace

try:  # progress bar exactly where the reference has it (optimizer.py:85-98); optional here
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None


def _clone(x):
    return x.clone() if hasattr(x, "clone") else copy.deepcopy(x)


class _NoBar:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def update(self, *_):
        pass


class Optimizer:
    """Base class of the optimization algorithms: run loop, trace, stopping rules."""

    def __init__(self, loss, trace_len=200, use_prox=True, tolerance=0, line_search=None,
                 save_first_iterations=5, label=None, seeds=None, tqdm=True):
        self.loss = loss
        self.trace_len = trace_len
        self.use_prox = use_prox and (self.loss.regularizer is not None)
        self.tolerance = tolerance
        self.line_search = line_search
        self.save_first_iterations = save_first_iterations
        self.label = label
        self.tqdm = tqdm
        self.initialized = False
        self.x_old_tol = None
        self.trace = Trace(loss=loss, label=label)
        self.seeds = [42] if seeds is None else seeds
        self.finished_seeds = []

    # ------------------------------------------------------------------ the loop
    def run(self, x0, t_max=np.inf, it_max=np.inf, ls_it_max=None):
        if t_max is np.inf and it_max is np.inf:
            it_max = 100
            print(f'{self.label}: The number of iterations is set to {it_max}.')
        self.t_max, self.it_max = t_max, it_max
        for seed in self.seeds:
            if seed in self.finished_seeds:
                continue
            if len(self.seeds) > 1:
                print(f'{self.label}: Running seed {seed}')
            self.rng = np.random.default_rng(seed)
            if ls_it_max is None:
                self.ls_it_max = it_max
            if not self.initialized:
                self.init_run(x0)
                self.initialized = True
            by_iterations = self.ls_it_max is not np.inf
            bar = _tqdm(total=self.ls_it_max if by_iterations else self.t_max) if (self.tqdm and _tqdm) else _NoBar()
            shown = 0
            with bar:
                while not self.check_convergence():
                    if self.tolerance > 0:
                        self.x_old_tol = _clone(self.x)
                    self.step()
                    self.save_checkpoint()
                    now = (self.ls_it if self.line_search is not None else self.it) if by_iterations else self.t
                    bar.update(now - shown)
                    shown = now
            self.finished_seeds.append(seed)
            self.initialized = False
        return self.trace

    def step_length(self):
        """|x - x_prev| of the last step; subclasses may return it from the device record."""
        return self.loss.norm(self.x - self.x_old_tol)

    def check_convergence(self):
        out_of_its = self.it >= self.it_max
        if self.line_search is not None:
            out_of_its = out_of_its or (self.line_search.it >= self.ls_it_max)
        out_of_time = time.perf_counter() - self.t_start >= self.t_max
        small_step = (self.tolerance > 0 and self.x_old_tol is not None
                      and self.step_length() < self.tolerance)
        return out_of_its or out_of_time or small_step

    def step(self):
        pass

    def init_run(self, x0):
        self.dim = x0.shape[0]
        self.x = _clone(x0)
        self.trace.xs, self.trace.its, self.trace.ts = [_clone(x0)], [0], [0]
        if self.line_search is not None:
            self.trace.ls_its, self.trace.lrs = [0], [self.line_search.lr]
        self.it, self.t = 0, 0
        self.t_start = time.perf_counter()
        self.time_progress = self.iterations_progress = self.max_progress = 0
        if self.line_search is not None:
            self.line_search.reset(self)

    def should_update_trace(self):
        if self.it <= self.save_first_iterations:
            return True
        budget = self.trace_len - self.save_first_iterations
        self.time_progress = int(budget * self.t / self.t_max)
        self.iterations_progress = int(budget * (self.it / self.it_max))
        if self.line_search is not None:
            self.iterations_progress = max(self.iterations_progress,
                                           int(budget * (self.line_search.it / self.it_max)))
        return max(self.time_progress, self.iterations_progress) > self.max_progress

    def save_checkpoint(self):
        self.it += 1
        if self.line_search is not None:
            self.ls_it = self.line_search.it
        self.t = time.perf_counter() - self.t_start
        if self.should_update_trace():
            self.update_trace()
        self.max_progress = max(self.time_progress, self.iterations_progress)

    def update_trace(self):
        self.trace.xs.append(_clone(self.x))
        self.trace.ts.append(self.t)
        self.trace.its.append(self.it)
        if self.line_search is not None:
            self.trace.ls_its.append(self.line_search.it)
            self.trace.lrs.append(self.line_search.lr)

    def compute_loss_of_iterates(self):
        self.loss.reset()
        self.trace.compute_loss_of_iterates()

    def reset(self, loss):
        self.initialized = False
        self.x_old_tol = None
        self.trace = Trace(loss=loss, label=self.label)
        self.finished_seeds = []
