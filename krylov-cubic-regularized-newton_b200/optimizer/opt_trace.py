"""Run log of an optimizer (mirror of the reference's optimizer/opt_trace.py:19-120, minus plotting).

Holds the checkpointed iterates ``xs``, wall-clock ``ts``, iteration counters ``its`` and, after
``compute_loss_of_iterates``, their loss values.  Iterates may be CUDA tensors; they are only handed
back to ``loss.value``.  Plotting needs matplotlib, which is optional here.
"""
from __future__ import annotations

import os
import pickle
import warnings
from pathlib import Path

import numpy as np


class Trace:
    def __init__(self, loss, label=None):
        self.loss, self.label = loss, label
        self.xs, self.ts, self.its = [], [], []
        self.loss_vals = []
        self.its_converted_to_epochs = False
        self.ls_its = None

    def compute_loss_of_iterates(self):
        if len(self.loss_vals) > 0:
            warnings.warn('Loss values have already been computed. Set .loss_vals = [] to recompute.')
            return
        batched = getattr(self.loss, "values_of_iterates", None)
        # the engine reads the matrix once per 8 stored iterates instead of once per iterate (kcrn_values_batch)
        self.loss_vals = batched(self.xs) if batched is not None else np.asarray([self.loss.value(x) for x in self.xs])

    def convert_its_to_epochs(self, batch_size=1):
        if self.its_converted_to_epochs:
            warnings.warn('The iteration count has already been converted to epochs.')
            return
        self.its = np.asarray(self.its) / (self.loss.n / batch_size)
        self.its_converted_to_epochs = True

    @property
    def best_loss_value(self):
        if len(self.loss_vals) == 0:
            self.compute_loss_of_iterates()
        return np.min(self.loss_vals)

    def _axis(self, its, use_ls_its, time):
        if its is not None:
            return its
        if use_ls_its and self.ls_its is not None:
            return self.ls_its
        return self.ts if time else self.its

    def plot_losses(self, its=None, f_opt=None, label=None, markevery=None, use_ls_its=True, time=False,
                    *args, **kwargs):
        import matplotlib.pyplot as plt  # optional dependency

        if len(self.loss_vals) == 0:
            self.compute_loss_of_iterates()
        f_opt = self.loss.f_opt if f_opt is None else f_opt
        markevery = max(1, len(self.loss_vals) // 20) if markevery is None else markevery
        plt.plot(self._axis(its, use_ls_its, time), self.loss_vals - f_opt, label=label or self.label,
                 markevery=markevery, *args, **kwargs)
        plt.ylabel(r'$f(x)-f^*$')

    def plot_distances(self, its=None, x_opt=None, label=None, markevery=None, use_ls_its=True, time=False,
                       *args, **kwargs):
        import matplotlib.pyplot as plt  # optional dependency

        if x_opt is None:
            x_opt = self.xs[-1] if self.loss.x_opt is None else self.loss.x_opt
        markevery = max(1, len(self.xs) // 20) if markevery is None else markevery
        dists = [self.loss.norm(x - x_opt) ** 2 for x in self.xs]
        plt.plot(self._axis(its, use_ls_its, time), dists, label=label or self.label, markevery=markevery,
                 *args, **kwargs)
        plt.ylabel(r'$\Vert x-x^*\Vert^2$')

    def save(self, file_name, path='./results/'):
        keep, self.loss = self.loss, None          # the oracle holds a device handle: never pickled
        xs = self.xs
        self.xs = [x.cpu().numpy() if hasattr(x, "cpu") else x for x in xs]
        try:
            Path(path).mkdir(parents=True, exist_ok=True)
            with open(os.path.join(path, file_name), 'wb') as f:
                pickle.dump(self, f)
        finally:
            self.loss, self.xs = keep, xs

    @classmethod
    def from_pickle(cls, path, loss=None):
        if not os.path.isfile(path):
            return None
        with open(path, 'rb') as f:
            trace = pickle.load(f)
        trace.loss = loss
        if loss is not None and len(trace.loss_vals) > 0:
            loss.f_opt = min(np.min(trace.loss_vals), loss.f_opt)
        return trace
