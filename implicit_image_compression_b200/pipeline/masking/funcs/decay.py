"""Prune-rate schedules (reference: pipeline/masking/funcs/decay.py).  Pure host arithmetic in Python
doubles, written so that every value equals what the reference's schedule objects return."""
import math


class Decay:
    mode = "current"

    def step(self, *args, **kwargs):
        raise NotImplementedError

    def get_dr(self):
        raise NotImplementedError


class CosineDecay(Decay):
    """decay.py:25-69.  The reference wraps torch's CosineAnnealingLR around a dummy SGD and drives it with
    an explicit epoch, which evaluates the closed form eta_min + (lr0 - eta_min)(1 + cos(pi t / T))/2."""

    def __init__(self, prune_rate=0.3, T_max=1000, eta_min=0.0, last_epoch=-1):
        self.mode = "current"
        self.base, self.T_max, self.eta_min = prune_rate, T_max, eta_min
        self._step = 0
        self._epoch = 0
        self._rate = prune_rate

    def _closed_form(self, epoch):
        return self.eta_min + (self.base - self.eta_min) * (1 + math.cos(math.pi * epoch / self.T_max)) / 2

    def step(self, step=-1):
        if step >= 0:
            if self._step < self.T_max:
                self._epoch = step
                self._rate = self._closed_form(step)
                self._step = step + 1
            else:
                self._step = self.T_max
            return
        if self._step < self.T_max:
            # recursive form of CosineAnnealingLR.get_lr(); only reached when no explicit step is given
            self._epoch += 1
            e, T = self._epoch, self.T_max
            if (e - 1 - T) % (2 * T) == 0:
                self._rate = self._rate + (self.base - self.eta_min) * (1 - math.cos(math.pi / T)) / 2
            else:
                self._rate = ((1 + math.cos(math.pi * e / T)) / (1 + math.cos(math.pi * (e - 1) / T))
                              * (self._rate - self.eta_min) + self.eta_min)
            self._step += 1

    def get_dr(self):
        return self._rate


class LinearDecay(Decay):
    """decay.py:72-108."""

    def __init__(self, prune_rate=0.3, T_max=1000):
        self.mode = "current"
        self._step = 0
        self.T_max = T_max
        self.decrement = prune_rate / float(T_max)
        self.current_prune_rate = prune_rate
        self.initial_prune_rate = prune_rate

    def step(self, step=-1):
        if step >= 0:
            if self._step < self.T_max:
                self.current_prune_rate = self.initial_prune_rate - self.decrement * (step + 1)
                self._step = step + 1
            else:
                self._step = self.T_max
            return
        if self._step < self.T_max:
            self.current_prune_rate -= self.decrement
            self._step += 1

    def get_dr(self):
        return self.current_prune_rate


class MagnitudePruneDecay(Decay):
    """decay.py:111-158 — Zhu & Gupta cubic cumulative-sparsity schedule; the prune rate is the finite
    difference between the target cumulative sparsity and the model's current sparsity."""

    def __init__(self, initial_sparsity=0.0, final_sparsity=0.3, T_max=30000, T_start=350, interval=100):
        self.mode = "cumulative"
        self.initial_sparsity, self.final_sparsity = initial_sparsity, final_sparsity
        self.T_max, self.T_start, self.interval = T_max, T_start, interval
        self.current_prune_rate = 0.0
        self._step = 0

    def cumulative_sparsity(self, step):
        if step < self.T_start:
            return self.initial_sparsity
        if step < self.T_max:
            mul = (1 - (step - self.T_start) / (self.T_max - self.T_start)) ** 3
            return self.final_sparsity + (self.initial_sparsity - self.final_sparsity) * mul
        return self.final_sparsity

    def step(self, step=-1, current_sparsity=-1):
        if step == -1:
            step = self._step
        if current_sparsity == -1:
            current_sparsity = self.cumulative_sparsity(step - self.interval)
        self.current_prune_rate = max(self.cumulative_sparsity(step) - current_sparsity, 0)
        self._step = step + 1

    def get_dr(self):
        return self.current_prune_rate


registry = {"cosine": CosineDecay, "linear": LinearDecay, "magnitude-prune": MagnitudePruneDecay}
