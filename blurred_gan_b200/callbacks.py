"""Sigma controllers of the reference (``callbacks.py:45-135``) for the torch train-step harness.

Host-side control only: they decide the next ``std`` and write it into ``model.std`` before a batch
(``model.assign_std``).  The blur layers keep that Parameter in host memory -- the radius fixes the launch geometry, and
reading sigma back from the device every step would serialise the stream.
"""
from __future__ import annotations


class BlurDecayController:
    """``callbacks.py:45-62``.  ``std(n_batches) = max_value * 0.96 ** (n_batches / (total_n_training_examples / 10))``.

    As in the reference the schedule is evaluated at the *batch* counter although ``decay_steps`` is given in examples
    (over a 10-epoch run sigma only falls to 0.987 * max_value), and ``min_value`` is accepted but unused."""

    def __init__(self, total_n_training_examples: int, max_value: float = 23.5, min_value: float = 0.01):
        self.initial = float(max_value)
        self.decay_steps = total_n_training_examples / 10
        self.decay_rate = 0.96
        self.min_value = min_value
        self.model = None

    def set_model(self, model):
        self.model = model

    def schedule(self, step: int) -> float:
        return self.initial * self.decay_rate ** (step / self.decay_steps)

    def on_batch_begin(self, batch=None, logs=None):
        self.model.assign_std(self.schedule(self.model.n_batches))


class LinearBlurSchedule:
    """The driver's synthetic schedule of BASELINE configs[0]: sigma annealed linearly ``max_value -> min_value`` over
    ``total_batches`` (not in the reference, whose own schedule barely anneals; used by the harness and its tests)."""

    def __init__(self, total_batches: int, max_value: float = 23.0, min_value: float = 0.0):
        self.total, self.max_value, self.min_value = max(int(total_batches) - 1, 1), float(max_value), float(min_value)
        self.model = None

    def set_model(self, model):
        self.model = model

    def schedule(self, step: int) -> float:
        t = min(max(step / self.total, 0.0), 1.0)
        return self.max_value + (self.min_value - self.max_value) * t

    def on_batch_begin(self, batch=None, logs=None):
        self.model.assign_std(self.schedule(self.model.n_batches))


class AdaptiveBlurController:
    """``callbacks.py:65-135``: lowers sigma by ``smoothing`` whenever the smoothed score ratio
    ``fake / (real + fake)`` sits within ``threshold`` of 0.5 and sigma was not touched for 100 batches.

    In the reference the assignment to the layer is commented out (``callbacks.py:102-103``), so the controller only
    tracks the sigma it *would* use; ``apply=True`` enables the assignment."""

    def __init__(self, smoothing=0.99, warmup_n_batches=100, threshold=0.05, min_value=0.01, max_value=23.5, apply=False):
        self.smoothing = smoothing
        self.warmup_n_batches = warmup_n_batches
        self.score_ratio = 0.5
        self.threshold = threshold
        self._last_modification_step = 0
        self.delay_between_modifications = 100
        self.std = float(max_value)
        self.min_value = min_value
        self.apply = apply
        self.model = None
        self.stop_training = False
        self.would_modify = 0

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        self.model.assign_std(self.std)

    def gan_problem_is_stable(self) -> bool:
        return 0.5 - self.threshold <= self.score_ratio <= 0.5 + self.threshold

    def decrease_blur_std(self, batch: int) -> None:
        std_was_just_modified = batch - self._last_modification_step < self.delay_between_modifications
        if not std_was_just_modified:
            self.std = self.smoothing * self.std
            if self.apply:
                self.model.assign_std(self.std)
            self.would_modify = 1
            self._last_modification_step = batch
        else:
            self.would_modify = 0

    def on_batch_end(self, batch, logs):
        fake_scores, real_scores = logs["fake_scores"], logs["real_scores"]
        ratio = fake_scores / (real_scores + fake_scores)
        self.score_ratio = self.smoothing * self.score_ratio + (1 - self.smoothing) * ratio
        if batch < self.warmup_n_batches:
            return
        if self.gan_problem_is_stable():
            self.decrease_blur_std(batch)
        if self.std < self.min_value:
            self.stop_training = True
            if self.model is not None:
                self.model.stop_training = True  # what the reference sets (callbacks.py:134) and a Keras-style loop checks
