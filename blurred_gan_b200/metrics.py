"""A *working* SWD metric for the train-step harness.

The reference's ``metrics.SWDMetric`` (``metrics.py:93-157``) cannot run: it reads the channel count as the image
resolution (``:119``), computes the "fake" descriptors from the real minibatch (``:131``) and its callback refers to
a non-existent attribute (``callbacks.py:194``) -- SURVEY Q11.  This class keeps its interface
(``update_state(reals, fakes)``, ``result()``, ``reset_states()``) on top of the one entry point that does work,
``sliced_wasserstein.API``, and takes CUDA tensors directly (no device -> host copy of the images).
"""
from __future__ import annotations

from . import sliced_wasserstein as sw


class SWDMetric:
    def __init__(self, image_shape, name: str = "swd"):
        """``image_shape`` = (C, H, W) of the NCHW minibatches that will be fed."""
        self.name = name
        self.image_shape = tuple(int(v) for v in image_shape)
        self.reset_states()

    def reset_states(self):
        self._reals, self._fakes = [], []

    def update_state(self, reals, fakes):
        assert tuple(reals.shape[1:]) == self.image_shape and tuple(fakes.shape[1:]) == self.image_shape
        self._reals.append(reals.detach())
        self._fakes.append(fakes.detach())

    def result(self):
        """[SWDx1e3 per resolution ..., average]; names from ``get_metric_names()``."""
        api = sw.API(self.image_shape)
        api.begin("reals")
        for r in self._reals:
            api.feed("reals", r)
        api.end("reals")
        api.begin("fakes")
        for f in self._fakes:
            api.feed("fakes", f)
        return api.end("fakes")

    def metric_names(self):
        return sw.API(self.image_shape).get_metric_names()
