"""Façade contract of the reference server (trt_inference/model_base.py:14-58), restated so the package is importable
without the reference tree. handler.py only relies on device(), resolution(), set_brush(), generate(), .image."""
from abc import ABC, abstractmethod


class ConditionalInpainterBase(ABC):
    @abstractmethod
    def device(self):
        ...

    @abstractmethod
    def resolution(self):
        ...

    @abstractmethod
    def set_brush(self, conditioning):
        ...

    @abstractmethod
    def generate_raw(self, canvas, **settings):
        ...

    def generate(self, canvas, **settings):
        result = self.generate_raw(canvas, **settings)
        alpha = canvas[:, 3:, ...]
        return canvas[:, :3, ...] * alpha + result[:, :3, ...] * (1 - alpha)
