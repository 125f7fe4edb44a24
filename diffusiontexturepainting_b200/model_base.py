"""Façade contract of the reference server. model_base.py is a KEPT file of the reference (trt_inference/model_base.py:14-58):
when the reference's flat module is importable (the server runs with cwd = trt_inference/) its class is used
as is, so `isinstance(model, model_base.ConditionalInpainterBase)` holds inside the unmodified handler. Only when it is
absent (tests, bench, the GPU box) the five-method contract below stands in: handler.py relies on device(), resolution(),
set_brush(), generate() and the attribute .image."""
from abc import ABC, abstractmethod


def _reference_base():
    try:
        import model_base as ref  # the reference's flat module, if on sys.path
    except Exception:
        return None
    cls = getattr(ref, "ConditionalInpainterBase", None)
    need = ("device", "resolution", "set_brush", "generate_raw", "generate")
    return cls if isinstance(cls, type) and all(hasattr(cls, n) for n in need) else None


ConditionalInpainterBase = _reference_base()

if ConditionalInpainterBase is None:
    class ConditionalInpainterBase(ABC):  # noqa: F811
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
