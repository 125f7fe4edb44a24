"""The drop-in claim, exercised: the reference's UNMODIFIED handler.py / server_io.py / run.py / model_base.py are imported
from /root/reference/trt_inference with dropin/ ahead of them on sys.path, tornado / flask / kornia
stubbed in sys.modules (not installed here; the handler only needs WebSocketHandler as a base class and gen.coroutine).

CPU (this file): run.py / handler.py resolve `trt_model` to this repository's class; real NEW_BRUSH / NEW_STAMP frames go
through InpaintWebSocketHandler._handle_binary_request on a recording model; tests/handler_twin.py (the restatement the GPU
box uses, where the reference tree does not exist) must make the same calls and write the same bytes.
GPU: tests/test_gpu_dropin.py drives the real CUDA model through the same sequence and checks the bytes against the oracle."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import handler_twin as twin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(os.environ.get("DTP_REFERENCE", "/root/reference"), "trt_inference")
FLAT = ("handler", "server_io", "run", "model_base", "trt_model", "models", "image_encoder", "inpaint_pipeline",
        "stable_diffusion_pipeline", "websocket_model")
STUBS = ("tornado", "tornado.websocket", "tornado.gen", "tornado.wsgi", "tornado.web", "tornado.ioloop", "flask", "kornia",
         "kornia.morphology")


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class WebSocketHandler:  # tornado.websocket.WebSocketHandler: the handler only overrides hooks and calls write_message
        def __init__(self, *a, **kw):
            self.written = []
            self.initialize(**kw)

        def write_message(self, message, binary=False):
            self.written.append((bytes(message), binary))

    class _Any:
        def __init__(self, *a, **kw):
            self.args, self.kw = a, kw

    t = mod("tornado")
    t.websocket = mod("tornado.websocket", WebSocketHandler=WebSocketHandler)
    t.gen = mod("tornado.gen", coroutine=lambda f: f)  # the handlers never yield: run them synchronously
    t.wsgi = mod("tornado.wsgi", WSGIContainer=_Any)
    t.web = mod("tornado.web", Application=_Any, FallbackHandler=_Any)
    t.ioloop = mod("tornado.ioloop", IOLoop=_Any)
    mod("flask", Flask=_Any, render_template=None, Response=None)

    def no_kornia(*a, **kw):
        raise AssertionError("the drop-in must not call kornia: add_extra_context runs in dtp_op_canvas_preprocess")
    k = mod("kornia")
    k.morphology = mod("kornia.morphology", dilation=no_kornia)


@pytest.fixture()
def reference_server():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present (GPU box): covered by tests/test_gpu_dropin.py through handler_twin")
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in FLAT + STUBS}
    for k in FLAT:
        sys.modules.pop(k, None)
    install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "dropin"))  # ahead of the reference: replaces exactly the five named modules
    try:
        import handler
        import run
        import server_io
        yield handler, run, server_io
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_unmodified_server_modules_resolve_to_this_repo(reference_server):
    handler, run, server_io = reference_server
    import diffusiontexturepainting_b200.trt_model as ours
    assert handler.__file__.startswith(REF) and run.__file__.startswith(REF) and server_io.__file__.startswith(REF)
    assert run.TRTConditionalInpainter is ours.TRTConditionalInpainter
    assert sys.modules["trt_model"].__file__.startswith(os.path.join(ROOT, "dropin"))
    assert sys.modules["model_base"].__file__.startswith(REF)
    # the facade derives from the reference's own contract class when that is importable (model_base.py is a kept file)
    assert issubclass(ours.TRTConditionalInpainter, handler.ConditionalInpainterBase) or \
        {"device", "resolution", "set_brush", "generate_raw", "generate"} <= set(dir(ours.TRTConditionalInpainter))


class Recorder:
    """Stands in for the model on the CPU: records what the handler hands over and returns a deterministic image."""

    def __init__(self, res):
        self.res, self.calls, self.image = res, [], None

    def device(self):
        return "cpu"

    def resolution(self):
        return self.res

    def set_brush(self, image):
        self.calls.append(("set_brush", str(image.dtype), tuple(image.shape), str(image.device), float(image.sum())))
        self.image = image[None, :, :self.res, :self.res].clone()

    def generate(self, canvas, **settings):
        self.calls.append(("generate", str(canvas.dtype), tuple(canvas.shape), str(canvas.device), float(canvas.sum()),
                           {k: (type(v).__name__, float(v)) for k, v in sorted(settings.items())}))
        return canvas[:, :3].flip(-1) * 0.75 + 0.125


def test_handler_twin_is_the_unmodified_handler(reference_server):
    handler, _, server_io = reference_server
    R = 64
    frames = twin.synthetic_frames(R, steps=7, context_pad=150, cfg_weight=2.5, tg_weight=0.5, tg_steps=3)
    # the frames are what the reference's own client-side encoders produce
    canvas = twin.binary_to_image(frames[1], 14)
    ref_frame = (server_io.encode_request_type(server_io.RequestType.NEW_STAMP)
                 + server_io.encode_inference_settings(steps=7, width=R, context_pad=150, cfg_weight=2.5, tg_weight=0.5,
                                                       tg_steps=3) + server_io.image_to_binary(canvas))
    assert ref_frame == frames[1]
    a, b = Recorder(R), Recorder(R)
    h = handler.InpaintWebSocketHandler(model=a, model_info_str="trt", debug_dir=None)
    for f in frames:
        h._handle_binary_request(f)
    mine = [twin.handle_binary_request(b, f) for f in frames]
    assert a.calls == b.calls and len(a.calls) == 3
    assert [w for w, binary in h.written] == mine and all(binary for _, binary in h.written)
    kinds = [server_io.decode_response(m)["type"] for m in mine]
    assert kinds == [server_io.RequestType.RETURN_PREVIEW.value, server_io.RequestType.RETURN_STAMP.value]
    # settings reach the model as numpy scalars straight from np.frombuffer (server_io.py:105-119)
    assert a.calls[2][5]["steps"][0] == "uint8" and a.calls[2][5]["cfg_weight"][0] == "float32"
    # an exception inside the model is swallowed by on_message (handler.py:83-89): nothing is written
    h2 = handler.InpaintWebSocketHandler(model=None, model_info_str="trt", debug_dir=None)
    h2.on_message(frames[1])
    assert h2.written == []


def test_fast_handler_uses_the_u8_path_and_answers_errors(reference_server):
    """opt-in subclass of the reference handler: uint8 wire image -> stamp_u8, error frame on failure."""
    handler, _, server_io = reference_server
    from diffusiontexturepainting_b200 import fast_handler as fh
    Fast = fh.make_fast_handler(handler.InpaintWebSocketHandler, server_io)
    R = 32

    class U8Model(Recorder):
        def stamp_u8(self, canvas_u8_hwc, **settings):
            assert canvas_u8_hwc.dtype == np.uint8 and canvas_u8_hwc.shape == (R, R, 4)
            self.calls.append(("stamp_u8", {k: type(v).__name__ for k, v in sorted(settings.items())}))
            return torch.from_numpy(canvas_u8_hwc[None, :, :, :3].copy()) // 2

    m = U8Model(R)
    h = Fast(model=m, model_info_str="trt", debug_dir=None)
    brush_frame, stamp_frame = twin.synthetic_frames(R, steps=5)
    h.on_message(brush_frame)
    h.on_message(stamp_frame)
    assert [c[0] for c in m.calls] == ["set_brush", "generate", "stamp_u8"]
    resp = server_io.decode_response(h.written[1][0])
    assert resp["type"] == server_io.RequestType.RETURN_STAMP.value
    assert np.array_equal(resp["image"], twin.binary_to_image(stamp_frame, 14)[..., :3] // 2)
    # a failure is answered, not swallowed
    h2 = Fast(model=None, model_info_str="trt", debug_dir=None)
    h2.on_message(stamp_frame)
    assert len(h2.written) == 1 and h2.written[0][1] is True
    msg = fh.decode_error_frame(h2.written[0][0])
    assert msg is not None and "AttributeError" in msg
    assert fh.decode_error_frame(h.written[1][0]) is None
    h2.on_message(b"\x07" + stamp_frame[1:])  # unknown request type
    assert "Unknown binary request type" in fh.decode_error_frame(h2.written[1][0])
