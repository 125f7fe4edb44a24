"""CPU tests of the host-side serving helpers: brush-embedding cache and stamp micro-batcher.
The model is a stand-in that records its calls; the arithmetic is not under test here."""
import threading
import time

import pytest
import torch

from diffusiontexturepainting_b200.serving import BrushCache, StampBatcher

SETTINGS = dict(steps=10, context_pad=150, tg_steps=10, width=256, cfg_weight=2.0, tg_weight=1.0)


def test_brush_cache_key_depends_on_content_shape_and_resolution():
    a = torch.rand(3, 40, 50)
    b = a.clone()
    assert BrushCache.key(a, 256) == BrushCache.key(b, 256)
    b[0, 0, 0] += 1e-3
    assert BrushCache.key(a, 256) != BrushCache.key(b, 256)
    assert BrushCache.key(a, 256) != BrushCache.key(a, 512)
    assert BrushCache.key(a, 256) != BrushCache.key(a.reshape(3, 50, 40), 256)
    # non-contiguous views hash by value
    t = torch.rand(3, 20, 30)
    assert BrushCache.key(t.transpose(1, 2), 64) == BrushCache.key(t.transpose(1, 2).contiguous(), 64)


def test_brush_cache_is_lru_with_bounded_size():
    c = BrushCache(3)
    for i in range(3):
        c.put(i, f"v{i}")
    assert c.get(0) == "v0"          # 0 becomes most recent
    c.put(3, "v3")                    # evicts 1
    assert c.get(1) is None and c.get(0) == "v0" and c.get(2) == "v2" and c.get(3) == "v3"
    assert len(c) == 3 and c.hits == 4 and c.misses == 1
    with pytest.raises(ValueError):
        BrushCache(0)


def test_set_brush_uses_the_cache_without_touching_the_encoder():
    """TRTConditionalInpainter.set_brush: a brush seen before costs set_condition only (no resize, no encoder forward)."""
    from diffusiontexturepainting_b200 import trt_model

    calls = {"encode": 0, "cond": 0}

    class Enc:
        def encode_image(self, image):
            calls["encode"] += 1
            return torch.full((1, 14, 8), float(calls["encode"])), torch.zeros(1, 14, 8)

    class Pipe:
        device = torch.device("cpu")

        def set_condition(self, a, b):
            calls["cond"] += 1
            self.last = a

    m = trt_model.TRTConditionalInpainter.__new__(trt_model.TRTConditionalInpainter)
    m.image_encoder, m.pipeline, m._resolution = Enc(), Pipe(), 32
    m.brush_cache = BrushCache(2)
    m.conditioning = m.image = None
    m.lock, m.brush_generation = threading.RLock(), 0
    b1, b2 = torch.rand(3, 48, 64), torch.rand(3, 64, 48)
    m.set_brush(b1)
    e1 = m.conditioning[0].clone()
    m.set_brush(b2)
    m.set_brush(b1)
    assert calls == {"encode": 2, "cond": 3}
    assert torch.equal(m.conditioning[0], e1) and torch.equal(m.pipeline.last, e1)
    assert m.image.shape == (1, 3, 32, 32) and m.brush_generation == 3
    m.brush_cache = None              # cache disabled: every switch re-encodes
    m.set_brush(b1)
    assert calls["encode"] == 3


class FakeModel:
    def __init__(self, delay=0.0, fail_on=None):
        self.calls, self.delay, self.fail_on = [], delay, fail_on
        self.lock = threading.Lock()

    def generate(self, canvas, **settings):
        with self.lock:
            self.calls.append((canvas.shape[0], settings["steps"], canvas.shape[-1]))
        if self.delay:
            time.sleep(self.delay)
        if self.fail_on is not None and settings["steps"] == self.fail_on:
            raise RuntimeError("engine failure")
        # stamp i = mean of canvas i broadcast to (3,R,R): lets the test check the rows were routed back correctly
        m = canvas.mean(dim=(1, 2, 3), keepdim=True)
        return m.expand(-1, 3, canvas.shape[-1], canvas.shape[-1]).clone()


def test_batcher_coalesces_concurrent_requests_and_routes_results():
    model = FakeModel(delay=0.02)
    with StampBatcher(model.generate, max_batch=4, max_wait_ms=50) as b:
        canv = [torch.full((4, 16, 16), float(i)) for i in range(6)]
        futs = [b.submit(c, SETTINGS) for c in canv]
        outs = [f.result(timeout=5) for f in futs]
    for i, o in enumerate(outs):
        assert o.shape == (3, 16, 16) and torch.all(o == float(i))
    assert sum(n for n, _, _ in model.calls) == 6
    assert max(n for n, _, _ in model.calls) == 4          # first batch fills up, the rest follows
    assert len(model.calls) <= 3


def test_batcher_never_mixes_settings_or_resolutions():
    model = FakeModel()
    other = dict(SETTINGS, steps=20)
    with StampBatcher(model.generate, max_batch=8, max_wait_ms=20) as b:
        futs = [b.submit(torch.zeros(4, 16, 16), SETTINGS), b.submit(torch.zeros(4, 16, 16), other),
                b.submit(torch.zeros(1, 4, 32, 32), SETTINGS), b.submit(torch.zeros(4, 16, 16), SETTINGS)]
        for f in futs:
            f.result(timeout=5)
    assert sorted(model.calls) == sorted([(2, 10, 16), (1, 20, 16), (1, 10, 32)])


def test_batcher_delivers_model_errors_to_every_caller_of_the_batch_and_keeps_running():
    model = FakeModel(fail_on=20)
    with StampBatcher(model.generate, max_batch=4, max_wait_ms=10) as b:
        bad = [b.submit(torch.zeros(4, 8, 8), dict(SETTINGS, steps=20)) for _ in range(2)]
        for f in bad:
            with pytest.raises(RuntimeError, match="engine failure"):
                f.result(timeout=5)
        ok = b.submit(torch.ones(4, 8, 8), SETTINGS)
        assert torch.all(ok.result(timeout=5) == 1.0)
    with pytest.raises(RuntimeError):
        b.submit(torch.zeros(4, 8, 8), SETTINGS)


def test_batcher_rejects_malformed_canvases():
    with StampBatcher(FakeModel().generate) as b:
        for bad in (torch.zeros(3, 8, 8), torch.zeros(2, 4, 8, 8), torch.zeros(4, 8, 9)):
            with pytest.raises(ValueError):
                b.submit(bad, SETTINGS)


def test_batcher_keeps_brushes_apart_and_fails_a_batch_whose_brush_was_replaced():
    """stamps submitted under different brushes must not be coalesced, and a batch must not be painted with a
    brush that replaced the one it was submitted under."""
    model = FakeModel(delay=0.05)
    gen = {"v": 0}
    lock = threading.RLock()
    with StampBatcher(model.generate, max_batch=8, max_wait_ms=30, brush_key=lambda: gen["v"], lock=lock) as b:
        f0 = b.submit(torch.zeros(4, 8, 8), SETTINGS)
        f0.result(timeout=5)
        f1 = b.submit(torch.zeros(4, 8, 8), SETTINGS)   # submitted under brush 0 ...
        with lock:
            gen["v"] = 1                                  # ... which is replaced before the batch runs
        f2 = b.submit(torch.ones(4, 8, 8), SETTINGS)     # submitted under brush 1
        with pytest.raises(RuntimeError, match="brush changed"):
            f1.result(timeout=5)
        assert torch.all(f2.result(timeout=5) == 1.0)
    assert [n for n, _, _ in model.calls] == [1, 1]      # f1 never reached the model, f2 ran alone
