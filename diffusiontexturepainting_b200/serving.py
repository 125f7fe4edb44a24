"""Host-side helpers around the stamp path for the websocket server. Pure Python: nothing
here touches the arithmetic, the CUDA library does all of it.

* BrushCache — the Kit client keeps a history of at most ten brushes and switches between them interactively
  (kit_app/.../extension.py:197-204); every switch re-runs crop/resize + the CLIP patch encoder in the reference
  (trt_model.py:79-88). The cache keys the brush by a digest of the image the client sent and returns the resized brush and
  its conditioning, so a switch back costs one K/V projection (dtp_set_condition) instead of an encoder forward.
* StampBatcher — the reference handles one stamp at a time inside the tornado IOLoop (handler.py:78-123, run.py:54-55), one
  model instance shared by all connections. Stamps of different connections are independent given the brush, so concurrent
  requests with identical settings can be coalesced into one `generate(B > 1)` call, which is where the GPU path is
  efficient (15 ms per stamp at B = 4 vs 72 ms at B = 1, 256 x 256 / 10 evaluations).
"""
from __future__ import annotations

import collections
import hashlib
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, Dict, Hashable, List, Optional, Tuple

import torch


class BrushCache:
    """LRU of (resized brush on the device, conditioning) keyed by a digest of the client's image and the resolution."""

    def __init__(self, capacity: int = 10):
        if capacity < 1:
            raise ValueError("capacity must be >= 1")
        self.capacity = int(capacity)
        self._items: "collections.OrderedDict[Hashable, Any]" = collections.OrderedDict()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def key(image: torch.Tensor, resolution: int) -> Tuple[int, Tuple[int, ...], str]:
        t = image.detach()
        if t.device.type != "cpu":
            t = t.cpu()
        t = t.contiguous()
        digest = hashlib.blake2b(t.numpy().tobytes(), digest_size=16).hexdigest()
        return int(resolution), tuple(t.shape), f"{t.dtype}:{digest}"

    def get(self, key):
        item = self._items.get(key)
        if item is None:
            self.misses += 1
            return None
        self._items.move_to_end(key)
        self.hits += 1
        return item

    def put(self, key, value) -> None:
        self._items[key] = value
        self._items.move_to_end(key)
        while len(self._items) > self.capacity:
            self._items.popitem(last=False)

    def clear(self) -> None:
        self._items.clear()

    def __len__(self) -> int:
        return len(self._items)


def _settings_key(settings: Dict[str, Any]) -> Tuple:
    """Requests can share a batch only if every scheduler / guidance setting agrees (the engine holds one schedule)."""
    return (int(settings["steps"]), int(settings["context_pad"]), int(settings["tg_steps"]), float(settings["cfg_weight"]),
            float(settings["tg_weight"]))


class StampBatcher:
    """Coalesces concurrent stamp requests into batched `generate` calls.

    submit(canvas (4,R,R) or (1,4,R,R), settings) -> Future resolving to the (3,R,R) stamp. A worker thread takes the oldest
    pending request, waits up to `max_wait_ms` for more requests with the same (resolution, settings), runs
    `generate(cat(canvases), **settings)` once (at most `max_batch` stamps) and distributes the rows. An exception of the
    model call is delivered to every future of that batch (the reference swallows it in the handler, handler.py:83-89).
    Consecutive stamps of ONE stroke depend on each other through the texture: a connection must wait
    for its previous stamp before submitting the next one, which the request/response protocol already enforces.
    """

    def __init__(self, generate: Callable[..., torch.Tensor], max_batch: int = 8, max_wait_ms: float = 2.0,
                 brush_key: Optional[Callable[[], Hashable]] = None, lock=None):
        """`lock`: the model's lock (TRTConditionalInpainter.lock); held across the brush check and the batched call.
        `generate`: the model's bound `generate` (TRTConditionalInpainter.generate takes the model lock, so a
        `set_brush` on the IOLoop thread cannot interleave with a batched call). `brush_key`: returns an id of the brush that
        is current when a stamp is SUBMITTED (e.g. `lambda: model.brush_generation`): stamps submitted under different
        brushes never share a batch, and a batch whose brush has been replaced before it runs fails instead of being
        painted with the wrong texture."""
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self._generate = generate
        self._brush_key = brush_key
        self._lock = lock if lock is not None else threading.RLock()
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._pending: List[Tuple[Tuple, torch.Tensor, Dict[str, Any], Future]] = []
        self._cv = threading.Condition()
        self._closed = False
        self.batches: List[int] = []  # sizes of the batches executed so far (observability / tests)
        self._worker = threading.Thread(target=self._run, name="dtp-stamp-batcher", daemon=True)
        self._worker.start()

    def submit(self, canvas: torch.Tensor, settings: Dict[str, Any]) -> Future:
        c = canvas if canvas.dim() == 4 else canvas.unsqueeze(0)
        if c.dim() != 4 or c.shape[0] != 1 or c.shape[1] != 4 or c.shape[2] != c.shape[3]:
            raise ValueError(f"canvas must be (4,R,R) or (1,4,R,R), got {tuple(canvas.shape)}")
        fut: Future = Future()
        key = (int(c.shape[-1]), self._brush_key() if self._brush_key else None) + _settings_key(settings)
        with self._cv:
            if self._closed:
                raise RuntimeError("StampBatcher is closed")
            self._pending.append((key, c, dict(settings), fut))
            self._cv.notify_all()
        return fut

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _take(self) -> Optional[List[Tuple[Tuple, torch.Tensor, Dict[str, Any], Future]]]:
        with self._cv:
            while not self._pending and not self._closed:
                self._cv.wait()
            if not self._pending:
                return None
            key = self._pending[0][0]
            deadline = time.monotonic() + self.max_wait
            while True:
                same = [r for r in self._pending if r[0] == key]
                left = deadline - time.monotonic()
                if len(same) >= self.max_batch or left <= 0 or self._closed:
                    break
                self._cv.wait(timeout=left)
            batch = same[:self.max_batch]
            taken = set(id(r) for r in batch)
            self._pending = [r for r in self._pending if id(r) not in taken]
            return batch

    def _run(self) -> None:
        while True:
            batch = self._take()
            if batch is None:
                return
            live = [r for r in batch if r[3].set_running_or_notify_cancel()]
            if not live:
                continue
            try:
                with self._lock:
                    if self._brush_key is not None and self._brush_key() != live[0][0][1]:
                        raise RuntimeError("the brush changed between submission and execution of this stamp batch")
                    canvases = torch.cat([r[1] for r in live], dim=0)
                    out = self._generate(canvases, **live[0][2])
                if out.shape[0] != len(live):
                    raise RuntimeError(f"model returned {out.shape[0]} stamps for a batch of {len(live)}")
                self.batches.append(len(live))
                for i, r in enumerate(live):
                    r[3].set_result(out[i])
            except BaseException as e:  # noqa: BLE001 - delivered to the callers
                for r in live:
                    if not r[3].done():
                        r[3].set_exception(e)
