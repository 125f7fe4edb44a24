"""Optional websocket handler for the reference server. The reference's handler is a KEPT file and
works unchanged with this package; this subclass is an opt-in that a maintainer selects in run.py with

    from diffusiontexturepainting_b200.fast_handler import make_fast_handler
    import handler, server_io
    InpaintWebSocketHandler = make_fast_handler(handler.InpaintWebSocketHandler, server_io)

It changes two things and nothing else (same wire format, same request types):
  * a stamp request takes the uint8 HWC wire image straight to the GPU (`TRTConditionalInpainter.stamp_u8`: one dtp_stamp
    call for de-quantisation, pre-process, inference, composite and the x255 truncation) instead of the fp32 round trip of
    handler.py:104-110 (np_to_torch -> .to(device) -> generate -> .cpu() -> torch_to_np);
  * a failing request is answered with an ERROR frame instead of being swallowed (handler.py:83-89 logs and returns, and the
    client then blocks in recv forever, trt_inference/websocket_model.py:81): one byte RESPONSE_ERROR (255), an int32
    length and the UTF-8 message. The reference's request / response types (server_io.py:19-24) end at 4.
The module imports neither tornado nor the reference: the base class and the codec module are passed in."""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)
RESPONSE_ERROR = 255


def encode_error_frame(message: str) -> bytes:
    payload = message.encode("utf-8", errors="replace")[:4096]
    return np.array([RESPONSE_ERROR], dtype=np.uint8).tobytes() + np.array([len(payload)], dtype=np.int32).tobytes() + payload


def decode_error_frame(frame: bytes):
    """-> message if `frame` is an error frame, else None (client side)."""
    if len(frame) < 5 or frame[0] != RESPONSE_ERROR:
        return None
    n = int(np.frombuffer(frame, dtype=np.int32, count=1, offset=1)[0])
    return bytes(frame[5:5 + n]).decode("utf-8", errors="replace")


def make_fast_handler(base_handler, server_io):
    class FastInpaintWebSocketHandler(base_handler):
        def on_message(self, message):
            try:
                if type(message) == bytes:
                    self._handle_binary_request(message)
                else:
                    self._handle_json_request(message)
            except Exception as e:  # noqa: BLE001 - reported to the client instead of leaving it blocked
                logger.error("Failed to handle incoming message: %s", e)
                self.write_message(encode_error_frame(f"{type(e).__name__}: {e}"), binary=True)

        def _handle_stamp_request(self, inference_settings, context):
            stamp_u8 = getattr(self.model, "stamp_u8", None)
            if stamp_u8 is None:
                return base_handler._handle_stamp_request(self, inference_settings, context)
            result = stamp_u8(np.ascontiguousarray(context), **inference_settings)  # (1, H, W, 3) uint8 on the device
            img = result[0].cpu().numpy()
            self.write_message(server_io.encode_generated_response(server_io.RequestType.RETURN_STAMP, img), binary=True)

    FastInpaintWebSocketHandler.__name__ = "FastInpaintWebSocketHandler"
    return FastInpaintWebSocketHandler
