"""Test helper: the call sequence of the reference's websocket handler for one binary frame, restated so the GPU box
(which has no copy of the reference tree) can drive the model exactly as the server does. It follows
trt_inference/handler.py:48-60 (preview_mask, torch_to_np, np_to_torch), :92-123 (_handle_new_image_brush_request,
_handle_stamp_request, _handle_binary_request) and the wire layout of trt_inference/server_io.py:65-165.

tests/test_dropin_handler.py proves, in the build container where the reference exists, that this twin and the UNMODIFIED
handler.py (imported through the dropin/ shims) make identical calls on the model and write identical bytes."""
import numpy as np
import torch

NEW_BRUSH_IMAGE, NEW_STAMP, RETURN_PREVIEW, RETURN_STAMP = 0, 2, 3, 4


def decode_request_metadata(msg, offset=0):
    """server_io.py:88-122: u8 type, u8 steps, u8 context_pad, u8 tg_steps, u16 width, f32 cfg_weight, f32 tg_weight."""
    b = np.frombuffer(msg, dtype=np.uint8, count=4, offset=offset)
    w = np.frombuffer(msg, dtype=np.uint16, count=1, offset=offset + 4)
    f = np.frombuffer(msg, dtype=np.float32, count=2, offset=offset + 6)
    settings = {"steps": b[1], "context_pad": b[2], "tg_steps": b[3], "width": w[0], "cfg_weight": f[0],
                "tg_weight": f[1]}
    return {"type": b[0]}, settings, offset + 14


def binary_to_image(msg, offset=0):
    """server_io.py:65-85: i32 width, height, channels, then H*W*C bytes."""
    w, h, c = (int(v) for v in np.frombuffer(msg, dtype=np.int32, count=3, offset=offset))
    data = np.frombuffer(msg, dtype=np.uint8, offset=offset + 12)
    return data[:h * w * c].reshape(h, w, c)


def image_to_binary(img):
    return np.array([img.shape[1], img.shape[0], img.shape[2]], dtype=np.int32).tobytes() + img.tobytes()


def encode_request(kind, image, steps=20, width=256, context_pad=150, cfg_weight=2.0, tg_weight=1.0, tg_steps=20):
    """client side of server_io.py:125-152"""
    return (np.array([kind], dtype=np.uint8).tobytes() + np.array([steps, context_pad, tg_steps], dtype=np.uint8).tobytes()
            + np.array([width], dtype=np.uint16).tobytes() + np.array([cfg_weight], dtype=np.float32).tobytes()
            + np.array([tg_weight], dtype=np.float32).tobytes() + image_to_binary(image))


def torch_to_np(img):  # handler.py:55-56
    return (img.detach() * 255).to(torch.uint8).permute(1, 2, 0).numpy()


def np_to_torch(img):  # handler.py:59-60
    return torch.from_numpy(img).to(torch.float32).permute(2, 0, 1) / 255


def handle_binary_request(model, raw_message):
    """-> the bytes the handler passes to write_message(binary=True) (handler.py:113-123)."""
    meta, settings, off = decode_request_metadata(raw_message)
    if meta["type"] == NEW_BRUSH_IMAGE:
        image = binary_to_image(raw_message, off)[..., :3]
        model.set_brush(np_to_torch(image))                                   # handler.py:94
        res = model.resolution()
        mask = torch.zeros(1, 1, res, res)                                    # preview_mask, handler.py:48-52
        mask[..., :res // 2, :res // 2] = 1
        mask = mask.to(model.device())
        context = torch.cat([model.image, mask], dim=1)                       # handler.py:97
        result = model.generate(context, **settings).cpu()                    # handler.py:98
        return np.array([RETURN_PREVIEW], dtype=np.uint8).tobytes() + image_to_binary(torch_to_np(result[0, ...]))
    if meta["type"] == NEW_STAMP:
        context = binary_to_image(raw_message, off)
        context = np_to_torch(context).unsqueeze(0).to(model.device())        # handler.py:106
        result = model.generate(context, **settings).cpu()                    # handler.py:107
        return np.array([RETURN_STAMP], dtype=np.uint8).tobytes() + image_to_binary(torch_to_np(result[0, ...]))
    raise NotImplementedError(f"Unknown binary request type {meta['type']}")


def synthetic_frames(R, seed=0, **settings):
    """(new-brush frame, stamp frame) with seeded smooth content."""
    g = torch.Generator().manual_seed(seed)

    def smooth(c):
        x = torch.rand(1, c, R // 8 + 1, R // 8 + 1, generator=g)
        return torch.nn.functional.interpolate(x, size=(R, R), mode="bilinear", align_corners=True)[0].clamp(0, 1)
    brush = (smooth(3).permute(1, 2, 0) * 255).to(torch.uint8).numpy()
    canvas = torch.cat([smooth(3), torch.zeros(1, R, R)])
    canvas[3, :int(0.4 * R)] = 1.0
    canvas = (canvas.permute(1, 2, 0) * 255).to(torch.uint8).numpy()
    settings.setdefault("width", R)
    return (encode_request(NEW_BRUSH_IMAGE, np.ascontiguousarray(brush), **settings),
            encode_request(NEW_STAMP, np.ascontiguousarray(canvas), **settings))
