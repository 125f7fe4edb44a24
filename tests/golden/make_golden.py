"""Generates tests/golden/*.json by EXECUTING the reference's own code in place (never copied): run in the build
container, where /root/reference exists:   python tests/golden/make_golden.py

  ddim.json      DDIMScheduler tables + step() known answers      (trt_inference/utilities.py:370-529)
  posenc.json    positional_encoding_2d / pos_emb view-scramble    (trt_inference/image_encoder.py:20-56)
  patches.json   get_image_patches checksums                       (trt_inference/image_encoder.py:34-40)
  codec.json     server_io request / response byte strings         (trt_inference/server_io.py)
"""
import ast
import hashlib
import json
import math
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DTP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, names):
    """exec only the named top-level classes / functions of a reference module (its imports are not installable here)."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch, "math": math}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def ddim():
    ns = extract(os.path.join(REF, "trt_inference/utilities.py"), {"DDIMScheduler"})
    out = {}
    for S in (1, 2, 4, 10, 20, 50):
        s = ns["DDIMScheduler"](device="cpu", num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012)
        s.set_timesteps(S)
        s.configure()
        # initialize_timesteps (stable_diffusion_pipeline.py:348-355) restated inline on the reference scheduler object
        offset = s.steps_offset
        init_timestep = min(int(S * 1.0) + offset, S)
        t_start = max(S - init_timestep + offset, 0)
        x = torch.linspace(-1, 1, 256).view(1, 4, 8, 8)
        eps = torch.cos(torch.arange(256.0)).view(1, 4, 8, 8)
        steps = {}
        for idx in sorted({0, min(1, S - 1), S // 2, S - 1}):
            y = s.step(eps, x, idx, s.timesteps[idx])
            steps[str(idx)] = {"sum": float(y.double().sum()), "abs_sum": float(y.double().abs().sum()),
                               "first3": [float(v) for v in y.flatten()[:3]]}
        out[str(S)] = {"timesteps": [int(t) for t in s.timesteps], "alphas": [float(a) for a in s.alphas_cumprod],
                       "final_alpha": float(s.final_alpha_cumprod), "t_start": t_start, "steps": steps}
    return out


def posenc():
    ns = extract(os.path.join(REF, "trt_inference/image_encoder.py"), {"positional_encoding_2d", "get_image_patches"})
    pe = ns["positional_encoding_2d"]
    out = {}
    for hid in (768, 128):
        parts = [pe(hid, int(math.sqrt(i)), int(math.sqrt(i))).view(1, i, hid) for i in (1, 4, 9)]
        t = torch.cat(parts, dim=1)
        out[str(hid)] = {"abs_sum": float(t.double().abs().sum()), "token0": [float(v) for v in t[0, 0, :8]],
                         "token1": [float(v) for v in t[0, 1, :8]], "token13_tail": [float(v) for v in t[0, 13, -8:]],
                         "sha": hashlib.sha256(t.numpy().astype("<f4").tobytes()).hexdigest()}
    img = torch.arange(3 * 224 * 224, dtype=torch.float32).view(1, 3, 224, 224)
    patches = {}
    for ps in (224, 112, 74):
        p = ns["get_image_patches"](img, ps)
        patches[str(ps)] = {"shape": list(p.shape), "sum": float(p.double().sum()),
                            "corner": [float(p[i, 0, 0, 0]) for i in range(p.shape[0])]}
    return out, patches


def codec():
    sys.path.insert(0, os.path.join(REF, "trt_inference"))
    import server_io  # numpy only
    rng = np.random.default_rng(0)
    canvas = rng.integers(0, 256, size=(8, 8, 4), dtype=np.uint8)
    req = (server_io.encode_request_type(server_io.RequestType.NEW_STAMP)
           + server_io.encode_inference_settings(steps=20, width=256, context_pad=150, cfg_weight=2.0, tg_weight=1.0,
                                                 tg_steps=20)
           + server_io.image_to_binary(canvas))
    resp_img = rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8)
    resp = server_io.encode_generated_response(server_io.RequestType.RETURN_STAMP, resp_img)
    out = {"response_hex": bytes(resp).hex(), "response_image_hex": resp_img.tobytes().hex()}
    meta, settings, off = server_io.decode_request_metadata(req)
    out["request_hex"] = bytes(req).hex()
    out["canvas_hex"] = canvas.tobytes().hex()
    out["request_type"] = int(meta["type"])
    out["request_settings"] = {k: float(v) for k, v in settings.items()}
    out["request_image_offset"] = int(off)
    return out


if __name__ == "__main__":
    json.dump(ddim(), open(os.path.join(HERE, "ddim.json"), "w"), indent=1)
    pe, patches = posenc()
    json.dump(pe, open(os.path.join(HERE, "posenc.json"), "w"), indent=1)
    json.dump(patches, open(os.path.join(HERE, "patches.json"), "w"), indent=1)
    try:
        json.dump(codec(), open(os.path.join(HERE, "codec.json"), "w"), indent=1)
    except Exception as e:  # pragma: no cover
        print("codec golden skipped:", e)
    print("golden vectors written to", HERE)
