"""Flat-module shim: put this directory AHEAD of the reference's trt_inference/ on PYTHONPATH and run.py's
`from trt_model import ...` resolves to the sm_100a implementation, while handler.py, server_io.py, model_base.py,
websocket_model.py and run.py keep resolving to the reference files."""
from diffusiontexturepainting_b200.trt_model import *  # noqa: F401,F403
from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter, crop_resize_square  # noqa: F401,E402
