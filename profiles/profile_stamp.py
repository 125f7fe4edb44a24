"""One warm + one profiled stamp of the benchmark workload; writes the per-op device-time table (CUDA events around every
launch on the launching stream) to gpurun_out/<tag>_ops.csv. Under `ncu --metrics gpu__time_duration.sum` the same
command yields the launch list committed next to it.
    python profiles/profile_stamp.py [--resolution 512] [--denoise-steps 20] [--tag r1] [--stamps 1] [--no-op-profile]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DTP_SYNTHETIC_WEIGHTS", "1")  # profiling runs on the seeded synthetic inventory
import torch  # noqa: E402

from diffusiontexturepainting_b200 import weights as W  # noqa: E402
from diffusiontexturepainting_b200.testdata import make_canvas, smooth_image  # noqa: E402
from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--resolution", type=int, default=512)
ap.add_argument("--denoise-steps", type=int, default=20)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--tag", default="r1")
ap.add_argument("--stamps", type=int, default=1)
ap.add_argument("--no-op-profile", action="store_true")
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--profiler-range", action="store_true",
                help="cudaProfilerStart/Stop around the measured stamps (ncu --profile-from-start off)")
ap.add_argument("--ablate", action="store_true",
                help="graph-mode stamp time with each kernel family skipped in turn (in-situ cost = difference)")
ap.add_argument("--ablate-labels", type=int, default=0, metavar="N",
                help="in-graph cost of the N most expensive op labels (stamp time with that label's launches left out)")
a = ap.parse_args()
R, S, B = a.resolution, a.denoise_steps, a.batch
model = TRTConditionalInpainter(R, device=0, model_config=W.sd15_config(), max_batch_size=B)
model.pipeline.strict_schedule = True
model.set_brush(smooth_image(1, 3, R))
canvas = make_canvas(B, R).cuda()
lat = torch.randn(B, 4, R // 8, R // 8, generator=torch.Generator().manual_seed(42)).cuda()
model.pipeline.update_infer_settings(S, 2.0, 1.0, S)
model.pipeline._push_schedule(1.0)
out = torch.empty(B, 3, R, R, device="cuda")
eng = model.engine
if a.no_graph:
    eng.set_option("graph", 0)
if a.ablate:
    KINDS = ["other", "contraction", "groupnorm", "layernorm", "softmax", "attn_small", "flash_attn"]

    def timed(n=4):
        for _ in range(2):
            eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    base = timed()
    print("graph-mode stamp: %.2f ms" % base)
    for k, name in enumerate(KINDS):
        if name in ("other", "softmax"):
            continue
        eng.set_option("debug_skip_kinds", 1 << k)
        t = timed()
        print("  without %-12s %.2f ms  -> in-situ cost %.2f ms" % (name, t, base - t), flush=True)
    eng.set_option("debug_skip_kinds", 0)
    sys.exit(0)
if a.ablate_labels:
    import csv

    def fnv(text):
        h = 2166136261
        for ch in text.encode():
            h = ((h ^ ch) * 16777619) & 0xffffffff
        return (h & 0x7fffffff) or 1

    def timed(n=3):
        for _ in range(2):
            eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    eng.set_option("graph", 0)
    eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
    eng.set_option("profile", 1)
    eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
    torch.cuda.synchronize()
    path = os.path.join(ROOT, "gpurun_out", f"{a.tag}_ops.csv")
    eng.profile_dump(path)
    eng.set_option("profile", 0)
    eng.set_option("graph", 1)
    rows = list(csv.DictReader(open(path)))[:a.ablate_labels]
    base = timed(5)
    lines = ["label,calls,eager_avg_us,in_graph_avg_us,in_graph_total_ms", "# graph-mode stamp %.3f ms" % base]
    total = 0.0
    for r in rows:
        eng.set_option("debug_skip_label", fnv(r["op"]))
        t = timed()
        calls = int(r["calls"])
        cost = base - t
        total += cost
        lines.append("%s,%d,%.2f,%.2f,%.3f" % (r["op"], calls, float(r["avg_us"]), cost * 1e3 / calls, cost))
        print(lines[-1], flush=True)
    eng.set_option("debug_skip_label", 0)
    lines.append("# sum of the listed in-graph costs %.2f ms of %.2f ms" % (total, base))
    print(lines[-1])
    open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_label_costs.csv"), "w").write("\n".join(lines) + "\n")
    sys.exit(0)
eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
torch.cuda.synchronize()
if not a.no_op_profile:
    eng.set_option("profile", 1)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if a.profiler_range:
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
t0.record()
for _ in range(a.stamps):
    eng.stamp(canvas, model.image, 150, lat, None, composite=True, out_f32=out)
t1.record()
torch.cuda.synchronize()
if a.profiler_range:
    torch.cuda.profiler.stop()
print("ms per stamp (with per-op events)" if not a.no_op_profile else "ms per stamp", t0.elapsed_time(t1) / a.stamps)
if not a.no_op_profile:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"{a.tag}_ops.csv")
    eng.profile_dump(path)
    print(eng.profile())
    print(open(path).read()[:6000])
