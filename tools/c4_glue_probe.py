"""Which Python lines launch the torch-native (non-library) CUDA kernels of the C4 self-supervised step?  Runs the step of
bench.py --workload c4 at a reduced batch under torch.profiler with stacks and prints, per aten kernel, total time, count and
the innermost repo frame.  Development aid; prints text."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    args = types.SimpleNamespace(gpus=1, batch=int(os.environ.get("BATCH", "16")), steps=2, warmup=3, per_term_losses=False)
    captured = {}
    orig_timer = bench._timer

    def fake_timer(world):
        def timed(fn, steps, warmup):
            if "fn" not in captured:
                captured["fn"] = fn
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            return 1.0
        return timed
    bench._timer = fake_timer
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        bench.run_ssl(args)
    bench._timer = orig_timer
    fn = captured["fn"]
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        fn()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages(group_by_stack_n=12, group_by_input_shape=True):
        t = getattr(ev, "self_device_time_total", 0) or getattr(ev, "self_cuda_time_total", 0)
        if t <= 0 or not ev.key.startswith("aten::"):
            continue
        frame = next((s for s in (ev.stack or []) if "/apla_b200/" in s or "bench.py" in s), "(torch internals / autograd)")
        rows.append((t, ev.count, ev.key, str(ev.input_shapes)[:70], frame.strip()[:120]))
    tot = sum(r[0] for r in rows)
    print(f"aten ops with device time of their own: {tot / 1e3:.2f} ms in one step (batch {args.batch})")
    for t, c, name, shp, frame in sorted(rows, key=lambda r: -r[0])[:30]:
        print(f"{t / 1e3:8.3f} ms  x{c:<4d} {name:26s} {shp:70s} {frame}")


if __name__ == "__main__":
    main()
