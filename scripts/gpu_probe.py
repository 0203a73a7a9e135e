"""Quick GPU probe: device info, TF32/BF16 cuBLAS peaks (roofline denominators), and a small
knn_rms throughput sample with phase timings.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_peak():
    import torch
    out = {}
    for name, dt, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False), ("fp32", torch.float32, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        n = 8192
        a = torch.randn(n, n, device="cuda", dtype=dt)
        b = torch.randn(n, n, device="cuda", dtype=dt)
        for _ in range(3):
            (a @ b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); (a @ b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        burst = 2 * n ** 3 / best / 1e9
        t0 = time.time(); it = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 3.0:
            for _ in range(10):
                (a @ b)
            it += 10
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        out[name] = {"burst_tflops": burst, "sustained_tflops": 2 * n ** 3 * it / e0.elapsed_time(e1) / 1e9}
    return out


def main():
    import torch
    import mdsctk_b200
    from mdsctk_b200 import synth
    res = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count(), "sms": torch.cuda.get_device_properties(0).multi_processor_count}
    n = int(os.environ.get("PROBE_N", "20000"))
    xyz = synth.traj_frames(n, 300, int(os.environ.get("PROBE_BASINS", "16")))
    mass = synth.traj_masses(300)
    ctx = mdsctk_b200.KnnContext(0)
    for kern in [int(x) for x in os.environ.get("PROBE_KERNELS", "0").split(",")]:
        ctx.set_option("rms_kernel", kern)
        if os.environ.get("PROBE_SLACK"):
            ctx.set_option("slack", int(os.environ["PROBE_SLACK"]))
        ctx.rms_set_reference(xyz, mass)
        for rep in range(3):
            t = time.time()
            ctx.rms_query(33, fetch=False)
            dt = time.time() - t
            st = ctx.stats()
        res[f"rms_kernel{kern}"] = {"n": n, "wall_s": dt, "pairs_per_s": n * n / (st["ms_sweep"] + st["ms_rescore"] + st["ms_fallback"]) * 1e3, **st}
        print(json.dumps(res[f"rms_kernel{kern}"]), flush=True)
    if os.environ.get("PROBE_PEAKS", "1") == "1":
        res["peaks"] = tf32_peak()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
