"""A/B a library switch that is read from the environment (e.g. AN3D_FWD_RING=1) on the GPU box.

    python tools/ab_env.py AN3D_FWD_RING=1 [--workload c3]      # 1 = accumulator ring, 2 = early slot release, 3 = both [--tests tests/test_gpu_conv_stack.py ...]

Runs, in separate processes (the library reads its switches once): the given GPU tests with the switch ON (a variant
that is not parity-green is not worth timing), then `bench.py` with the switch OFF and ON, and prints ms/step, the
metric and the per-kernel device times side by side.  Results also go to gpurun_out/ab_<NAME>.json."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run_bench(env, workload, steps, warmup):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", workload, "--steps", str(steps), "--warmup",
                        str(warmup), "--no-cpu-baseline", "--no-configs"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError(f"bench failed (rc={r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return json.loads(lines[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("switch", help="NAME=VALUE")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--tests", nargs="*", default=["tests/test_gpu_conv_stack.py", "tests/test_gpu_bf16.py"])
    a = ap.parse_args()
    name, value = a.switch.split("=", 1)
    off = {k: v for k, v in os.environ.items() if k != name}
    on = dict(off, **{name: value})
    if a.tests:
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", *a.tests], env=on, cwd=ROOT, timeout=900)
        if r.returncode != 0:
            sys.exit(f"{a.switch}: parity tests fail with the switch on -- not timing it")
    res = {"off": run_bench(off, a.workload, a.steps, a.warmup), "on": run_bench(on, a.workload, a.steps, a.warmup)}
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"ab_{name}.json").write_text(json.dumps(res))
    print(f"{'':28s}{'off':>12s}{'on':>12s}")
    print(f"{'ms_per_step':28s}{res['off']['ms_per_step']:12.3f}{res['on']['ms_per_step']:12.3f}")
    print(f"{res['off']['unit']:28s}{res['off']['value']:12.0f}{res['on']['value']:12.0f}")
    names = {"0": "conv layer-2 statistics", "1": "conv full pass", "2": "t1_sparse", "3": "dgrad3", "4": "bwd_l2", "5": "fc gemm"}
    tags_off, tags_on = res["off"].get("kernel_ms_by_tag", {}), res["on"].get("kernel_ms_by_tag", {})
    for k in sorted(set(tags_off) | set(tags_on)):
        print(f"{names.get(k, k) + ' ms':28s}{tags_off.get(k, float('nan')):12.3f}{tags_on.get(k, float('nan')):12.3f}")


if __name__ == "__main__":
    main()
