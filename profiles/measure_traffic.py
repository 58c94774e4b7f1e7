#!/usr/bin/env python
"""Regenerate profiles/traffic.json: DRAM bytes per env step (whole batch) of the bench workloads, measured with ncu.

    python profiles/measure_traffic.py            # on a B200 box, from the repository root

`bench.py` reports `roofline.traffic` from this file ONLY while its `source_hash` equals the hash of the kernel sources
of the running build (bench.source_hash()); after any kernel change the figure is dropped (traffic = null, with the
reason) until this script has been run again -- the number can no longer go stale silently.

How: the script re-invokes itself under
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
with `--inner <workload>`; the inner program builds the bench batch (bench.make_batch: same pre-roll, same actions),
runs warm-up steps, then brackets exactly K env steps with cudaProfilerStart/Stop.  Sum of the two counters over every
kernel launched inside the bracket / K = bytes per env step.  (ncu serialises kernels and flushes caches between
replays, so this is the cold-cache figure, an upper bound of the in-flight traffic.)
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_STEPS = 2
# (name, bench workload, extra environment, ncu --cache-control): "all" = ncu flushes the caches before every kernel (the cold-cache
# upper bound, comparable with earlier rounds); "none" = caches keep their contents between the (serialised) kernels, which is
# what the L2-resident work records of trex_config.chunk_envs need to show up
WORKLOADS = (
    ("random", "random", {}, "all"),
    ("standing", "standing", {}, "all"),
    ("c2", "c2", {}, "all"),
    ("random_warm", "random", {}, "none"),
    ("random_chunk8192_warm", "random", {"TREX_CHUNK": "8192", "TREX_PIPES": "2"}, "none"),
    ("random_chunk4096_warm", "random", {"TREX_CHUNK": "4096", "TREX_PIPES": "2"}, "none"),
)


def inner(workload):
    import torch

    import bench

    n = bench.WORKLOADS[workload][0]
    args = argparse.Namespace(no_contacts=False, substeps=5, warps_per_block=0, horizon=0, preroll=300 if workload != "standing" else 40)
    sim, acts = bench.make_batch(workload, n, 0, 0, args)
    for t in range(3):
        sim.step(acts[(args.preroll + t) % len(acts)])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for t in range(K_STEPS):
        sim.step(acts[(args.preroll + 3 + t) % len(acts)])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    st = sim.stats()
    print("INNER", json.dumps({"envs": n, "num_substeps": sim.num_substeps, "mean_contacts": st["mean_contacts"]}))


def outer():
    import bench

    out = {"source_hash": bench.source_hash(), "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every kernel of %d env steps / %d (profiles/measure_traffic.py)" % (K_STEPS, K_STEPS),
           "workloads": {}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for w, bench_w, extra_env, cache in WORKLOADS:
        log = os.path.join(ROOT, "gpurun_out", "traffic_%s.csv" % w)
        cmd = ["ncu", "--profile-from-start", "off", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
               "--cache-control", cache, "--csv", "--log-file", log, sys.executable, os.path.abspath(__file__), "--inner", bench_w]
        p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, **extra_env))
        meta = None
        for line in p.stdout.splitlines():
            if line.startswith("INNER "):
                meta = json.loads(line[6:])
        if meta is None:
            raise SystemExit("inner run failed for %s:\n%s\n%s" % (w, p.stdout[-2000:], p.stderr[-2000:]))
        rows = list(csv.reader(open(log)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        h = rows[hdr]
        ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total, per_kernel, launches = 0.0, {}, 0
        for r in rows[hdr + 1:]:
            if len(r) <= vi or "dram__bytes" not in r[mi]:
                continue
            b = float(r[vi].replace(",", "")) * scale[r[ui]]
            total += b
            name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            per_kernel[name] = per_kernel.get(name, 0.0) + b / K_STEPS
            launches += r[mi].endswith("read.sum")
        out["workloads"][w] = {"envs": meta["envs"], "num_substeps": meta["num_substeps"], "mean_contacts_per_env": meta["mean_contacts"],
                               "bytes_per_env_step_batch": int(total / K_STEPS), "bytes_per_env_step": total / K_STEPS / meta["envs"],
                               "kernel_launches_per_step": launches / K_STEPS,
                               "per_kernel_bytes_per_step": {k: int(v) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
                               "ncu_cache_control": cache, "environment": extra_env, "capture": "gpurun_out/traffic_%s.csv" % w}
        print(w, "%.3f GB per env step of %d envs" % (total / K_STEPS / 1e9, meta["envs"]), flush=True)
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote profiles/traffic.json for kernel sources", out["source_hash"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--inner", default=None)
    a = ap.parse_args()
    inner(a.inner) if a.inner else outer()
