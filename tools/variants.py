#!/usr/bin/env python
"""Build / time experimental builds of the library (same ABI, extra -D switches).
  python tools/variants.py build name1=-DX=1,-DY ... # (CPU box) -> tools/variants/lib_<name>.so
  python tools/variants.py run OUT.txt [bench args]   # (GPU box) bench.py --no-cpu per variant
Measurement tooling only."""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "variants")


def build(specs):
    from unconfined_b200.build import NVCC_FLAGS, CSRC, _nvcc
    os.makedirs(VDIR, exist_ok=True)

    def one(spec):
        name, _, defs = spec.partition("=")
        so = os.path.join(VDIR, f"lib_{name}.so")
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + [d for d in defs.split(",") if d] + ["-Xptxas", "-v", "-o", so,
                           os.path.join(CSRC, "capi.cu")], capture_output=True, text=True)
        info = [l for l in r.stderr.splitlines() if "lh_grid8" in l or "error" in l]
        return name, r.returncode, r.stderr[-1500:] if r.returncode else "\n".join(info[:2])
    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, rc, msg in ex.map(one, specs):
            print(name, "ok" if rc == 0 else "FAILED", msg, flush=True)


def run(out, extra):
    rows = []
    names = sorted(f[4:-3] for f in os.listdir(VDIR) if f.startswith("lib_g_") and f.endswith(".so"))
    for name in names:
        env = dict(os.environ, UNC_B200_LIB=os.path.join(VDIR, f"lib_{name}.so"))
        for tag, args in (("full", []), ("nt1", ["--nt", "1"])):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--steps", "5", "--warmup", "3"]
                               + args + extra, capture_output=True, text=True, env=env, timeout=600)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                rows.append(f"{name:24s} {tag:5s} ms_per_step {j['ms_per_step']:9.3f}  points/s {j['value']:.4g}  e2e {j['e2e']['value']:.4g}")
            except Exception:  # noqa: BLE001
                rows.append(f"{name:24s} {tag:5s} FAILED {r.stderr[-300:]}")
            print(rows[-1], flush=True)
    open(out, "w").write("\n".join(rows) + "\n")


def run_cmd(out, prefix, cmd):
    """`cmd` once per variant whose name starts with `prefix`; last stdout line of each is kept."""
    rows = []
    names = sorted(f[4:-3] for f in os.listdir(VDIR) if f.startswith("lib_" + prefix) and f.endswith(".so"))
    for name in names:
        env = dict(os.environ, UNC_B200_LIB=os.path.join(VDIR, f"lib_{name}.so"))
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
        last = (r.stdout.strip().splitlines() or ["FAILED " + r.stderr[-300:]])[-1]
        rows.append(f"{name:20s} {last}")
        print(rows[-1], flush=True)
    open(out, "w").write("\n".join(rows) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "cmd":      # cmd OUT.txt PREFIX command...
        run_cmd(sys.argv[2], sys.argv[3], sys.argv[4:])
    else:
        run(sys.argv[2], sys.argv[3:])
