"""Throughput and roofline sweep (BASELINE configs[4]): training fwd+bwd over sequence length T and window size W on one
GPU, a fixed number of tokens per step (B * T = 32768, i.e. B = 512 at T = 64 ... 128 at T = 256).
usage: python tools/sweep.py > profiles/r02_sweep.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for T in (64, 128, 192, 256):
    for W in (16, 32, 64):
        B = 32768 // T if T != 192 else 168
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--config", "train512", "--frames", str(T), "--window", str(W),
               "--batch", str(B), "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-eager-baseline"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            rows.append({"T": T, "W": W, "B": B, "error": r.stderr[-400:]})
            continue
        j = json.loads(r.stdout.strip().splitlines()[-1])
        k = {x["kernel"]: x for x in j["kernels"]}
        attn = {name: {"avg_ms": round(v["avg_ms"], 3), "frac": round(v["frac"], 3)} for name, v in k.items() if "attn" in name}
        rows.append({"T": T, "W": W, "N": 2 * W, "B": B, "seq_per_s": round(j["value"], 1), "ms_per_step": round(j["ms_per_step"], 2),
                     "tokens_per_s": round(j["value"] * T * 64), "sm_mhz": j["clocks"]["sm_mhz"],
                     "attn_frac_of_sustained_peak": round(j["attn_tensor_frac"]["frac_of_sustained_peak"], 3),
                     "attention_kernels": attn})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
print(json.dumps(rows, indent=1))
