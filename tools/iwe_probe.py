"""Timing probe of the event-warping loss kernels (graph replays): python tools/iwe_probe.py; switches through the environment
(EF_IWE_COOP=0 plain launch, EF_IWE_SKIP=mask phase ablation, EF_IWE_GRID=n forced grid) -- one process per setting."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, bench
    pk, _ = bench.peaks()
    r = bench.iwe_bench(torch.device("cuda"), pk["hbm_gbs"], reps=20)
    print("RESULT " + json.dumps({k: (round(v["fwd_ms"] * 1e3, 1), round((v["fwd_bwd_ms"] - v["fwd_ms"]) * 1e3, 1)) for k, v in r.items()}))
else:
    for env in ({}, {"EF_IWE_SKIP": "1"}, {"EF_IWE_SKIP": "2"}, {"EF_IWE_SKIP": "4"}, {"EF_IWE_SKIP": "7"}):
        out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        print(env, res[0][7:] if res else out.stderr[-400:], "(fwd us, bwd us)", flush=True)
