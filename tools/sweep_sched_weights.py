import os, sys, subprocess, json
sets = ["32,36,40,44", "32,37,41,45", "32,36,42,48", "32,35,40,44", "32,37,42,47", "32,36,39,42", "32,38,42,46"]
for w in sets:
    env = dict(os.environ, MMN_SCHED_WT_FWD=w, MMN_SCHED_WT_BWD=w)
    out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--steps", "6", "--warmup", "3", "--no-graph"], env=env, capture_output=True, text=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    k = d["roofline"]["kernels_ms"]
    print(w, "fwd %.4f bwd %.4f" % (k["winattn_fwd"], k["winattn_bwd"]), flush=True)
