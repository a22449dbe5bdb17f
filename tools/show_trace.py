"""Print a trace written by tools/trace_fwd.py / trace_bwd.py: per-role phase deltas (cycles) and the
spread of per-CTA run times (ns).  Usage: python tools/show_trace.py gpurun_out/trace_fwd.txt [first last]"""
import sys

path = sys.argv[1]
first, last = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 20)
ev, ctas = {}, []
for line in open(path):
    t = line.split()
    if t[0] == "cta":
        ctas.append((int(t[1]), int(t[2]), int(t[3])))
    else:
        ev[(int(t[0]), int(t[1]))] = [int(x) for x in t[2:]]
t0 = min(v for k in ev for v in ev[k] if v > 0)
for role in range(5):
    print("role", role)
    prev_last = None
    for item in range(first, last):
        e = ev.get((role, item))
        if not e or not any(e):
            continue
        pairs = sorted((x, k) for k, x in enumerate(e) if x > 0)
        st = [x for x, _ in pairs]
        deltas = [f"{k}:{b - a}" for (a, _), (b, k) in zip(pairs, pairs[1:])]
        gap = st[0] - prev_last if prev_last else 0
        prev_last = st[-1]
        print(f"  item {item:3d} start {st[0] - t0:7d} gap {gap:5d} phases {" ".join(deltas)} total {st[-1] - st[0]}")
if ctas:
    durs = sorted(b - a for _, a, b in ctas)
    s0 = min(a for _, a, _ in ctas)
    e1 = max(b for _, _, b in ctas)
    print(f"ctas {len(ctas)}: run ns min {durs[0]} median {durs[len(durs) // 2]} max {durs[-1]}; kernel span {e1 - s0} ns")
    ends = sorted((b - s0, c) for c, _, b in ctas)
    print("  earliest ends", ends[:4], "latest ends", ends[-4:])
