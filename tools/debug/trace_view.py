"""Prints the per-warp (tag, clock) timeline written by BBMPC_TC_TRACE (see csrc/rollout_tc.cu)."""
import collections, sys
fn = sys.argv[1]; warps = [int(x) for x in sys.argv[2].split(",")]; n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
rows = [l.split() for l in open(fn)]
by = collections.defaultdict(list)
for w, tag, clk in rows: by[int(w)].append((int(tag, 16), int(clk)))
t0 = min(v[0][1] for v in by.values())
for w in warps:
    print("warp", w, "records", len(by[w]))
    prev = None
    for tag, clk in by[w][:n]:
        print(f"  {tag:6x} {clk - t0:8d} {'' if prev is None else clk - prev}")
        prev = clk
