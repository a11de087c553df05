"""Where the end-to-end path spends its time: shard.evaluate(Registration, host=True) on 3 batches of the headline shape,
with wall-clock timers around the host preparation, the native call and the post-processing.
    python scripts/e2e_phases.py [pairs] [iters]"""
import sys, time
sys.path.insert(0, ".")
import torch
from deformationpyramid_b200 import shard
from deformationpyramid_b200.config import ndp_config
from deformationpyramid_b200.model.registration import Registration
from deformationpyramid_b200.synthetic import make_pair

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 500
N = 8192
cfg = ndp_config(samples=N, m=9, iters=iters, max_break_count=10 ** 9, device=0)
reg = Registration(cfg)
T = {"prep": 0.0, "run": 0.0}
_p, _r = reg._prepare_batch, reg._run_prepared
def prep(*a, **k):
    t = time.perf_counter(); out = _p(*a, **k); T["prep"] += time.perf_counter() - t; return out
def run(*a, **k):
    t = time.perf_counter(); out = _r(*a, **k); T["run"] += time.perf_counter() - t; return out
reg._prepare_batch, reg._run_prepared = prep, run
pairs = {}
def get_item(i):
    g = i % B
    if g not in pairs:
        s, t = make_pair(g, N, N); pairs[g] = dict(src_pcd=s.numpy(), tgt_pcd=t.numpy())
    return pairs[g]
for i in range(B): get_item(i)
shard.evaluate(reg, B, get_item, batch=B, compute_metrics=False, host=True)
for k in T: T[k] = 0.0
torch.cuda.synchronize()
t0 = time.perf_counter()
shard.evaluate(reg, 3 * B, get_item, batch=B, compute_metrics=False, host=True, checksum=True)
tot = time.perf_counter() - t0
print(f"total {tot:.3f} s for {3 * B} pairs = {3 * B / tot:.2f} pairs/s; host prep (sum over batches, overlapped after the first) "
      f"{T['prep']:.3f} s, native calls {T['run']:.3f} s, rest {tot - T['run']:.3f} s")
