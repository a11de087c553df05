"""Debug: per-phase globaltimer stamps of CTA (0,0) of the tensor-core kernels (last launch of a short solver run).
    python scripts/phase_times.py [pairs] [streams] [tpc] [fwd_rounds]"""
import ctypes, sys
sys.path.insert(0, '.')
import torch
from deformationpyramid_b200 import ops, _lib
from deformationpyramid_b200.synthetic import make_pair
from oracle import ndp_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tpc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = 'cuda:0'
lib = _lib.load()
specs = O.make_specs(3, 128, -8, 1, "axis_angle")
pairs = [make_pair(p, 8192, 8192) for p in range(B)]
solver = ops.Solver(max_pairs=B, max_src_points=8192, max_tgt_points=8192, samples=8192, levels=1, k0=-8, depth=3, width=128,
                    motion="SE3", rotation_format="axis_angle", iters=int(sys.argv[5]) if len(sys.argv) > 5 else 20, max_break_count=10**9, break_threshold_ratio=0.001, lr=0.01,
                    streams=streams, tiles_per_bwd_cta=tpc, fwd_rounds=rounds)
torch.manual_seed(0)
flats = [torch.cat([O.flatten_params(s, O.init_params(s)) for s in specs]).to(dev) for _ in range(B)]
solver.register([s.to(dev) for s, _ in pairs], [t.to(dev) for _, t in pairs], flats)
buf = (ctypes.c_ulonglong * 64)()
lib.ndp_debug_phase_times.argtypes = [ctypes.c_int, ctypes.c_void_p]
print(f"pairs={B} streams={streams} tpc={tpc} rounds={rounds}")
for which, name in ((0, 'fwd'), (2, 'bwd_rc (chain 0: [0..12], chain 1: [24..36], issuer [40..57]: F1 c0 40-41 c1 42-43 reload 44; F2 c0 45/46/47 c1 48/49/50; B1 c0 51/52/53 c1 54/55/56 reload 57)')):
    lib.ndp_debug_phase_times(which, buf)
    t = [int(v) for v in buf]
    t0 = t[0]
    print(name, ' '.join(f"[{i}]{(v - t0)/1000:.2f}" for i, v in enumerate(t[:62]) if v))
    print(name.split()[0], f"CTA(0,*) duration of the last launch to finish: {t[62]/1000:.2f} us")
    a, b = (58, 59) if which == 0 else (60, 61)
    if t[b]:
        print(name.split()[0], f"mean CTA duration over all {t[b]} CTAs of the run: {t[a]/t[b]/1000:.2f} us")
