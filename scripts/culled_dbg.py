import torch, sys, inspect
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from deformationpyramid_b200 import _lib
import parity_cases as P
lib = _lib.load()
src = inspect.getsource(P.check_culled_search_equals_brute_force)
code = src.replace("assert torch.allclose(c0, c1, rtol=tol, atol=0), (c0, c1)", "print('rel diff per iteration', ((c0-c1).abs()/c1.abs()).amax(dim=0))")
ns = dict(P.__dict__); exec(code, ns)
f = ns["check_culled_search_equals_brute_force"]
f(lib, "cuda:0")
f(lib, "cuda:0", n=4200, m=4100, samples=4096, levels=2, iters=12)
f(lib, "cuda:0", n=8300, m=8250, samples=8192, levels=1, iters=6)
