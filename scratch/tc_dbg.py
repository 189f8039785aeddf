import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
def t(f, n=20):
    f(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
g = torch.Generator(device="cuda").manual_seed(0)
M = 65536
for (Nn, K) in [(256, 256), (512, 256), (256, 64)]:
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nn, device="cuda", generator=g)
    for dbg in (0, 8, 16, 32, 1, 4, 12):
        ctx.set_option("tc_debug", dbg)
        print(f"fwd N={Nn} K={K} dbg={dbg:2d}: {t(lambda: ctx.tc_linear(1, A, W, False, bias=b, variant=3)):.1f} us", flush=True)
ctx.set_option("tc_debug", 0)
