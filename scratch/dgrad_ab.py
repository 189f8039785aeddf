"""Timing-only A/B of the TANH_BWD epilogue of the CTA-pair GEMM: tc_debug 8192 skips the activation (Hact) loads, 16384 the column sums."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
def t(f, n=30):
    f(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
g = torch.Generator(device="cuda").manual_seed(0)
M = 65536
for (Nn, K) in [(256, 512), (256, 256)]:
    sets = []
    for _ in range(3):
        A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(K, Nn, device="cuda", generator=g) / K ** 0.5
        Hact = torch.tanh(torch.randn(M, Nn, device="cuda", generator=g))
        sets.append((A, W, Hact))
    for dbg in (0, 8192, 16384, 8192 + 16384, 0):
        ctx.set_option("tc_debug", dbg)
        i = [0]
        def f():
            A, W, Hact = sets[i[0] % 3]; i[0] += 1
            ctx.tc_linear(2, A, W, True, Hact=Hact, colsum=True)
        print(f"dgrad N={Nn} K={K} tc_debug={dbg}: {t(f):.1f} us (incl. weight prep)", flush=True)
ctx.set_option("tc_debug", 0)
