"""Timing-only A/B of the BIAS_TANH epilogue of the CTA-pair GEMM: tc_debug 32 = bias add without tanh, 16 = no TMA stores,
8 = no epilogue work at all (wrong results)."""
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
for (Nn, K) in [(256, 64), (256, 256), (512, 256)]:
    sets = []
    for _ in range(3):
        A = torch.randn(M, K, device="cuda", generator=g); out = torch.empty(M, Nn, device="cuda")
        sets.append((A, out))
    W = torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nn, device="cuda", generator=g)
    ws = torch.empty(ctx.lib.dppo_tc_linear_workspace_bytes(Nn, K), device="cuda", dtype=torch.uint8)
    ctx.tc_linear(1, sets[0][0], W, False, bias=b, out=sets[0][1], ws=ws)
    for dbg in (0, 32, 16, 32 + 16, 8, 0):
        ctx.set_option("tc_debug", dbg)
        i = [0]
        def f():
            A, out = sets[i[0] % 3]; i[0] += 1
            ctx.tc_linear(1, A, W, False, bias=b, out=out, ws=ws, prepared=True)
        print(f"fwd N={Nn} K={K} tc_debug={dbg}: {t(f):.1f} us", flush=True)
ctx.set_option("tc_debug", 0)
