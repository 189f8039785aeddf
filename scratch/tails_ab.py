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
for (Nn, K) in [(256, 64), (256, 256), (512, 256)]:
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nn, device="cuda", generator=g)
    ws = torch.empty(ctx.lib.dppo_tc_linear_workspace_bytes(Nn, K), device="cuda", dtype=torch.uint8)
    out = torch.empty(M, Nn, device="cuda")
    ctx.tc_linear(1, A, W, False, bias=b, out=out, ws=ws)
    ref = torch.tanh(A.double() @ W.double().T + b.double())
    for dbg in (0, 128, 0, 128):
        ctx.set_option("tc_debug", dbg)
        us = t(lambda: ctx.tc_linear(1, A, W, False, bias=b, out=out, ws=ws, prepared=True))
        err = (out.double() - ref).abs().max().item()
        print(f"fwd N={Nn} K={K} tc_debug={dbg}: {us:.1f} us  err {err:.2e}", flush=True)
ctx.set_option("tc_debug", 0)
for (Nn, K) in [(256, 512), (256, 256)]:
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(K, Nn, device="cuda", generator=g) / K ** 0.5
    Hact = torch.tanh(torch.randn(M, Nn, device="cuda", generator=g))
    ref = (A.double() @ W.double()) * (1.0 - Hact.double() ** 2)
    for dbg in (0, 128, 0, 128):
        ctx.set_option("tc_debug", dbg)
        C, cs = ctx.tc_linear(2, A, W, True, Hact=Hact, colsum=True)
        err = (C.double() - ref).abs().max().item(); cs = cs[:cs.shape[0] // 5]; ce = (cs.double().sum(0) - ref.sum(0)).abs().max().item()
        us = t(lambda: ctx.tc_linear(2, A, W, True, Hact=Hact, colsum=True))
        print(f"dgrad N={Nn} K={K} tails={'full' if dbg else 'half'}: {us:.1f} us (incl. prep)  err {err:.2e} colsum {ce:.2e}", flush=True)
ctx.set_option("tc_debug", 0)
