import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
for pair in (0, 1):
    for bf16 in (0, 1):
        for n in (64, 128, 256):
            c = ctx.mma_probe(pair, bf16, n, 2000)
            print(f"pair={pair} bf16={bf16} N={n}: clk/MMA mean {c.mean().item():.1f} min {c.min().item():.1f} max {c.max().item():.1f} (n={c.numel()})", flush=True)
