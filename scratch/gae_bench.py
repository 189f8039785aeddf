import sys, os, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
import bench
from diamond import _native as N
ctx = N.get_context(0)
for variant in (0, 2, 1):
    ctx.set_option("gae_variant", variant)
    for _ in range(2):
        r = bench.bench_gae(ctx, 6543.4)
        print("variant", variant, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
