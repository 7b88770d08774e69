"""Where the in-batch strain block's latency goes (host + device), B = 128 by default."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strainer_b200 as sb
from oracle import strainer_oracle as O

def t(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

def t_sync_each(fn, n=200, warm=20):
    for _ in range(warm): fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); tot += time.perf_counter() - t0
    return tot / n * 1e6

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
netD = O.make_discriminator(O.SEED).eval()
real = sb.synth_images(0, B, O.SEED, dev)
fake = torch.randn(B, 3, 64, 64, device=dev)
sc = sb.get_scorer(netD, dev, "bf16", max_batch=max(B, 512))
prob = torch.empty(B, device=dev)
out = {"B": B}
out["get_scorer_us"] = t(lambda: sb.get_scorer(netD, dev, "bf16", max_batch=max(B, 512)))
out["score_into_async_us"] = t(lambda: sc.score_into(real, None, prob, None))
out["score_into_sync_each_us"] = t_sync_each(lambda: sc.score_into(real, None, prob, None))
for layer in range(1, 6):
    out[f"layer{layer}_sync_each_us"] = t_sync_each(lambda: sc.run_layer(real, layer, None, prob, None))
out["strain_scores_us"] = t(lambda: sb.strain_scores(real, prob, 0.1))
fr, ff, _, _ = sb.strain_scores(real, prob, 0.1)
out["concat_fake_us"] = t(lambda: sb.concat_fake(fake[:fr.shape[0]], ff))
out["torch_cat_us"] = t(lambda: torch.cat([fake[:fr.shape[0]], ff], 0))
out["strain_batch_us"] = t(lambda: sb.strain_batch(netD, real, 0.1, conv_mode="bf16"))
out["empty_sync_us"] = t_sync_each(lambda: None)
netDg = O.make_discriminator(O.SEED).eval().to(dev)
def eager():
    with torch.no_grad():
        s = netDg(real).view(-1); thr = torch.quantile(s, 0.1); m = s >= thr
        return real[m], real[~m]
out["torch_eager_block_us"] = t(eager)
print(json.dumps(out, indent=1))
