import cProfile, pstats, io, sys, os
sys.path.insert(0, '/root/repo')
import torch
import strainer_b200 as sb
from oracle import strainer_oracle as O
dev = torch.device('cuda', 0)
netD = O.make_discriminator(O.SEED).to(dev).eval()
real = sb.synth_images(0, 128, O.SEED, dev)
fake = torch.randn(128, 3, 64, 64, device=dev)
def block():
    fr, ff, _, _ = sb.strain_batch(netD, real, 0.1)
    return sb.concat_fake(fake[:fr.shape[0]], ff)
for _ in range(20): block()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200): block()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
