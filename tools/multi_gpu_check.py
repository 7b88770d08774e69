"""torchrun --nproc-per-node N tools/multi_gpu_check.py
Every rank strains its shard over NCCL; rank 0 also strains the whole dataset alone and checks that the
global threshold and the concatenated kept-index list are BIT-IDENTICAL to the single-GPU result
(BASELINE.json north_star: 'identical to single-GPU results')."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    per = 12288
    n = per * world
    netD = O.make_discriminator(O.SEED)
    results = {}
    comm = sb.PeerComm.for_group(dist.group.WORLD, device)
    assert comm is not None, "NVLink peer buffers could not be mapped"
    for mode in ("auto", "fp32", "bf16"):
        imgs = sb.synth_images(rank * per, per, O.SEED, device)
        idx, thr, losses = sb.strain_shard(imgs, netD, 0.1, group=dist.group.WORLD, index_base=rank * per,
                                           n_global=n, conv_mode=mode, device=device)            # NVLink peer all-reduce
        idx_n, thr_n, _ = sb.strain_shard(imgs, netD, 0.1, group=dist.group.WORLD, index_base=rank * per,
                                          n_global=n, conv_mode=mode, device=device, comm=None)  # NCCL all-reduces
        assert thr == thr_n and np.array_equal(idx, idx_n), "peer-memory and NCCL selects disagree"
        gathered = [None] * world
        dist.all_gather_object(gathered, (idx, thr))
        if rank == 0:
            all_idx = np.concatenate([g[0] for g in gathered])
            thrs = [g[1] for g in gathered]
            assert all(t == thrs[0] for t in thrs), thrs
            full = sb.synth_images(0, n, O.SEED, device)
            idx1, thr1, _ = sb.strain_shard(full, netD, 0.1, conv_mode=mode, device=device)
            assert thr1 == thrs[0], (thr1, thrs[0])
            assert np.array_equal(all_idx, idx1)
            results[mode] = (float(thr1), len(idx1), n)
    # sharded deterministic GMM EM (8-double all-reduce per iteration): every rank holds the identical fit, and it
    # agrees with the single-GPU fit of the whole vector to fp64 summation-order accuracy
    v = O.synth_losses(40000 * world, seed=7)
    shard = torch.from_numpy(v[rank * 40000:(rank + 1) * 40000]).to(device)
    g = sb.gmm_fit_device(shard, group=dist.group.WORLD, n_global=v.size)
    allg = [None] * world
    dist.all_gather_object(allg, {k: np.asarray(g[k]).tolist() for k in ("weights", "means", "stds")})
    if rank == 0:
        assert all(a == allg[0] for a in allg), allg
        g1 = sb.gmm_fit_device(torch.from_numpy(v).to(device))
        for k in ("weights", "means", "stds"):
            assert np.allclose(g[k], g1[k], rtol=1e-9), (k, g[k], g1[k])
        assert g["n_iter"] == g1["n_iter"]
        results["gmm"] = (g["means"].tolist(), g["n_iter"])
    # adversarial vectors through the peer path: ties at the threshold, NaN, all-equal, tiny shards, every quantile
    rng = np.random.default_rng(5)
    cases = {"ties": np.round(rng.standard_normal(50000 * world), 1).astype(np.float32),
             "equal": np.full(4097 * world, 0.25, np.float32),
             "nan": np.where(rng.random(3000 * world) < 1e-3, np.nan, rng.random(3000 * world)).astype(np.float32),
             "tiny": rng.random(3 * world).astype(np.float32)}
    for name, v in cases.items():
        m = v.size // world
        shard = torch.from_numpy(v[rank * m:(rank + 1) * m]).to(device)
        for q in (0.0, 10.0, 50.0, 90.0, 99.9, 100.0):
            got = sb.percentile_device(shard, q, dist.group.WORLD, v.size, comm=comm).cpu().numpy()[0]
            want = np.percentile(v, q)
            assert (np.isnan(got) and np.isnan(want)) or got == want, (name, q, got, want)
    # latency of the two select paths on this rank's 12288-loss shard (CUDA events, median of 50)
    times = {}
    for tag, c in (("peer", comm), ("nccl", None)):
        evs = []
        for i in range(60):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sb.percentile_device(losses, 90.0, dist.group.WORLD, n, comm=c)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        times[tag] = float(np.median([a.elapsed_time(b) for a, b in evs[10:]]))
    if rank == 0:
        results["select_ms"] = times
        print("multi_gpu_check OK", world, "ranks", results, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
