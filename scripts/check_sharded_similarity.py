"""Multi-GPU check (run under torchrun): the pair set of `cosine_pairs_sharded` over the shards of all ranks equals the
single-matrix scan of the gathered matrix on rank 0; prints timings of both.
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/check_sharded_similarity.py --rows-per-gpu 65536"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=65536)
    ap.add_argument("--tau", type=float, default=0.9)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from facet_b200 import ops
    from facet_b200.utils.duplicate import cosine_pairs_sharded, gather_pairs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_local, d = args.rows_per_gpu, 768
    # identical data on every rank (same seed), near-duplicate clusters spread over all shards by a fixed permutation
    g = torch.Generator(device=dev).manual_seed(5)
    n = world * n_local
    base = torch.randn((n // 4, d), device=dev, generator=g)
    full = base[torch.randint(0, n // 4, (n,), device=dev, generator=g)] + 0.12 * torch.randn((n, d), device=dev, generator=g)
    full = torch.nn.functional.normalize(full, dim=1).contiguous()
    mine = full[rank * n_local:(rank + 1) * n_local].contiguous()
    for _ in range(2):
        pairs, sims = cosine_pairs_sharded(mine, args.tau)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pairs, sims = cosine_pairs_sharded(mine, args.tau)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    allp = gather_pairs(pairs)
    if rank == 0:
        e0.record()
        want, _ = ops.cosine_pairs(full, args.tau)
        e1.record()
        torch.cuda.synchronize()
        a = set(map(tuple, allp.tolist()))
        b = set(map(tuple, want.cpu().numpy().tolist()))
        print(f"world {world}: {n} rows, sharded scan {float(t):.2f} ms (max over ranks), single-GPU scan {e0.elapsed_time(e1):.2f} ms, "
              f"pairs {len(a)} vs {len(b)}, equal: {a == b}", flush=True)
        assert a == b and len(a) == allp.shape[0], "pair sets differ"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
