"""BASELINE.json configs[3]: all-pairs cosine similarity over N 768-d embeddings (planted near-duplicate clusters),
sharded across the ranks of one node in shard-against-shard blocks (utils/duplicate.py cosine_pairs_sharded); the bf16 /
float32 shards are all-gathered with NCCL inside the timed region, behind the scan of each rank's own block.

    python scripts/bench_similarity.py --embeddings 1000000            (1 GPU: the whole upper triangle)
    torchrun --nproc-per-node 8 ... scripts/bench_similarity.py --embeddings 1000000

Prints one JSON line: pairs/s over the N(N-1)/2 upper triangle, TFLOP/s on N(N-1)*768 flops, pair count.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--embeddings", dest="n", type=int, default=1_000_000)
    ap.add_argument("--tau", type=float, default=0.90)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from facet_b200 import ops
    from facet_b200.utils.duplicate import cosine_pairs_sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d = args.n, 768
    n_local = n // world
    # planted structure: clusters of 2..6 near-duplicates (e + 0.05 * noise, renormalised) over 20 % of the rows
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    e = torch.randn((n_local, d), device=dev, generator=g)
    n_cl = n_local // 20
    members = torch.randint(0, n_cl, (n_local // 5,), device=dev, generator=g)
    e[n_local - n_local // 5:] = e[members] + 0.05 * torch.randn((n_local // 5, d), device=dev, generator=g) / (d ** 0.5) * (d ** 0.5) * 0.2
    e = torch.nn.functional.normalize(e, dim=1)

    def step():
        # world > 1: shard-against-shard blocks; the rank's own diagonal block runs while NCCL gathers the other shards
        return cosine_pairs_sharded(e, args.tau)

    pairs, _ = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        pairs, _ = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    t = torch.tensor([ms, float(pairs.shape[0])], device=dev, dtype=torch.float64)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, npairs = float(mx[0]), int(sm[1])
    else:
        npairs = int(t[1])
    if rank == 0:
        flops = float(n) * (n - 1) * d
        print(json.dumps({"workload": "all-pairs cosine >= tau over N x 768 f32 embeddings (bf16 tensor-core scan + exact f32 recheck)",
                          "n": n, "tau": args.tau, "n_gpus": world, "ms": ms, "pairs_found": npairs,
                          "pair_comparisons_per_s": n * (n - 1) / 2 / (ms * 1e-3),
                          "tflops_upper_triangle": flops / (ms * 1e-3) / 1e12}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
