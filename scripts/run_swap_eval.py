"""Stand-alone run of the swap-evaluation workload (poy5_b200.swap_eval; the `swap_eval` sub-record of bench.py).

    python scripts/run_swap_eval.py --prunings 64 --chunk 32
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_swap_eval.py --prunings 128
    config #5:  --taxa 1000 --loci 1 --lmin 10000 --lmax 10000 --seed 5 --prunings 32
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=200)
    ap.add_argument("--loci", type=int, default=5)
    ap.add_argument("--lmin", type=int, default=1000)
    ap.add_argument("--lmax", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--prunings", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--check", type=int, default=16)
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--merge-edges", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import poy5_b200 as pb
    from poy5_b200 import swap_eval
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pb.Context(local)
    out, sample = swap_eval.run(ctx, rank=rank, world=world, device=dev, taxa=a.taxa, nloci=a.loci, lmin=a.lmin, lmax=a.lmax,
                                seed=a.seed, prunings=a.prunings, chunk=a.chunk, check=a.check, lanes=a.lanes, merge_edges=a.merge_edges)
    if rank == 0:
        if sample is not None:
            from tests.oracle_backend import replay_sample
            t = time.perf_counter()
            out["parity"] = replay_sample(sample[0], sample[1], (1, 1, 3))
            out["parity"]["replay_s"] = time.perf_counter() - t
        out["ctx_stats"] = ctx.stats()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
