#!/usr/bin/env python
"""What the box moves between device and pinned host memory with N GPUs copying at once (plain cudaMemcpyAsync, no kernels).

    python tools/d2h_ceiling.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/d2h_ceiling.py

Every rank copies `--gib` GiB in 1 GiB chunks from device memory into a pinned ring (D2H), then the other way (H2D), then both
at once on two streams.  Barrier + device sync on both sides, CUDA events, max over ranks; rank 0 prints one JSON line with the
aggregate GB/s.  This is the denominator of bench.py's e2e figure (`e2e.d2h_ceiling_gbs` is measured inside the bench the same
way): on this pool's virtualised hosts it does not scale with the GPU count, which is what caps e2e at N > 1.
"""
import argparse
import json
import os


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=int, default=16)
    ap.add_argument("--slots", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chunk = 1 << 28  # floats: 1 GiB
    dev = torch.empty(args.slots * chunk, dtype=torch.float32, device="cuda").fill_(1.0)
    ring = [torch.empty(chunk, dtype=torch.float32).pin_memory() for _ in range(args.slots)]
    for r in ring:
        r.fill_(0.0)  # first touch by this rank's thread
    s2 = torch.cuda.Stream()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def d2h():
        for i in range(args.gib):
            ring[i % args.slots].copy_(dev[(i % args.slots) * chunk:(i % args.slots + 1) * chunk], non_blocking=True)

    def h2d():
        for i in range(args.gib):
            dev[(i % args.slots) * chunk:(i % args.slots + 1) * chunk].copy_(ring[i % args.slots], non_blocking=True)

    def both():
        s2.wait_stream(torch.cuda.current_stream())
        for i in range(args.gib):
            j = i % args.slots
            if i % 2 == 0:
                ring[j].copy_(dev[j * chunk:(j + 1) * chunk], non_blocking=True)
            else:
                with torch.cuda.stream(s2):
                    dev[j * chunk:(j + 1) * chunk].copy_(ring[j], non_blocking=True)

    total = world * args.gib * chunk * 4 / 1e9
    res = {"n_gpus": world, "gib_per_rank": args.gib, "d2h_GB/s": round(total / (timed(d2h) * 1e-3), 1), "h2d_GB/s": round(total / (timed(h2d) * 1e-3), 1),
           "both_GB/s": round(total / (timed(both) * 1e-3), 1), "host_cpus": os.cpu_count()}
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
