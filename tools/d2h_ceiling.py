"""Host-ingest ceiling of the box: every rank copies a result-sized block (default 1.8 GB, one C2 batch) device -> pinned host
memory with plain cudaMemcpyAsync, all ranks at once.  The aggregate GB/s is what bounds bench.py's end-to-end number at N GPUs
(results are D2H-copied every step): e2e queries/s <= ceiling / (d2h_bytes_per_step / queries per step).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/d2h_ceiling.py [MB] [reps]
    python tools/d2h_ceiling.py            (one GPU)

Rank 0 prints one JSON line: per-rank and aggregate GB/s for D2H alone, and for D2H while a copy kernel keeps HBM busy."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1800
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = mb << 20
    src = torch.empty(n, dtype=torch.uint8, device=dev).fill_(7)
    dst = torch.empty(n, dtype=torch.uint8).pin_memory()
    busy_a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    busy_b = torch.empty_like(busy_a)
    copy_stream = torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(with_kernels):
        dst.copy_(src)                                       # warm: pinned pages touched
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(copy_stream):
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
        if with_kernels:
            while not copy_stream.query():
                busy_b.copy_(busy_a)
        copy_stream.synchronize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    alone = run(False)
    loaded = run(True)
    if rank == 0:
        tot = world * reps * n / 1e9
        print(json.dumps({"n_gpus": world, "mb_per_copy": mb, "copies_per_rank": reps, "host_cpus": os.cpu_count(),
                          "d2h_aggregate_gbs": tot / alone, "d2h_per_rank_gbs": tot / alone / world,
                          "d2h_aggregate_gbs_while_kernels_run": tot / loaded, "seconds": alone, "seconds_while_kernels_run": loaded}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
