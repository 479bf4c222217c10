"""CPU, world_size 2 over gloo: the N>1 host logic -- token-balanced contiguous query shards that cover every
query exactly once, and the shape/array broadcast plumbing used before the NCCL index broadcast."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_queries():
    from cgx_b200.dist import shard_queries
    rng = np.random.default_rng(0)
    for Q in (0, 1, 2, 7, 100, 1001):
        lens = rng.integers(0, 60, size=Q)
        off = np.concatenate([[0], np.cumsum(lens)])
        for world in (1, 2, 3, 8):
            cuts = [shard_queries(off, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == Q
            for a, b in zip(cuts[:-1], cuts[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            if Q >= 100:
                toks = [off[e] - off[b] for b, e in cuts]
                assert max(toks) - min(toks) <= 2 * 60 + 1


def test_gloo_world2_broadcast_and_shards(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        from cgx_b200.dist import shard_queries, broadcast_shape
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        vec = torch.zeros(104, dtype=torch.int64)
        if rank == 0:
            vec[:4] = torch.tensor([1000, 900, 77, 321]); vec[4:] = torch.arange(100)
        broadcast_shape(vec, 0)
        assert vec[:4].tolist() == [1000, 900, 77, 321] and vec[4:].tolist() == list(range(100))
        arr = torch.arange(1000, dtype=torch.int32) if rank == 0 else torch.zeros(1000, dtype=torch.int32)
        dist.broadcast(arr, src=0)
        assert int(arr.sum()) == 999 * 1000 // 2
        off = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(1, 40, size=501))])
        b, e = shard_queries(off, world, rank)
        cnt = torch.tensor([e - b, int(off[e] - off[b])])
        dist.all_reduce(cnt)
        assert cnt.tolist() == [501, int(off[-1])]
        dist.barrier()
        print("rank", rank, "ok", b, e)
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
