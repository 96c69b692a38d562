"""world_size-2 tests of the multi-GPU plumbing on CPU (gloo): the shards partition
a batch, the timing reduction is a max over ranks, and the whole-job rate counts
every rank's units.  The data path itself has no collective."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from gid_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_images = 37
        mine = sharding.shard(n_images, rank, world)
        # every rank reports which images it owns; the union must be the batch, disjointly
        flags = torch.zeros(n_images, dtype=torch.int32)
        flags[mine] = 1
        dist.all_reduce(flags, op=dist.ReduceOp.SUM)
        assert flags.tolist() == [1] * n_images
        assert all(sharding.owner(i, world) == rank for i in mine)
        # pretend rank r needed (10 + 5 r) ms: the job's time is the slowest rank's
        ms = 10.0 + 5.0 * rank
        dist.barrier()
        ms_max = sharding.max_over_ranks(ms)
        assert ms_max == 10.0 + 5.0 * (world - 1)
        units = float(len(mine))
        total = sharding.sum_over_ranks(units)
        assert total == n_images
        rate = sharding.whole_job_rate([units] if world == 1 else [total / world] * world, world, ms_max * 1e-3)
        with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
            json.dump({"rank": rank, "mine": mine, "ms_max": ms_max, "rate": rate}, f)
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    world, port = 2, 29517 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(world)]
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(37))
    assert res[0]["ms_max"] == res[1]["ms_max"] == 15.0
    assert res[0]["rate"] == pytest.approx(37 / 0.015)


def test_shard_is_round_robin():
    from gid_b200 import sharding
    assert sharding.shard(10, 1, 4) == [1, 5, 9]
    assert sharding.shard(3, 3, 4) == []
    assert sum(len(sharding.shard(1024, r, 8)) for r in range(8)) == 1024
    with pytest.raises(ValueError):
        sharding.shard(4, 4, 4)


def test_reference_arm_under_torchrun_rank0_only():
    """bench.py --impl reference launched like the driver does for N = 2: rank 0 prints the one
    JSON line (impl = reference, cpu_baseline, e2e with zero copy bytes); rank 1 exits 0 silently."""
    env = dict(os.environ, GIDBENCH_REFERENCE_SECONDS="0.3")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29800 + os.getpid() % 100),
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--workload", "cfg1"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_numa_binding_is_best_effort():
    """bench.py binds each rank to its GPU's NUMA node when it can; without NVML / sysfs information
    (as here) it says so and leaves the affinity alone."""
    import bench
    before = os.sched_getaffinity(0)
    msg = bench.bind_to_gpu_numa_node(0)
    assert isinstance(msg, str) and msg
    if not msg.startswith("bound"):
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
