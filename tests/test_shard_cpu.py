"""CPU tests of the multi-GPU partitioning (world_size 2, gloo): GOP segments of one stream are
assigned to ranks, decoded independently (the oracle stands in for the GPU engine on this box)
and gathered; the result equals the sequential decode = the reference golden."""
import os
import subprocess
import sys
import textwrap

import cases
from broadway_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_is_balanced_and_complete():
    cost = [90, 10, 50, 50, 30, 70, 5]
    plan = shard.assign(cost, 3)
    assert sorted(i for p in plan for i in p) == list(range(len(cost)))
    loads = [sum(cost[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(cost)
    assert shard.assign(cost, 3) == plan
    assert shard.assign([], 2) == [[], []]
    assert shard.assign([5], 4) == [[0], [], [], []]


def test_single_process_path(golden):
    import util
    from broadway_b200 import capi
    case = next(c for c in cases.SMALL if c[0] == "idr_period")
    segs = capi.split_gops(cases.make_stream(case))
    res = shard.decode_sharded(segs, lambda us: [util.oracle_md5(u)[0] for u in us], 0, 1)
    assert [m for r in res for m in r] == golden["idr_period"]["frame_md5"]


WORKER = textwrap.dedent('''
    import json, os, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import torch.distributed as dist
    import cases, util
    from broadway_b200 import capi, shard
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = next(c for c in cases.SMALL if c[0] == "idr_period")
    segs = capi.split_gops(cases.make_stream(case))
    res = shard.decode_sharded(segs, lambda us: [util.oracle_md5(u)[0] for u in us], rank, world, dist)
    golden = json.load(open(os.path.join(%(root)r, "tests", "golden", "streams.json")))["idr_period"]["frame_md5"]
    ok = [m for r in res for m in r] == golden
    mine = shard.assign([len(s) for s in segs], world)[rank]
    open(os.path.join(%(out)r, "rank%%d.json" %% rank), "w").write(json.dumps({"ok": ok, "units": mine}))
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if ok else 1)
''')


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    res = [json.load(open(tmp_path / ("rank%d.json" % k))) for k in range(2)]
    assert all(x["ok"] for x in res)
    # the three segments were split 2 + 1 across the ranks, every segment exactly once
    assert sorted(res[0]["units"] + res[1]["units"]) == [0, 1, 2] and {len(res[0]["units"]), len(res[1]["units"])} == {1, 2}
