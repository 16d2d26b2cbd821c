"""Hash-sharded build over 2 real GPUs (NCCL): merged result == the oracle's table.
Skipped on a one-GPU box; the same host logic is covered on CPU by tests/test_dist_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

from tests.helpers import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KTG_ROOT"])
import numpy as np, torch, torch.distributed as dist
from katome_b200.dist import ShardedGIR
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
for k, n, L, G in ((31, 6000, 150, 200_000), (63, 3000, 150, 100_000)):
    reads = O.synth_reads(99, G, L, 5000, 0, n)
    cpu = O.OracleGIR(k)
    cpu.add_reads(reads, np.arange(n + 1, dtype=np.uint64) * L, True)
    half = n // world
    mine = torch.from_numpy(reads[rank * half * L:(rank + 1) * half * L]).cuda()
    offs = torch.arange(0, (half + 1) * L, L, dtype=torch.int64, device="cuda")
    for kw in ({}, {"exchange": "direct"}, {"exchange": "keys"}, {"exchange": "skm"}, {"fused": False},
               {"force_pages": True, "exchange": "skm"}, {"force_pages": True, "exchange": "direct"},
               {"force_pages": True, "exchange": "keys"},
               {"fused": False, "force_partition": True, "sub_table_log2_bytes": 16}):
        sg = ShardedGIR(k, True, **kw)
        # super-k-mer records need 23 <= k <= 31; below 8 ranks they are opt-in
        assert sg.exchange == ("nccl" if kw.get("fused") is False else "skm" if k <= 31 and kw.get("exchange") == "skm" else "keys")
        # the direct exchange is the default below 4 ranks; "keys" / "skm" / fused=False ask for another one
        assert sg.direct == (kw.get("exchange") == "direct" or (kw.get("exchange") is None and kw.get("fused") is not False))
        for _ in range(2):  # reset + rebuild gives the same table
            sg.reset()
            sg.add_reads_device(mine[: (half // 2) * L], offs[: half // 2 + 1], half // 2, (half // 2) * L)
            sg.add_reads_device(mine[(half // 2) * L:], offs[: half - half // 2 + 1], half - half // 2, (half - half // 2) * L)
            sg.finalize()
            assert sg.digest() == cpu.digest(), (rank, k, kw, sg.digest(), cpu.digest())
        # the same reads from pinned host memory, copied in chunks behind the exchange
        sg.reset()
        hb = torch.from_numpy(reads[rank * half * L:(rank + 1) * half * L].copy()).pin_memory()
        ho = torch.arange(0, (half + 1) * L, L, dtype=torch.int64).pin_memory()
        sg.MIN_CHUNK_READS = 500
        sg.add_reads_host(hb, ho, half, chunks=3)
        sg.finalize()
        assert sg.digest() == cpu.digest(), (rank, k, kw, "host path")
        assert sg.last_exchange == ("direct" if sg.direct else sg.exchange), (sg.last_exchange, kw)
        a, b = sg.collection_stats(), cpu.collection_stats()
        for key in b:
            assert a[key] == b[key] or (a[key] != a[key] and b[key] != b[key]), (rank, k, key, a[key], b[key])
        assert sg.counts() == cpu.counts()
        sg.remove_weak_edges(3)
        cpu2 = O.OracleGIR(k)
        cpu2.add_reads(reads, np.arange(n + 1, dtype=np.uint64) * L, True)
        cpu2.remove_weak_edges(3)
        assert sg.digest() == cpu2.digest()
        assert sg.counts() == cpu2.counts()
        sg.standardize_edges(G, k, 3)
        cpu2.standardize_edges(G, k, 3)
        assert sg.digest() == cpu2.digest() and sg.counts() == cpu2.counts()
        sg.close()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_gpu_sharded_build_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KTG_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_one_handle_over_two_real_gpus(tmp_path):
    """ktg_config.n_devices over DISTINCT devices (peer memory over NVLink, no torch.distributed, one
    process): against the oracle and against the same build on one GPU, exports included."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import katome_b200 as K
    from oracle import oracle as O
    from tests import helpers as H
    n_dev = min(torch.cuda.device_count(), 8)
    for k, ids in ((31, [0, 1]), (63, [1, 0]), (31, list(range(n_dev))), (40, list(range(n_dev)))):
        n, L, G = 40_000, 150, 300_000
        reads = O.synth_reads(1234 + k, G, L, 5000, 0, n)
        offsets = np.arange(n + 1, dtype=np.uint64) * L
        cpu = O.OracleGIR(k)
        cpu.add_reads(reads, offsets, True)
        g = K.GpuGIR(k, True, device_ids=ids, options={"chunk_mb": 1})
        assert g.add_reads(reads, offsets) == (cpu.accepted_reads, cpu.accepted_bytes)
        assert g.digest() == cpu.digest() and g.counts() == cpu.counts()
        one = K.GpuGIR(k, True, device=0)
        one.add_reads(reads, offsets)
        ga, gb = g.export_graph(), one.export_graph()
        for name in ga:
            assert np.array_equal(ga[name], gb[name]), (k, ids, name)
        for a, b in zip(g.export_externals(), one.export_externals()):
            assert np.array_equal(a, b)
        a, b = g.collection_stats(), cpu.collection_stats()
        for key in b:
            assert a[key] == b[key] or (a[key] != a[key] and b[key] != b[key]), (k, ids, key)
        g.remove_weak_edges(3), cpu.remove_weak_edges(3)
        g.standardize_edges(G, k, 3), cpu.standardize_edges(G, k, 3)
        assert g.digest() == cpu.digest() and g.counts() == cpu.counts()
        g.close(), one.close()
