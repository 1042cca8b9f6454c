"""N > 1 path on the CPU: two ranks (gloo), each an independent engine over its contig-range shard, no collective on the data path,
then the host-side gather of calitas_b200.multi.  The gathered table must equal the single-engine table byte for byte.
The engine here is the test-only host simulation (no GPU in this tier); the -m gpu twin runs the product library (test_sharding.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lib_path, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from calitas_b200 import synth, multi
    from calitas_b200._capi import Engine, Library, Limits
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        g = synth.config1_genome(scale=0.02, n_sites=80)
        contigs = [(n, b) for n, b in g.contigs()]
        guides = [synth.BASELINE_GUIDE, ("CTTGCCCCACAGGGCAGTAAngg", ["nag"]), "GGGGCCACTAGGGACAGGAT"]
        lim = Limits(5, 1, 3, -1, 10)
        e = Engine(0, lib=Library(lib_path))
        ref = e.load_reference(contigs, shard=(rank, world, 4000))
        local = e.search(ref, guides, lim, dedup=True).records()
        merged = multi.gather_hits(local, dst=0)
        if rank == 0:
            whole_ref = e.load_reference(contigs)
            whole = e.search(whole_ref, guides, lim, dedup=True).records()
            np.save(out_path, np.array([merged.tobytes() == whole.tobytes(), merged.size, whole.size, local.size]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_shards_gather_to_the_single_engine_table(tmp_path, world):
    import torch.multiprocessing as mp
    import backends
    lib_path = backends.get("hostsim").t.lib.path
    out = str(tmp_path / "result.npy")
    mp.spawn(_worker, args=(world, _free_port(), lib_path, out), nprocs=world, join=True)
    same, n_merged, n_whole, n_local0 = np.load(out)
    assert n_whole > 60 and n_merged == n_whole and 0 < n_local0 < n_whole
    assert same == 1


def test_merge_shard_records_orders_guide_major():
    from calitas_b200 import multi
    from calitas_b200._capi import hit_dtype
    a = np.zeros(3, dtype=hit_dtype()); a["guide_idx"] = [0, 0, 1]; a["guide_start_offset"] = [5, 9, 2]
    b = np.zeros(2, dtype=hit_dtype()); b["guide_idx"] = [0, 1]; b["guide_start_offset"] = [100, 50]
    m = multi.merge_shard_records([a, b])
    assert m["guide_idx"].tolist() == [0, 0, 0, 1, 1] and m["guide_start_offset"].tolist() == [5, 9, 100, 2, 50]
    assert multi.merge_shard_records([]).size == 0
