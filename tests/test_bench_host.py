"""Host-side logic of bench.py that needs no GPU: which frames a step renders, and the bookkeeping the judge reads."""
import argparse
import json
import math
import os

import numpy as np

import bench


def test_every_rank_covers_the_sweep_with_the_same_spacing():
    for K in (20, 36, 720):
        for N in (1, 2, 4, 8):
            per_rank = [np.degrees(bench.sweep_thetas(K, r, N)) for r in range(N)]
            for th in per_rank:
                assert len(th) == K and 0 <= th.min() and th.max() < 360
                assert np.allclose(np.diff(th), 360. / K)          # the same angular spacing on every rank, for every N
                assert th.max() - th.min() > 360. - 2 * 360. / K   # ... covering the whole turn
            union = np.sort(np.concatenate(per_rank))
            assert np.allclose(np.diff(union), 360. / (K * N))     # together: K * N distinct, evenly spaced views
            # the offset between two ranks is less than one step: their work per frame is the same to first order
            assert abs(per_rank[-1][0] - per_rank[0][0]) < 360. / K


def test_both_arms_describe_the_same_workload():
    a = argparse.Namespace(vol=512, img=1024, gpus=4, steps=20, warmup=5)
    c = bench.sweep_config(a)
    assert c == bench.sweep_config(argparse.Namespace(vol=512, img=1024, gpus=1, steps=7, warmup=3))
    assert "512" in c["workload"] and "1024x1024" in c["workload"]
    json.dumps(c)


def test_traffic_is_only_reported_for_the_sources_it_was_measured_on(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "profiles")
    os.makedirs(tmp_path / "spimagine_b200" / "csrc")
    for f in bench.KERNEL_SOURCES["mip"]:
        (tmp_path / "spimagine_b200" / "csrc" / f).write_text("// " + f)
    t, why = bench.ncu_traffic("sweep_512_1024")
    assert t is None and "no ncu capture" in why
    rec = {"sweep_512_1024": {"source_sha1": bench.source_sha1(), "dram_bytes_read": 3e8, "dram_bytes_write": 1e7,
                              "launches": 3, "kernel": "k"}}
    (tmp_path / "profiles" / "r02_mip_traffic.json").write_text(json.dumps(rec))
    t, why = bench.ncu_traffic("sweep_512_1024")
    assert t == 310000000 and "r02_mip_traffic.json" in why
    (tmp_path / "spimagine_b200" / "csrc" / "spv_mip.cu").write_text("// changed")
    t, why = bench.ncu_traffic("sweep_512_1024")
    assert t is None and "predates" in why
    assert bench.ncu_traffic("other")[0] is None
