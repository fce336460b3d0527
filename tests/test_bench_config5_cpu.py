"""bench.py's config-5 driver (256 concurrent requests: untimed warm-up server, three serving modes, gather) run on
the CPU against the fake model of tests/test_serving.py: the host logic of the measurement, no kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_config5_host_logic(monkeypatch):
    import bench
    from test_serving import FakeModel

    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    n, frames = 24, 52
    out = bench.config5(FakeModel({}), torch.device("cpu"), 0, 1, n, frames, False)
    assert out["warm_up"].startswith("throwaway server")
    assert out["total_frames"] == sum(frames - (r * 7) % 41 for r in range(n))
    for mode in ("per_gpu_batch", "micro_batch_32", "lane_groups_4x32"):
        assert "error" not in out[mode], out[mode]
        assert out[mode]["frames_per_s"] > 0 and out[mode]["decode_calls"] > 0
