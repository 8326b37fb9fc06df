"""Host-side logic that needs no GPU: weight-cache invalidation hooks of the module, NUMA helper parsing, bench helpers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_cache_is_invalidated_by_every_parameter_changing_entry_point():
    from tokenpacker_b200 import TokenPackerB200
    m = TokenPackerB200(hidden_size=64, scale_factor=2)
    sentinel = object()

    def arm():
        m._packed, m._packed_key = sentinel, ("armed",)

    arm()
    m.load_state_dict(m.state_dict())
    assert m._packed is None and m._packed_key is None
    arm()
    m.to(torch.float16)                       # _apply: .to() / .cuda() / .half() / .float()
    assert m._packed is None
    arm()
    m.train()
    assert m._packed is None
    arm()
    m.eval()
    assert m._packed is None
    arm()
    m.invalidate_packed()
    assert m._packed is None


def test_cpu_inputs_still_fail_loudly():
    import pytest
    from tokenpacker_b200 import TokenPackerB200
    m = TokenPackerB200(hidden_size=64, scale_factor=2)
    with pytest.raises(RuntimeError):
        m((torch.zeros(1, 576, 1024), torch.zeros(1, 576, 4096)))


def test_numa_cpulist_parsing_and_graceful_fallback():
    from tokenpacker_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    rep = numa.bind_to_gpu_node(0)            # no GPU / no sysfs topology here: must report, not raise
    assert isinstance(rep, dict) and "bound" in rep


def test_bench_reference_arm_runs_the_configs1_batch():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.N_CROPS == 64 and bench.SCALE == 2 and bench.HIDDEN == 4096
    assert sum(a * b + (1 if a * b > 1 else 0) for a, b in bench.HD5_GRIDS) == 256      # BASELINE configs[4]: 256 crops
