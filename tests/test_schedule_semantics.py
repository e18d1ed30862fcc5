"""The dependency semantics libvrt's schedule rests on (every (cell, sweep) visit reads FINAL / THIS / LAG / ZERO
operands, any topological order of the visits is valid), checked without a GPU: profiles/microbench/tile_schedule_spec.py
builds the visits from the oracle's layers and stencil, executes them in a cache-friendly NON-reference order with one
slot per visit, and must reproduce the sequential oracle bit for bit."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("theta,phi", [(152.7, 315.5), (67.2, 155.8)])
def test_reordered_visits_reproduce_the_sequential_sweep(oracle, theta, phi):
    sys.path.insert(0, os.path.join(ROOT, "profiles", "microbench"))
    import tile_schedule_spec as spec
    assert spec.run(theta, phi)
    assert spec.run(theta, phi, cells_per_tile=500)


@pytest.mark.parametrize("theta,phi", [(152.7, 315.5), (27.3, 135.5)])
def test_blocked_order_of_schedule_rule_5(oracle, theta, phi):
    """the key of schedule.cu rule 5 (column blocks in upwind order x level slabs, pushed behind the producers) in numpy: a
    topological order whose equal keys are independent, executing to the sequential oracle's bits"""
    sys.path.insert(0, os.path.join(ROOT, "profiles", "microbench"))
    import tile_schedule_spec as spec
    assert spec.run(theta, phi, blocks=(3, 2), slab=5)
    assert spec.run(theta, phi, blocks=(4, 4), slab=0)
