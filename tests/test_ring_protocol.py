"""Queue protocol of the experimental barrier-free wavefront kernel (k_wf_ring, RT_WF_GRAIN=ring): the claim /
push / termination code of csrc/rt_ring.hpp, compiled for the host and driven by threads that play CTAs
(tests/ring_sim.cpp).  Checks what only the protocol can break — lost or duplicated slots, paths that start or
end twice, frames that never terminate, counters that do not balance, lap tags that alias — not the kernel's
arithmetic, which is the arithmetic of the other wavefront kernels."""
import os
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def ring_sim(tmp_path_factory):
    exe = tmp_path_factory.mktemp("ring") / "ring_sim"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-Wno-unknown-pragmas", "-o", str(exe), str(ROOT / "tests" / "ring_sim.cpp")])
    return str(exe)


def _run(exe, *args):
    out = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("ok "), out.stdout
    return dict(kv.split("=") for kv in out.stdout.split()[1:])


# threads, slots, log2(ring capacity), chunk, paths per frame, frames, seed
@pytest.mark.parametrize("cfg", [
    (1, 64, 7, 8, 20000, 10, 1),     # tiny ring, one consumer: the lap tags wrap many times (64 laps per wrap)
    (1, 4, 3, 4, 3000, 20, 2),       # ring of 8 entries, 4 slots
    (1, 1000, 11, 128, 50000, 3, 3), # the kernel's chunk size, slots not a power of two
])
def test_lap_tags_and_bookkeeping_single_consumer(ring_sim, cfg):
    got = _run(ring_sim, *cfg)
    assert int(got["laps"]) > 64 or cfg[1] == 1000


@pytest.mark.parametrize("cfg", [
    (8, 4096, 20, 16, 200000, 3, 7),  # many chunks in flight
    (6, 1024, 20, 128, 200000, 2, 3), # the kernel's chunk size: partial chunks most of the time
    (8, 64, 20, 4, 60000, 4, 11),     # few slots: consumers fight over every entry (credits go negative and come back)
    (3, 16, 20, 1, 30000, 3, 5),      # one entry per claim
    (8, 256, 20, 32, 100, 20, 13),    # fewer paths than slots: frames of a few entries, termination raced every time
])
def test_concurrent_consumers(ring_sim, cfg):
    threads = min(cfg[0], max(2, os.cpu_count() or 2))
    _run(ring_sim, threads, *cfg[1:])


@pytest.mark.parametrize("cfg", [
    (8, 65536, 20, 4, 400000, 2, 7, 4),  # batch claims (WF_RING_BATCH): 4 chunks at once while a class holds >= 64 batches
    (1, 8192, 14, 2, 100000, 3, 3, 4),   # ... single consumer, ring laps
    (6, 16384, 20, 1, 200000, 2, 9, 8),
])
def test_batch_claims(ring_sim, cfg):
    threads = min(cfg[0], max(1, os.cpu_count() or 2))
    got = _run(ring_sim, threads, *cfg[1:])
    assert got["batch"] == str(cfg[7])


def test_rejects_ring_smaller_than_two_laps(ring_sim):
    out = subprocess.run([ring_sim, "1", "64", "6", "8", "100", "1", "1"], capture_output=True, text=True)
    assert out.returncode == 64
