"""World-size-2 gloo tests of the multi-GPU plumbing (host logic only, no compute)."""

import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def test_shard_bounds_cover_the_batch_exactly():
    from cheetah_b200.sharding import shard_bounds

    for n in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank: int, world: int, port: int, results) -> None:
    sys.path.insert(0, str(REPO))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cheetah_b200 as cb
    import workloads
    from cheetah_b200 import lowering, sharding

    n_settings = 10
    # every rank draws the FULL batch from the same seed and keeps its own contiguous slice
    begin, end = sharding.shard_bounds(n_settings, rank, world)
    description = workloads.ares_config3(n_settings, torch.float32, begin, end)
    segment = workloads.product_segment(description, "cpu", torch.float32)

    # rank 0's beam is broadcast once at set-up: afterwards all ranks hold identical particles
    torch.manual_seed(100 + rank)
    beam = cb.ParticleBeam.from_parameters(num_particles=256, dtype=torch.float32)
    nbytes = sharding.broadcast_module(beam)
    gathered = [torch.zeros_like(beam.particles) for _ in range(world)]
    dist.all_gather(gathered, beam.particles)

    program = lowering.lower(list(segment.elements), torch.device("cpu"))
    section = program.stages[0]
    k1 = segment.AREAMQZM1.k1.clone()
    all_k1 = [torch.zeros(end - begin) for _ in range(world)] if (n_settings % world == 0) else None
    if all_k1 is not None:
        dist.all_gather(all_k1, k1)
    results[rank] = {
        "bounds": (begin, end),
        "lattice_shape": section.lattice_shape,
        "n_stages": len(program.stages),
        "beam_equal": all(torch.equal(g, gathered[0]) for g in gathered),
        "broadcast_bytes": nbytes,
        "k1": torch.cat(all_k1) if all_k1 is not None else k1,
    }
    dist.destroy_process_group()


def test_two_ranks_shard_settings_and_share_the_beam():
    import workloads

    world = 2
    manager = mp.Manager()
    results = manager.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert results[0]["bounds"] == (0, 5) and results[1]["bounds"] == (5, 10)
    for rank in range(world):
        assert results[rank]["lattice_shape"] == (5,)
        assert results[rank]["n_stages"] == 1
        assert results[rank]["beam_equal"]
        assert results[rank]["broadcast_bytes"] > 256 * 7 * 4
    # the shards concatenate to the single-process batch: no setting lost or duplicated
    full = workloads.ares_config3(10, torch.float32)
    k1_full = next(e for e in full if e["name"] == "AREAMQZM1")["k1"]
    assert torch.equal(results[0]["k1"], k1_full)


def _gather_worker(rank: int, world: int, port: int, results) -> None:
    sys.path.insert(0, str(REPO))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cheetah_b200 import sharding
    from cheetah_b200.tracking import BeamMoments

    n_settings = 5  # 3 + 2: unequal shards
    begin, end = sharding.shard_bounds(n_settings, rank, world)
    index = torch.arange(begin, end, dtype=torch.float64)
    mine = BeamMoments(
        mu=index[:, None] + torch.arange(6, dtype=torch.float64) * 0.1,
        sigma=(index[:, None] + 1.0).expand(-1, 6).contiguous(),
        num_particles_survived=index * 100.0, energy=torch.tensor(1e8, dtype=torch.float64),
        s=index * 0.5, cov=index[:, None, None] * torch.eye(6, dtype=torch.float64),
    )
    whole = sharding.gather_moments(mine, n_settings)
    results[rank] = {"mu": whole.mu, "sigma": whole.sigma, "survived": whole.num_particles_survived,
                     "s": whole.s, "cov": whole.cov, "energy": whole.energy}
    dist.destroy_process_group()


def test_two_ranks_gather_reduced_observables():
    """The only data-path collective the design allows: a final gather of per-setting moments
    (never the particles), with unequal shard sizes."""
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(world, port, results), nprocs=world, join=True)
    index = torch.arange(5, dtype=torch.float64)
    for rank in range(world):
        got = results[rank]
        assert torch.equal(got["mu"], index[:, None] + torch.arange(6, dtype=torch.float64) * 0.1)
        assert torch.equal(got["sigma"], (index[:, None] + 1.0).expand(-1, 6))
        assert torch.equal(got["survived"], index * 100.0)
        assert torch.equal(got["s"], index * 0.5)
        assert torch.equal(got["cov"], index[:, None, None] * torch.eye(6, dtype=torch.float64))
        assert got["energy"].dim() == 0


def test_shard_segment_respects_the_inner_dimensions_of_a_field():
    """A (2,) ``pixel_size`` / ``misalignment`` is not two settings and a (7, 7) transfer map not
    seven; per-setting parameters registered as nn.Parameter are sliced too."""
    import torch

    import cheetah_b200 as cb
    from cheetah_b200 import sharding

    quad = cb.Quadrupole(length=torch.tensor(0.2), k1=torch.tensor([1.0, 2.0]),
                         misalignment=torch.tensor([1e-3, -2e-3]))
    screen = cb.Screen(pixel_size=torch.tensor([1e-5, 2e-5]), resolution=(10, 10))
    sharding.shard_segment(cb.Segment([quad, screen]), 2, rank=1, world_size=2)
    assert quad.k1.shape == (1,) and float(quad.k1) == 2.0
    assert quad.misalignment.shape == (2,) and screen.pixel_size.shape == (2,)

    tm = torch.eye(7).repeat(7, 1, 1)
    custom = cb.CustomTransferMap(predefined_transfer_map=tm.clone(), length=torch.ones(7))
    one = cb.CustomTransferMap(predefined_transfer_map=torch.eye(7))
    sharding.shard_segment(cb.Segment([custom, one]), 7, rank=0, world_size=7)
    assert custom.predefined_transfer_map.shape == (1, 7, 7) and custom.length.shape == (1,)
    assert one.predefined_transfer_map.shape == (7, 7)

    drift = cb.Drift(length=torch.tensor([0.5, 0.6, 0.7, 0.8]))
    drift._parameters["length"] = torch.nn.Parameter(drift._buffers.pop("length"),
                                                     requires_grad=False)
    sharding.shard_segment(cb.Segment([drift]), 4, rank=1, world_size=2)
    assert drift.length.shape == (2,) and float(drift.length[0]) == pytest.approx(0.7)
