"""The N>1 path on CPU (gloo, world_size 2): the one collective (weight broadcast) and the round-robin batch
sharding of Asyrp.run_test — no data-path collective exists."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import argparse
    from asyrp_official_b200 import modules, synthetic
    from asyrp_official_b200.configs import load_config
    from asyrp_official_b200.diffusion_latent import Asyrp, broadcast_weights
    cfg = load_config("celeba")
    cfg.model.ch, cfg.model.ch_mult, cfg.data.image_size = 64, [1, 2], 32
    m = modules.DDPM(cfg)
    m.setattr_layers(1)
    synthetic.randomize_(m, seed=100 + rank)  # ranks start with DIFFERENT weights
    v0 = m._version
    broadcast_weights(m, src=0)
    ref = modules.DDPM(cfg)
    ref.setattr_layers(1)
    synthetic.randomize_(ref, seed=100)
    same = all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), ref.state_dict().values()))
    # sharding rule of run_test: batch b goes to rank b % world
    r = Asyrp(argparse.Namespace(user_defined_t_edit=500, user_defined_t_addnoise=200), cfg, device="cpu")
    mine = [b for b in range(7) if b % r.world == r.rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # cache files (precomputed/*.pth, diffusion_latent.py:974-982,1082,1167): rank 0 writes through a temp file + rename,
    # the other ranks wait at a barrier and load — nobody reads a half-written file, nobody duplicates the inversion
    from asyrp_official_b200.diffusion_latent import _atomic_save
    path = os.path.join(os.environ["ASYRP_TEST_TMP"], "pairs.pth")
    value = None
    if rank == 0:
        value = [[torch.full((1, 3, 4, 4), float(i)), torch.zeros(1, 3, 4, 4), torch.ones(1, 3, 4, 4) * i] for i in range(3)]
        _atomic_save(value, path)
    got = r._sync_cache(path, value)
    cache_ok = len(got) == 3 and all(torch.equal(t[0], torch.full((1, 3, 4, 4), float(i))) for i, t in enumerate(got)) \
        and not [f for f in os.listdir(os.environ["ASYRP_TEST_TMP"]) if ".tmp." in f]
    q.put((rank, same and cache_ok, m._version > v0, gathered))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2(tmp_path):
    os.environ["ASYRP_TEST_TMP"] = str(tmp_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, same, bumped, gathered in res:
        assert same, f"rank {rank}: weights differ from rank 0 after broadcast, or the rank-0 cache was not received"
        assert bumped
        assert sorted(gathered[0] + gathered[1]) == list(range(7)) and not set(gathered[0]) & set(gathered[1])
