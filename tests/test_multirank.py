"""World-size-2 host logic on CPU (gloo): games shard by global slot (rank + world*slot), every rank plays only its
own games, and the union equals the single-rank result and the oracle -- i.e. results are independent of how games
are sharded over GPUs.  The engines run on the host emulator build of the device code; the
record exchange that NCCL does on the GPU box (agz_replay_gather) is done here with a gloo all_gather_object."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, total_games, readouts, seed, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import pkg
    from emu.build_emu import build
    agz = pkg.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = agz.Engine(9, lib_path=build(), n_games=2, readouts=readouts, seed=seed, world_size=world, rank=rank)
    recs = eng.selfplay_run(total_games)
    mine = [(int(r.game_id), r.moves.tolist(), r.visits.copy(), int(r.result)) for r in recs]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out_q.put([g for part in gathered for g in part])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_single_rank_and_oracle():
    import torch.multiprocessing as mp
    from oracle import go as ogo, selfplay as osp
    total, readouts, seed = 5, 16, 21
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, readouts, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    games = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    games.sort(key=lambda g: g[0])
    assert [g[0] for g in games] == list(range(total))

    class Dummy:
        fake_priors = (np.ones(82) / 82).astype(np.float32)
        fake_value = np.float32(0)

        def __call__(self, positions):
            n = len(positions)
            return np.repeat(self.fake_priors[:, None], n, axis=1), np.repeat(self.fake_value, n)

    oenv = ogo.GoEnv(9)
    for gid, moves, visits, result in games:
        op = osp.selfplay(oenv, Dummy(), readouts, seed=seed, game_id=gid)
        assert moves == [ogo.to_flat(m.move, oenv) for m in op.root.position.recent], gid
        assert np.array_equal(np.array(op.searches_N), visits) and result == op.result
