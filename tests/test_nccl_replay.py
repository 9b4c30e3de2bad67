"""The one collective of the path on real GPUs: agz_replay_gather over NCCL.  Two ranks, one B200 each, play disjoint
games (global slot = rank + world * slot); after the gather BOTH replay rings must hold the tuples of ALL games -- ragged counts,
padded ncclAllGather -- and equal what the oracle's extract_data produces.  Skipped with fewer than two GPUs."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run(total_games=6, readouts=16, seed=5):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker_lockstep, args=(r, 2, total_games, readouts, seed, uid_q, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = [out_q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return msgs


def _worker_lockstep(rank, world, total_games, readouts, seed, uid_q, out_q):
    """Fixed number of (step, gather) iterations on both ranks, so the collectives pair up."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import pkg
    agz = pkg.load()
    eng = agz.Engine(9, n_games=2, readouts=readouts, seed=seed, device=rank, world_size=world, rank=rank)
    eng.set_dummy_evaluator(None, 0.0)
    if rank == 0:
        uid = eng.nccl_unique_id()
        uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    eng.nccl_init(uid)
    eng.selfplay_start(total_games)
    recs, total = [], 0
    for _ in range(60):                       # 60 x 64 rounds: far more than 3 games of <= 113 moves x 2 rounds need
        pr = eng.selfplay_step(64)
        assert pr.error == 0
        total = eng.replay_gather()
        recs += eng.selfplay_harvest(16)
    boards, tp, pis, zs = eng.replay_read(0, total)
    out_q.put({"rank": rank, "total": int(total), "boards": boards, "tp": tp, "pis": pis, "zs": zs,
               "recs": [(int(r.game_id), r.moves.tolist(), r.searches_pi.copy(), int(r.result)) for r in recs]})
    eng.close()


@pytest.mark.gpu
def test_nccl_replay_all_gather_two_gpus():
    from oracle import go as ogo
    msgs = _run()
    by_rank = {m["rank"]: m for m in msgs}
    games = sorted(by_rank[0]["recs"] + by_rank[1]["recs"])
    assert [g[0] for g in games] == list(range(6))
    assert {g[0] % 2 for g in by_rank[0]["recs"]} == {0} and {g[0] % 2 for g in by_rank[1]["recs"]} == {1}
    n_all = sum(len(g[1]) for g in games)
    oenv = ogo.GoEnv(9)
    expected = []
    for gid, moves, pis, result in games:
        pos = ogo.GoPosition(oenv)
        for t, m in enumerate(moves):
            expected.append((pos.board.flatten(order="F").tobytes(), int(pos.to_play), pis[t].tobytes(), result))
            pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))
    want = sorted(expected)
    for r in (0, 1):                                      # every rank's ring holds every rank's tuples
        m = by_rank[r]
        assert m["total"] == n_all
        got = sorted((m["boards"][k].tobytes(), int(m["tp"][k]), m["pis"][k].tobytes(), int(m["zs"][k])) for k in range(n_all))
        assert got == want, "rank %d ring differs from the oracle's tuples" % r


def _worker_ragged(rank, world, uid_q, out_q):
    """ONE gather after all games are over: rank 0 holds two finished games, rank 1 one -- the largest block is then bigger than
    the smaller rank's own payload (round 1 re-allocated the send buffer after packing in that case and sent garbage)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import pkg
    agz = pkg.load()
    eng = agz.Engine(9, n_games=4, readouts=16, seed=9, device=rank, world_size=world, rank=rank)
    eng.set_dummy_evaluator(None, 0.0)
    if rank == 0:
        uid = eng.nccl_unique_id()
        uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    eng.nccl_init(uid)
    eng.selfplay_start(3)                     # game ids 0, 2 on rank 0; 1 on rank 1
    mine = 2 if rank == 0 else 1
    for _ in range(400):
        pr = eng.selfplay_step(64)
        assert pr.error == 0
        if pr.games_finished == mine:
            break
    assert pr.games_finished == mine
    total = eng.replay_gather()               # the only collective call: both ranks make exactly one
    recs = eng.selfplay_harvest(16)
    boards, tp, pis, zs = eng.replay_read(0, total)
    info = eng.replay_info()
    out_q.put({"rank": rank, "total": int(total), "boards": boards, "tp": tp, "pis": pis, "zs": zs, "last_gather_bytes": info["last_gather_bytes"],
               "tuple_bytes": info["tuple_bytes"], "recs": [(int(r.game_id), r.moves.tolist(), r.searches_pi.copy(), int(r.result)) for r in recs]})
    eng.close()


@pytest.mark.gpu
def test_nccl_replay_ragged_single_gather_two_gpus():
    import torch
    import torch.multiprocessing as mp
    from oracle import go as ogo
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker_ragged, args=(r, 2, uid_q, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = {m["rank"]: m for m in (out_q.get(timeout=600) for _ in range(2))}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    games = sorted(msgs[0]["recs"] + msgs[1]["recs"])
    assert [g[0] for g in games] == [0, 1, 2] and len(msgs[0]["recs"]) == 2 and len(msgs[1]["recs"]) == 1
    oenv = ogo.GoEnv(9)
    expected = []
    for gid, moves, pis, result in games:
        pos = ogo.GoPosition(oenv)
        for t, m in enumerate(moves):
            expected.append((pos.board.flatten(order="F").tobytes(), int(pos.to_play), pis[t].tobytes(), result))
            pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))
    want = sorted(expected)
    for r in (0, 1):
        m = msgs[r]
        assert m["total"] == len(want) and m["last_gather_bytes"] == len(want) * m["tuple_bytes"]
        got = sorted((m["boards"][k].tobytes(), int(m["tp"][k]), m["pis"][k].tobytes(), int(m["zs"][k])) for k in range(m["total"]))
        assert got == want, "rank %d ring differs from the oracle's tuples" % r


def _train_worker(rank, world, uid_q, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import pkg
    agz = pkg.load()
    from oracle import net as onet, train as otrain
    import test_abi_train as tt
    N, T, B = 9, 1, 6
    onn = onet.NeuralNet(N, T, seed=3)
    onn.randomize_bn(seed=4)
    eng = agz.Engine(N, n_games=8, readouts=8, tower_height=T, evaluator=agz.EVAL_NN_TC, device=rank, world_size=world, rank=rank)
    flat = otrain.flat_params(onn)
    bns = [onn.base_bns(), [onn.v_bn], [onn.p_bn]]
    for k in range(3):
        eng.net_set_params(k, flat[k])
        eng.net_set_bn_stats(k, np.concatenate([b.mu for b in bns[k]]), np.concatenate([b.sigma for b in bns[k]]), agz.BN_VAR_EPS)
    if rank == 0:
        uid = eng.nccl_unique_id()
        uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    eng.nccl_init(uid)
    losses = []
    for step in range(2):
        positions, pis, zs = tt._batch(N, B, 100 * step + rank)      # every rank trains on its own minibatch
        bh, tp = tt._hist(positions, N)
        losses.append(eng.train_step(bh, tp, pis, zs, lr=0.02, momentum=0.9))
    out_q.put({"rank": rank, "losses": losses, "grads": [eng.train_read_grads(k) for k in range(3)], "params": [eng.net_get_params(k) for k in range(3)],
               "bn": [eng.net_get_bn_stats(k)[:2] for k in range(3)]})
    eng.close()


@pytest.mark.gpu
def test_data_parallel_train_step_two_gpus():
    """agz_train_step with an NCCL communicator: gradients, loss and moved running statistics are averaged over the ranks, both
    ranks end with identical parameters, and they equal the oracle's step on the averaged gradients of the two minibatches."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import net as onet, train as otrain
    import test_abi_train as tt
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, uid_q, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = {m["rank"]: m for m in (out_q.get(timeout=600) for _ in range(2))}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    N, T, B = 9, 1, 6
    onn = onet.NeuralNet(N, T, seed=3)
    onn.randomize_bn(seed=4)
    tr = otrain.Trainer(onn)
    want_losses = [tr.step_data_parallel([tt._batch(N, B, 100 * step + r) for r in range(2)]) for step in range(2)]
    want = otrain.flat_params(onn)
    bns = [onn.base_bns(), [onn.v_bn], [onn.p_bn]]
    for r in (0, 1):
        m = msgs[r]
        assert np.allclose(m["losses"], want_losses, rtol=3e-4)
        for k in range(3):
            assert tt._rel(m["grads"][k], tr.last_grads[k]) < 2e-3
            assert np.max(np.abs(m["params"][k] - want[k])) < 2e-5
            assert np.allclose(m["bn"][k][0], np.concatenate([b.mu for b in bns[k]]), atol=2e-5, rtol=1e-4)
            assert np.allclose(m["bn"][k][1], np.concatenate([b.sigma for b in bns[k]]), atol=2e-5, rtol=1e-4)
    for k in range(3):
        assert np.array_equal(msgs[0]["params"][k], msgs[1]["params"][k]), "ranks diverged"
