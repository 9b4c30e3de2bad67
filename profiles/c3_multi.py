"""BASELINE config C3 on N GPUs of one box: 19x19, 512 concurrent games per GPU (4096 on 8), 800 readouts, tower_height 19,
game-sharded, NCCL replay all-gather every step.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/c3_multi.py [rounds]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import pkg  # noqa: E402

agz = pkg.load()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = agz.GoEnv(19, device=local)
nn = agz.NeuralNet(env, tower_height=19, seed=0)
eng = agz.Engine(19, n_games=512, readouts=800, tower_height=19, seed=0, device=local, world_size=world, rank=rank, evaluator=agz.EVAL_NN_TC)
nn.push(eng)
if world > 1:
    ids = [eng.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.nccl_init(ids[0])
eng.selfplay_start(-1)
eng.selfplay_step(3)
eng.replay_gather()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


barrier()
t0 = time.perf_counter()
pr0 = eng.selfplay_step(1)
pr = eng.selfplay_step(rounds)
eng.replay_gather()
barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda", dtype=torch.float64)
r = torch.tensor([float(pr.readouts - pr0.readouts + 0)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(r, op=dist.ReduceOp.SUM)
if rank == 0:
    ms_round = 1e3 * t.item() / (rounds + 1)
    print(json.dumps({"config": "C3: 19x19, %d games (512 per GPU), 800 readouts, T=19, %d x B200" % (512 * world, world), "n_gpus": world,
                      "ms_per_round": ms_round, "moves_per_s_at_100_rounds_per_move": 512 * world / (ms_round * 100 / 1e3),
                      "positions_per_s": 512 * 8 * world / (ms_round / 1e3), "error": pr.error}), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
