#!/usr/bin/env python
"""bench.py -- self-play moves/sec of the B200 engine (BASELINE.json metric) and of the CPU reference arm.

  python bench.py --gpus N --steps K --warmup W            our arm (under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (oracle port), rank 0 only

Headline workload (config C2 of BASELINE.json): 9x9 Go, 1024 concurrent self-play games per GPU, 400 readouts per move,
tower_height 6, random-init weights (Flux default init restated, seed 0), all games from empty boards, finished games
refilled.  One "step" = 50 batched tree_search! rounds = 400 readouts for every live game = one move-step of the whole job;
moves are counted from the device counter, so `value` = moves actually played / device time.

Steady state (src/train.jl:56-61): the slots are started staggered (option selfplay.stagger_rounds: slot g begins
g*R/n_games rounds late, R = --burnin move-steps) and --burnin untimed move-steps are played first, so that inside the timed
region the games sit at every ply, some finish and are refilled in every step, the harvest / replay-pack / NCCL all-gather
carry real payloads, and arena compaction runs under load.

Extra legs (same engine and kernels, each a few seconds, reported as extra keys of the one JSON line):
  roofline_tree  C5: MCTS-only 9x9, uniform prior (DummyNet), 8192 trees x 1600 readouts -> tree-kernel HBM roofline
  c4             C4: network only, 19x19, batch 8192, tower_height 19 -> conv tensor roofline at the north-star shape
  c3             C3 per-GPU share: 19x19, 512 games per GPU, 800 readouts, tower_height 19 (+ NCCL replay gather when N > 1)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BOARD, GAMES, READOUTS, TOWER, ROUNDS_PER_STEP = 9, 1024, 400, 6, 50
METRIC = "self-play moves/sec (9x9, 400 readouts)"
# DRAM bytes per tower-conv launch (8192 positions) from the committed ncu --set full capture of this workload
NCU_CONV_DRAM_BYTES_PER_LAUNCH = 0.5 * ((341.14 + 299.02) + (680.71 + 314.34)) * 1e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


def oracle_moves_per_sec(max_moves, budget_s):
    """The reference's CPU path (oracle port): one game at a time, <= 8 leaves per network call, fp32 torch-CPU net on ALL host
    cores (torchrun exports OMP_NUM_THREADS=1, so the thread count is set explicitly: the arm is the same at every N)."""
    import torch
    from oracle import go as ogo, net as onet, selfplay as osp
    torch.set_num_threads(os.cpu_count() or 1)
    env = ogo.GoEnv(BOARD)
    nn = onet.NeuralNet(BOARD, TOWER, seed=0)
    state = {"moves": 0, "t0": time.perf_counter(), "elapsed": 0.0}

    class Stop(Exception):
        pass

    def on_move(player, move):
        state["moves"] += 1
        state["elapsed"] = time.perf_counter() - state["t0"]
        if state["moves"] >= max_moves or state["elapsed"] > budget_s:
            raise Stop()

    gid = 0
    try:
        while True:
            osp.selfplay(env, nn, READOUTS, seed=0, game_id=gid, on_move=on_move)
            gid += 1
    except Stop:
        pass
    return state["moves"] / state["elapsed"], state["moves"], state["elapsed"], torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_moves = 2
    t0 = time.perf_counter()
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_moves_per_sec(1, 30)
    vals, cores = [], os.cpu_count() or 1
    for _ in range(args.steps):
        v, m, el, cores = oracle_moves_per_sec(per_step_moves, 60)
        vals.append((m, el))
        if time.perf_counter() - t0 > 240:
            break
    moves = sum(m for m, _ in vals)
    el = sum(e for _, e in vals)
    value = moves / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "moves/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1e3 * el / max(1, len(vals)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C2 sample: 9x9 Go, 1 game at a time, 400 readouts/move, tower_height 6, random-init weights, %d moves per step" % per_step_moves},
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": cores, "kind": "port",
                         "sample": "%d moves of sequential self-play (oracle port of src/selfplay.jl: python tree, torch-CPU fp32 net, <=8 leaves per call)" % moves},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ extra legs
def tree_bytes_per_readout(N, d):
    """(d-1)*(3*4*A + A) select reads + 4*4*A expand writes + 8*N^2 child boards + 16*d path read-modify-writes."""
    A = N * N + 1
    return (d - 1) * 13 * A + 16 * A + 8 * N * N + 16 * d


def leg_c5(agz, device, hbm, lib_path=None):
    """BASELINE config C5: MCTS-only, 9x9, uniform prior / value 0 (DummyNet), 8192 trees x 1600 readouts per move."""
    trees, readouts, rounds = 8192, 1600, 200
    eng = agz.Engine(9, n_games=trees, readouts=readouts, tower_height=1, seed=0, evaluator=agz.EVAL_DUMMY, nodes_per_game=3600, device=device,
                     **({"lib_path": lib_path} if lib_path else {}))
    try:
        eng.selfplay_start(-1)
        pr0 = eng.selfplay_step(rounds + 10)              # warm-up: past the first move of every tree
        ms, pr1 = 0.0, pr0
        reps = 3
        for _ in range(reps):                             # timed: device time (CUDA events on the engine's stream)
            pr1 = eng.selfplay_step(rounds)
            ms += pr1.step_ms
        ro, pn, mv = pr1.readouts - pr0.readouts, pr1.path_nodes - pr0.path_nodes, pr1.moves_played - pr0.moves_played
        d = pn / max(1, ro)
        bpr = tree_bytes_per_readout(9, d)
        achieved = ro * bpr / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "k_warps<DummyRoundsOp<3,1>> (select_leaf + expand + virtual loss + incorporate + backup + move logic, one warp per tree, %d rounds per launch)" % rounds,
                "workload": "C5: MCTS-only 9x9, uniform prior (DummyNet), %d trees x %d readouts/move; %d x %d rounds timed after %d warm-up rounds" % (trees, readouts, reps, rounds, rounds + 10),
                "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                "bytes_per_readout": bpr, "mean_path_nodes": d, "readouts_per_launch": ro / reps, "ms_per_launch": ms / reps, "ms_per_round": ms / (reps * rounds),
                "moves_per_s": mv / (ms * 1e-3), "readouts_per_s": ro / (ms * 1e-3), "error": int(pr1.error)}
    finally:
        eng.close()


def leg_19(agz, device, tf_sus, games, rounds, world, rank, dist, name):
    """19x19, tower_height 19, 800 readouts: `games` concurrent games on this GPU (8 leaves each per round).  C4 = the network on a
    batch of 8192 positions harvested from this self-play; C3 = one GPU's share of the 4096-game config."""
    env = agz.GoEnv(19, device=device)
    nn = agz.NeuralNet(env, tower_height=19, seed=0)
    eng = agz.Engine(19, n_games=games, readouts=800, tower_height=19, seed=0, device=device, world_size=world, rank=rank, evaluator=agz.EVAL_NN_TC,
                     nodes_per_game=4000)
    try:
        nn.push(eng)
        if dist is not None and name == "c3":
            ids = [eng.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            eng.nccl_init(ids[0])
        eng.selfplay_start(-1)
        # >= 4 s of back-to-back rounds first: the roofline denominator is the SUSTAINED tensor peak (MEASURED_PEAKS.json: 4 s of
        # back-to-back GEMMs); a leg timed from a cool start would run at burst clocks and "beat" it
        pr = eng.selfplay_step(3)
        pr = eng.selfplay_step(max(1, int(4000.0 / max(1.0, pr.step_ms / 3))))
        pr = eng.selfplay_step(rounds)                   # plain rounds: device time of the whole round
        ms_round = pr.step_ms / rounds
        if name == "c3":
            eng.replay_gather()
        eng.set_timing(True)
        eng.phase_times(reset=True)
        eng.selfplay_step(4)
        kms, kln = eng.phase_times(reset=True)
        eng.set_timing(False)
        fpos, fconv = eng.net_flops()
        rows = games * 8
        conv_ms = kms[3] / max(1, kln[3])
        net_ms = (kms[2] + kms[3] + kms[4]) / max(1, kln[0])
        out = {"workload": "19x19, %d games x 8 leaves = %d positions per batch, tower_height 19, 256 filters, positions from live self-play; timed after >= 4 s of back-to-back rounds (sustained clocks)" % (games, rows),
               "ms_per_round": ms_round, "network_ms_per_batch": net_ms, "positions_per_s": rows / (net_ms * 1e-3),
               "network_tflops": fpos * rows / (net_ms * 1e-3) / 1e12,
               "roofline": {"bound": "tensor", "kernel": "conv3x3_tc5_kernel / conv3x3_tc6_kernel", "achieved": fconv * rows / (conv_ms * 1e-3) / 1e12, "peak": tf_sus,
                            "unit": "TFLOP/s", "frac": fconv * rows / (conv_ms * 1e-3) / 1e12 / tf_sus, "flops_per_launch": fconv * rows, "ms_per_launch": conv_ms, "traffic": None},
               "kernel_ms_per_round": {n: kms[i] / max(1, kln[0]) for i, n in enumerate(agz.binding.KERNEL_NAMES)}, "error": int(pr.error)}
        if name == "c3":
            out["moves_per_s_per_gpu_at_100_rounds_per_move"] = games / (ms_round * 100 / 1e3)
        return out
    finally:
        eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="agz")
    ap.add_argument("--games", type=int, default=GAMES)
    ap.add_argument("--burnin", type=int, default=112, help="untimed move-steps before the warm-up (staggered start; ~ one mean game length)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the C5 / C4 / C3 legs")
    ap.add_argument("--pipeline", type=int, default=0, help="option schedule.pipeline (two half batches on separate streams)")
    ap.add_argument("--precision", type=int, default=1, help="option conv.precision: 1 = fp16 operands (default), 2 = split precision (hi + lo fp16 pairs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import pkg
    agz = pkg.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(3, args.warmup)

    env = agz.GoEnv(BOARD, device=local)
    nn = agz.NeuralNet(env, tower_height=TOWER, seed=0)
    eng = agz.Engine(BOARD, n_games=args.games, readouts=READOUTS, tower_height=TOWER, seed=0, device=local, world_size=world, rank=rank,
                     evaluator=agz.EVAL_NN_TC, options={"selfplay.stagger_rounds": args.burnin * ROUNDS_PER_STEP, "schedule.pipeline": args.pipeline, "conv.precision": args.precision})
    nn.push(eng)
    if world > 1:  # NCCL communicator of the replay all-gather: rank 0's unique id goes round through torch.distributed
        ids = [eng.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.nccl_init(ids[0])
    flat_params = [np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in nn.params[k]]) for k in range(3)]
    param_bytes = int(sum(p.nbytes for p in flat_params))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        """one step, inputs resident: 50 rounds + replay all-gather of the games that finished"""
        pr = eng.selfplay_step(ROUNDS_PER_STEP)
        eng.replay_gather()
        eng.selfplay_harvest_discard()
        return pr

    eng.selfplay_start(-1)
    t_burn = time.perf_counter()
    for _ in range(args.burnin):          # untimed: reach the steady state (every slot live, games spread over all plies)
        pr = device_step()
    t_burn = time.perf_counter() - t_burn
    for _ in range(warmup):
        pr = device_step()
    slots_live = pr.games_live
    # ---- timed region 1: device-resident throughput (`value`) ---------------------------------------------------
    l0 = eng.kernel_launches()
    stats0 = eng.selfplay_stats()
    g0 = eng.replay_info()["gathered_bytes"]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    m0, p0, r0, f0 = pr.moves_played, pr.positions_evaluated, pr.readouts, pr.games_finished
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        pr = device_step()
        dev_ms += pr.step_ms
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    launches = eng.kernel_launches() - l0
    gathered = eng.replay_info()["gathered_bytes"] - g0
    stats1 = eng.selfplay_stats()
    moves = pr.moves_played - m0
    fill = (pr.positions_evaluated - p0) / float(args.steps * ROUNDS_PER_STEP * args.games * 8)
    readouts_done, finished = pr.readouts - r0, pr.games_finished - f0
    # device time of the steps (CUDA events on the engine stream) plus the gather/harvest tail measured by wall clock
    t_rank = max(wall, dev_ms / 1e3)
    # ---- per-kernel CUDA-event timing on the same workload, sequential schedule ---------------------------------
    eng.set_timing(True)
    eng.phase_times(reset=True)
    for _ in range(min(2, args.steps)):
        device_step()
    kms, kln = eng.phase_times(reset=True)
    eng.set_timing(False)
    pr = eng.selfplay_step(1)
    # ---- timed region 2: end to end through the public API with host buffers (`e2e`) ----------------------------
    barrier()
    m1 = pr.moves_played
    d2h = 0
    glen = []
    L, A = eng.L, eng.A
    t1 = time.perf_counter()
    for _ in range(args.steps):
        for k in range(3):                      # H2D: the caller's current network parameters (train.jl hands selfplay cur_nn)
            eng.net_set_params(k, flat_params[k])
        pr = eng.selfplay_step(ROUNDS_PER_STEP)
        eng.replay_gather()
        recs = eng.selfplay_harvest(4 * args.games)   # D2H: finished games (headers, moves, q, pi, visits)
        glen += [r.n_moves for r in recs]
        d2h += len(recs) * (40 + L * (2 + 4 + 8 * A)) + 80   # what agz_selfplay_harvest transfers: fixed-stride records + counters
    barrier()
    t_e2e = time.perf_counter() - t1
    moves_e2e = pr.moves_played - m1
    info = eng.info()
    _, conv_flops_pos = eng.net_flops()
    flops_pos, _ = eng.net_flops()
    err = pr.error
    eng.close()

    if err:
        print(json.dumps({"metric": METRIC, "error": "a game stopped on the device with status %d (6 = node arena full)" % err}), flush=True)
        sys.exit(2)
    hbm, tf_sus, tf_burst, src = peaks()
    legs = {}
    if not args.no_legs:
        legs["roofline_tree"] = leg_c5(agz, local, hbm)
        legs["c4"] = leg_19(agz, local, tf_sus, 1024, 5, 1, 0, None, "c4")
        legs["c3"] = leg_19(agz, local, tf_sus, 512, 8, world, rank, dist, "c3")
    tot = torch.tensor([float(moves), float(moves_e2e), float(launches), float(gathered), float(finished), float(len(glen)), float(d2h)], device="cuda", dtype=torch.float64)
    tmx = torch.tensor([t_rank, t_e2e, legs["c3"]["ms_per_round"] if legs else 0.0], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
    tot, tmx = tot.cpu().numpy(), tmx.cpu().numpy()
    if rank == 0:
        rows = args.games * 8
        conv_ms = kms[3] / max(1, kln[3])
        achieved = conv_flops_pos * rows / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": tot[0] / tmx[0], "unit": "moves/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": 1e3 * tmx[0] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == 1 else "f16x2-split",
            "data": "synthetic",
            "config": {"workload": "C2: 9x9 Go, %d concurrent self-play games per GPU, 400 readouts/move (50 rounds x 8 leaves per step), tower_height 6, 256 filters, random-init weights seed 0, empty-board starts, finished games refilled" % args.games,
                       "step": "50 tree_search rounds over all games (select -> leaf features -> stem + 12 tower convs -> heads -> incorporate/move logic, one stream) + replay pack / all-gather of the games that finished",
                       "steady_state": "staggered start + %d untimed burn-in move-steps (%.0f s) before the %d warm-up steps: %d of %d slots live, games at every ply" % (args.burnin, t_burn, warmup, slots_live, args.games),
                       "l2": "inputs larger than L2: tree arenas %.1f GB (%d nodes per game) and %.0f MB per activation buffer vs 126 MB L2" % (info["n_games"] * info["nodes_per_game"] * info["bytes_per_node"] / 1e9, info["nodes_per_game"], rows * 81 * 512 / 1e6)},
            "e2e": {"value": tot[1] / tmx[1], "unit": "moves/s", "h2d_bytes_per_step": param_bytes, "d2h_bytes_per_step": int(tot[6] / world / args.steps)},
            "gpu_launches": int(tot[2]),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "conv3x3_tc5_kernel / conv3x3_tc6_kernel (tower 3x3 conv 256->256, fp16 tcgen05 cta_group::2 + TMA im2col; tc6 = second conv of a block, shortcut tile by TMA)",
                         "measured_in": "CUDA events around every kernel, sequential schedule, %d steps of the same workload right after the timed region" % min(2, args.steps), "achieved": achieved, "peak": tf_sus, "unit": "TFLOP/s",
                         "frac": achieved / tf_sus, "traffic": NCU_CONV_DRAM_BYTES_PER_LAUNCH if args.games == GAMES else None,
                         "traffic_detail": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full on this workload (profiles/r02_conv_raw.csv): tc5 341.1 + 299.0 MB, tc6 680.7 + 314.3 MB, mean of the two; algorithmic 680 / 1020 MB",
                         "peak_source": src + " bf16 sustained",
                         "flops_per_launch": conv_flops_pos * rows, "ms_per_launch": conv_ms},
            "kernel_ms_per_round": {n: kms[i] / max(1, kln[0]) for i, n in enumerate(agz.binding.KERNEL_NAMES)},
            "games_finished_timed": int(tot[4]), "harvested_games_e2e": int(tot[5]), "mean_game_length_e2e": (sum(glen) / len(glen)) if glen else None,
            "replay_gathered_bytes_per_step": tot[3] / world / args.steps,   # tuple bytes every rank appended to its ring per step (all ranks' payload)
            "leaf_fill": fill, "readouts_per_s": readouts_done / tmx[0],
            "duplicate_leaf_frac": (stats1["duplicate_leaves"] - stats0["duplicate_leaves"]) / max(1, stats1["positions_evaluated"] - stats0["positions_evaluated"]),
            "arena_prunes": stats1["arena_prunes"],
            "network_tflops": flops_pos * rows / ((kms[2] + kms[3] + kms[4]) / max(1, kln[0]) * 1e-3) / 1e12 if kln[0] else None,
        }
        if legs:
            c3 = legs["c3"]
            c3["n_gpus"] = world
            c3["ms_per_round_max_over_ranks"] = tmx[2]
            c3["moves_per_s_all_gpus_at_100_rounds_per_move"] = 512 * world / (tmx[2] * 100 / 1e3)
            line.update(legs)
        if world == 1 and not args.no_cpu_baseline:
            v, m, el, cores = oracle_moves_per_sec(8, 25)
            line["cpu_baseline"] = {"value": v, "unit": "moves/s", "cores": cores, "kind": "port",
                                    "sample": "first %d moves of one 9x9 game, 400 readouts, tower_height 6 (oracle port of src/selfplay.jl, %.1f s)" % (m, el)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
