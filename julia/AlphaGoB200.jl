# AlphaGoB200.jl -- Julia front end of libagz (include/agz.h): AlphaGo.jl's self-play surface on a B200.
#
# STATUS: UNVERIFIED HERE.  The build image has no Julia binary, so this file has never been parsed or run; the very same
# entry points are exercised end to end by the ctypes twin of this file (alphago.jl_b200/binding.py + api.py) in tests/.
# Struct layouts are pinned by tests/test_struct_layout.py (offsetof of every field, compiled from include/agz.h).
#
# Surface mirrored (reference paths): GoEnv (src/game/go/go.jl:1-26), NeuralNet (src/neural_net.jl:7-33,57-73),
# MCTSPlayer fields root / searches_π / qs / result / result_string (src/mcts_play.jl:3-24), selfplay (src/selfplay.jl:1-45),
# extract_data (src/mcts_play.jl:126-139), get_replay_batch + the train loop (src/train.jl:4-12,38-92), evaluate
# (src/neural_net.jl:103-158).  Indices cross the ABI 0-based (flat move f0 = N*(j-1) + (i-1), pass = N*N) and are converted
# to the reference's 1-based (i, j) / `nothing` here.
#
# Use inside AlphaGo.jl:   include("AlphaGoB200.jl"); using .AlphaGoB200
#   env = AlphaGoB200.GoEnv(9); nn = AlphaGoB200.NeuralNet(env; tower_height = 6)
#   players = AlphaGoB200.selfplay(env, nn, 400; n_games = 1024)       # Vector of finished players
#   pos, πs, res = AlphaGoB200.extract_data(players[1])
# With Flux present, `NeuralNet(env, flux_nn)` copies the parameters of an AlphaGo.NeuralNet (base_net / value / policy chains).
module AlphaGoB200

export GoEnv, GomokuEnv, Go, GameEnv, NeuralNet, MCTSPlayer, B200Engine, IllegalMove, selfplay, extract_data, get_replay_batch, train, evaluate,
       set_option!, get_option, load_weights!, read_weights

const libagz = get(ENV, "LIBAGZ", joinpath(@__DIR__, "..", "alphago.jl_b200", "libagz.so"))

struct IllegalMove <: Exception end                      # src/AlphaGo.jl:8

const AGZ_EVAL_DUMMY, AGZ_EVAL_NN_TC, AGZ_EVAL_NN_F32 = Int32(0), Int32(1), Int32(2)
const AGZ_BN_VAR_EPS, AGZ_BN_STD = Int32(0), Int32(1)

# ---- PODs of include/agz.h (field order, types and therefore offsets must match: tests/test_struct_layout.py) ---------------
struct AgzConfig
  board_n::Int32; planes::Int32; filters::Int32; tower_height::Int32
  c_puct::Float64; noise_weight::Float64; noise_alpha::Float64
  max_game_length::Int32; tau_threshold::Int32; parallel_readouts::Int32; max_parallel::Int32
  komi::Float32; resign_threshold::Float64; resign_disable_frac::Float64
  n_games::Int32; readouts::Int32; nodes_per_game::Int32; seed::UInt64
  device::Int32; world_size::Int32; rank::Int32; record_ring::Int32; evaluator::Int32; inject_noise::Int32
  game::Int32; n_in_row::Int32
end

struct AgzGameHeader
  game_id::Int64; n_moves::Int32; result::Int32; resigned::Int32; final_score::Float32; resign_threshold::Float64
end

struct AgzProgress
  moves_played::Int64; games_finished::Int64; games_started::Int64; positions_evaluated::Int64; readouts::Int64; path_nodes::Int64
  games_live::Int32; error::Int32; step_ms::Float32; arena_prunes::Int32
end

function check(h::Ptr{Cvoid}, rc::Int32)
  rc == 0 && return nothing
  msg = unsafe_string(ccall((:agz_last_error, libagz), Cstring, (Ptr{Cvoid},), h))
  rc == 1 && throw(IllegalMove())
  rc == 2 && throw(AssertionError(msg))
  error("libagz error $rc: $msg")
end

# ---- GameEnv: GoEnv (go.jl:1-26) and GomokuEnv (gomoku/gomoku.jl:1-19) -------------------------------------------------------
abstract type GameEnv end
struct GoEnv <: GameEnv
  N::Int; action_space::Int; planes::Int; max_action_space::Int
end
GoEnv(board_size::Int = 19, planes::Int = 17) = GoEnv(board_size, board_size^2 + 1, (planes - 1) ÷ 2, 361)
struct GomokuEnv <: GameEnv
  N::Int; n_in_row::Int; action_space::Int; planes::Int; max_action_space::Int
end
GomokuEnv(board_size::Int = 15, connect_row::Int = 5, planes::Int = 17) = GomokuEnv(board_size, connect_row, board_size^2, (planes - 1) ÷ 2, 361)
Go(n) = GoEnv(n)                                                             # game/env.jl:2
const AGZ_GAME_GO = Int32(0)
const AGZ_GAME_GOMOKU = Int32(1)
game_of(::GoEnv) = (AGZ_GAME_GO, Int32(0))
game_of(env::GomokuEnv) = (AGZ_GAME_GOMOKU, Int32(env.n_in_row))

to_flat0(c::Nothing, env::GameEnv) = env.N^2                                 # coords.jl:5-7, 0-based (not an action of Gomoku)
to_flat0(c::Tuple{Int,Int}, env::GameEnv) = env.N * (c[2] - 1) + (c[1] - 1)
from_flat0(f::Integer, env::GameEnv) = f == env.N^2 ? nothing : (Int(f % env.N) + 1, Int(f ÷ env.N) + 1)

# ---- engine handle ----------------------------------------------------------------------------------------------------------
mutable struct B200Engine
  h::Ptr{Cvoid}
  cfg::AgzConfig
  env::GameEnv
end

"""One engine = one GPU.  Keyword arguments are the fields of agz_config that callers of the reference set through
`MCTSPlayer(...)` / `train(...)` kwargs; everything else keeps the reference's defaults (agz_config_default)."""
function B200Engine(env::GameEnv; n_games::Int = 1024, readouts::Int = 800, tower_height::Int = 19, seed::Integer = 0, device::Int = 0,
                    two_player_mode::Bool = false, resign_threshold::Float64 = -0.9, evaluator::Int32 = AGZ_EVAL_NN_TC,
                    world_size::Int = 1, rank::Int = 0, nodes_per_game::Int = 0)
  r = Ref{AgzConfig}()
  game, n_in_row = game_of(env)
  check(C_NULL, ccall((:agz_config_default_game, libagz), Int32, (Ref{AgzConfig}, Int32, Int32, Int32), r, game, env.N, n_in_row))
  c = r[]
  c = AgzConfig(c.board_n, c.planes, c.filters, tower_height, c.c_puct, c.noise_weight, c.noise_alpha, c.max_game_length,
                two_player_mode ? Int32(-1) : c.tau_threshold, c.parallel_readouts, c.max_parallel, c.komi, resign_threshold,
                c.resign_disable_frac, n_games, readouts, nodes_per_game, UInt64(seed), device, world_size, rank, 0, evaluator,
                two_player_mode ? Int32(0) : Int32(1), c.game, c.n_in_row)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(C_NULL, ccall((:agz_engine_create, libagz), Int32, (Ref{AgzConfig}, Ref{Ptr{Cvoid}}), c, h))
  e = B200Engine(h[], c, env)
  finalizer(close!, e)
  e
end

function close!(e::B200Engine)
  e.h == C_NULL && return
  ccall((:agz_engine_destroy, libagz), Cvoid, (Ptr{Cvoid},), e.h)
  e.h = C_NULL
  nothing
end

set_option!(e::B200Engine, key::String, value::Integer) =
  check(e.h, ccall((:agz_set_option, libagz), Int32, (Ptr{Cvoid}, Cstring, Int64), e.h, key, value))
function get_option(e::B200Engine, key::String)
  v = Ref{Int64}(0)
  check(e.h, ccall((:agz_get_option, libagz), Int32, (Ptr{Cvoid}, Cstring, Ref{Int64}), e.h, key, v))
  v[]
end

# ---- NeuralNet (neural_net.jl:7-33): the three Flux `params` lists save_model writes (train.jl:27-33) -------------------------
"""Parameters as plain arrays in Flux `params` order: base = Conv(W,b), BatchNorm(β,γ), then per ResidualBlock
W1,b1,W2,b2,β1,γ1,β2,γ2 (resnet.jl:3-5); value = Conv, BatchNorm, Dense, Dense; policy = Conv, BatchNorm, Dense.
`bn_mu` / `bn_sigma` per chain in layer order; `bn_mode` = AGZ_BN_VAR_EPS (Flux 0.10.4: σ² with ε = 1e-5) or AGZ_BN_STD (shipped
models/agz_*.bson: moving standard deviation)."""
mutable struct NeuralNet
  env::GameEnv
  tower_height::Int
  params::Vector{Vector{Array{Float32}}}        # [base, value, policy]
  bn_mu::Vector{Vector{Float32}}
  bn_sigma::Vector{Vector{Float32}}
  bn_mode::Int32
end

glorot_uniform(dims...) = (rand(Float32, dims...) .- 0.5f0) .* sqrt(24.0f0 / sum(length(dims) == 2 ? (dims[2], dims[1]) :
                              (dims[end-1] * prod(dims[1:end-2]), dims[end] * prod(dims[1:end-2]))))

"""NeuralNet(env; tower_height = 19): Flux-default initialisation restated (Glorot-uniform Conv / Dense weights, zero biases,
BatchNorm γ = 1, β = 0, μ = 0, σ² = 1), same shapes as neural_net.jl:16-30."""
function NeuralNet(env::GameEnv; tower_height::Int = 19)
  N, C, P = env.N, 256, 2 * env.planes + 1
  z(n) = zeros(Float32, n); o(n) = ones(Float32, n)
  base = Array{Float32}[glorot_uniform(3, 3, P, C), z(C), z(C), o(C)]
  for _ in 1:tower_height
    append!(base, Array{Float32}[glorot_uniform(3, 3, C, C), z(C), glorot_uniform(3, 3, C, C), z(C), z(C), o(C), z(C), o(C)])
  end
  value = Array{Float32}[glorot_uniform(1, 1, C, 1), z(1), z(1), o(1), glorot_uniform(256, N * N), z(256), glorot_uniform(1, 256), z(1)]
  policy = Array{Float32}[glorot_uniform(1, 1, C, 2), z(2), z(2), o(2), glorot_uniform(env.action_space, 2 * N * N), z(env.action_space)]
  nbn = (C * (1 + 2 * tower_height), 1, 2)
  NeuralNet(env, tower_height, [base, value, policy], [z(n) for n in nbn], [o(n) for n in nbn], AGZ_BN_VAR_EPS)
end

"""NeuralNet(env, flux_nn): copy an AlphaGo.NeuralNet (fields base_net, value, policy: Flux chains).  `flux_params(chain)` must
return the chain's parameter arrays in Flux `params` order and `flux_batchnorms(chain)` its BatchNorm layers in order; with
Flux 0.10.4 these are `collect(Flux.params(chain))` and the `BatchNorm` entries of a walk over `chain.layers` (ResidualBlock:
its `norm_layers`, resnet.jl:3-5).  Passed in as functions so that this file does not depend on Flux."""
function NeuralNet(env::GameEnv, flux_nn; tower_height::Int, flux_params::Function, flux_batchnorms::Function, cpu::Function = identity)
  chains = (flux_nn.base_net, flux_nn.value, flux_nn.policy)
  ps = [Array{Float32}[Array{Float32}(cpu(p)) for p in flux_params(ch)] for ch in chains]
  mu = [reduce(vcat, [vec(Float32.(cpu(b.μ))) for b in flux_batchnorms(ch)]) for ch in chains]
  s2 = [reduce(vcat, [vec(Float32.(cpu(b.σ²))) for b in flux_batchnorms(ch)]) for ch in chains]
  NeuralNet(env, tower_height, ps, mu, s2, AGZ_BN_VAR_EPS)
end

flat(list::Vector{<:Array{Float32}}) = reduce(vcat, [vec(a) for a in list])      # column-major, as the ABI expects

function load_weights!(e::B200Engine, nn::NeuralNet)
  for chain in 0:2
    f = flat(nn.params[chain + 1]); mu = nn.bn_mu[chain + 1]; sg = nn.bn_sigma[chain + 1]
    GC.@preserve f mu sg begin
      check(e.h, ccall((:agz_net_set_params, libagz), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Csize_t), e.h, chain, f, length(f)))
      check(e.h, ccall((:agz_net_set_bn_stats, libagz), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Csize_t, Int32),
                       e.h, chain, mu, sg, length(mu), nn.bn_mode))
    end
  end
  nothing
end

"""Current parameters / running statistics of the engine back into `nn` (what save_model writes after training, train.jl:14-35)."""
function read_weights(e::B200Engine, nn::NeuralNet)
  for chain in 0:2
    n = ccall((:agz_net_param_count, libagz), Csize_t, (Ptr{Cvoid}, Int32), e.h, chain)
    f = Vector{Float32}(undef, n)
    GC.@preserve f check(e.h, ccall((:agz_net_get_params, libagz), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Csize_t), e.h, chain, f, n))
    off = 0
    for a in nn.params[chain + 1]
      copyto!(a, 1, f, off + 1, length(a)); off += length(a)
    end
    nb = ccall((:agz_net_bn_count, libagz), Csize_t, (Ptr{Cvoid}, Int32), e.h, chain)
    mu = Vector{Float32}(undef, nb); sg = Vector{Float32}(undef, nb); mode = Ref{Int32}(0)
    GC.@preserve mu sg check(e.h, ccall((:agz_net_get_bn_stats, libagz), Int32,
                                        (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Csize_t, Ref{Int32}), e.h, chain, mu, sg, nb, mode))
    nn.bn_mu[chain + 1] = mu; nn.bn_sigma[chain + 1] = sg; nn.bn_mode = mode[]
  end
  nn
end

"""(nn::NeuralNet)(boards_hist, to_play) -> (π :: A×B, v :: 1×B)  (neural_net.jl:57-68).  `boards_hist` is N²×8×B Int8: board k
moves ago in flat (column-major) order, what features.jl:7-14 rebuilds from board_deltas; `to_play` is B Int8."""
function (nn::NeuralNet)(e::B200Engine, boards_hist::Array{Int8,3}, to_play::Vector{Int8})
  B = length(to_play); A = nn.env.action_space
  @assert size(boards_hist) == (nn.env.N^2, 8, B)
  load_weights!(e, nn)
  π = Matrix{Float32}(undef, A, B); v = Matrix{Float32}(undef, 1, B)
  GC.@preserve boards_hist to_play π v check(e.h, ccall((:agz_net_forward, libagz), Int32,
      (Ptr{Cvoid}, Int32, Ptr{Int8}, Ptr{Int8}, Int32, Ptr{Float32}, Ptr{Float32}), e.h, AGZ_EVAL_NN_TC, boards_hist, to_play, B, π, v))
  π, v
end

# ---- MCTSPlayer as callers of selfplay read it (mcts_play.jl:3-24; train.jl:57-58,73) -----------------------------------------
"""The finished player `selfplay` returns.  `root` holds what callers read from `player.root.position` (`n`, `recent`, komi);
`moves` are the reference's coordinates ((i, j) 1-based or `nothing` for a pass)."""
struct RootView
  n::Int
  recent::Vector{Union{Nothing,Tuple{Int,Int}}}
  komi::Float32
end
struct MCTSPlayer
  env::GameEnv
  root::RootView
  searches_π::Vector{Vector{Float32}}
  qs::Vector{Float32}
  result::Int
  result_string::String
  num_readouts::Int
  resign_threshold::Float64
  game_id::Int64
end

function result_string(hd::AgzGameHeader, env::GameEnv = GoEnv())   # set_result! (mcts_play.jl:100-108), result_string (board.jl:546-555)
  hd.resigned != 0 && return hd.result == 1 ? "B+R" : "W+R"
  if env isa GomokuEnv                                    # gomoku/board.jl:184-193: final_score carries the winner's colour
    return hd.final_score > 0 ? "B" : (hd.final_score < 0 ? "W" : "DRAW")
  end
  hd.final_score > 0 && return "B+" * string(round(hd.final_score; digits = 1))
  hd.final_score < 0 && return "W+" * string(round(abs(hd.final_score); digits = 1))
  "DRAW"
end

function players_from_records(env::GameEnv, readouts::Int, hd::Vector{AgzGameHeader}, moves::Matrix{Int16}, qs::Matrix{Float32},
                              pis::Array{Float32,3}, n::Int)
  out = MCTSPlayer[]
  for g in 1:n
    nm = Int(hd[g].n_moves)
    mv = Union{Nothing,Tuple{Int,Int}}[from_flat0(moves[t, g], env) for t in 1:nm]
    push!(out, MCTSPlayer(env, RootView(nm, mv, 7.5f0), [pis[:, t, g] for t in 1:nm], qs[1:nm, g], Int(hd[g].result), result_string(hd[g], env),
                          readouts, hd[g].resign_threshold, hd[g].game_id))
  end
  out
end

"""selfplay(env, nn, num_ro = 800; n_games = 1) -> the finished player (n_games == 1, like src/selfplay.jl:1) or a Vector of them:
all games run concurrently on the GPU, game ids 0 … n_games-1 key the random streams (oracle/rng.py)."""
function selfplay(env::GameEnv, nn::NeuralNet, num_ro::Int = 800; n_games::Int = 1, seed::Integer = 0, device::Int = 0, engine = nothing)
  e = engine === nothing ? B200Engine(env; n_games = n_games, readouts = num_ro, tower_height = nn.tower_height, seed = seed, device = device) : engine
  load_weights!(e, nn)
  L = Int(e.cfg.max_game_length) + 2; A = env.action_space
  hd = Vector{AgzGameHeader}(undef, n_games); moves = Matrix{Int16}(undef, L, n_games); qs = Matrix{Float32}(undef, L, n_games)
  pis = Array{Float32}(undef, A, L, n_games); vis = Array{Float32}(undef, A, L, n_games)
  GC.@preserve hd moves qs pis vis check(e.h, ccall((:agz_selfplay_run, libagz), Int32,
      (Ptr{Cvoid}, Int32, Ptr{AgzGameHeader}, Ptr{Int16}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), e.h, n_games, hd, moves, qs, pis, vis))
  engine === nothing && close!(e)
  ps = players_from_records(env, num_ro, hd, moves, qs, pis, n_games)
  n_games == 1 ? ps[1] : ps
end

"""extract_data(player) -> (positions, πs, results)  (mcts_play.jl:126-139): `positions[t]` is the board before move t as an N×N
Int8 matrix (+1 Black, -1 White, row i from the top, column j) together with the side to move; results repeat the game result
from Black's view (board.jl:574).  `replay(board, move, color)` replays one move with the rules of board.jl:451-509 -- inside
AlphaGo.jl pass `(pos, c) -> play_move!(pos, c)` over GoPosition instead; the engine-side equivalent (no host replay at all) is
`agz_replay_gather`, which packs the same tuples on the device."""
function extract_data(player::MCTSPlayer; position0, play)
  @assert length(player.searches_π) == player.root.n
  positions = Any[]; pos = position0
  for mv in player.root.recent
    push!(positions, pos)
    pos = play(pos, mv)
  end
  positions, player.searches_π, fill(player.result, length(positions))
end

# ---- replay ring + training (train.jl:4-12,38-92; neural_net.jl:75-101) -------------------------------------------------------
"""get_replay_batch: `batch` distinct tuples drawn uniformly from the device replay ring: (boards_hist N²×8×B, to_play, π A×B, z)."""
function get_replay_batch(e::B200Engine, batch::Int; seed::Integer = 0)
  N2 = e.env.N^2; A = e.env.action_space
  bh = Array{Int8}(undef, N2, 8, batch); tp = Vector{Int8}(undef, batch); pis = Matrix{Float32}(undef, A, batch)
  zs = Vector{Int8}(undef, batch); idx = Vector{Int64}(undef, batch)
  GC.@preserve bh tp pis zs idx check(e.h, ccall((:agz_replay_sample_hist, libagz), Int32,
      (Ptr{Cvoid}, Int32, UInt64, Ptr{Int8}, Ptr{Int8}, Ptr{Float32}, Ptr{Int8}, Ptr{Int64}), e.h, batch, UInt64(seed), bh, tp, pis, zs, idx))
  bh, tp, pis, zs
end

"""train(env; …) -> NeuralNet  (train.jl:38-92).  `concurrent` games run at once with the current network; every finished game
triggers `epochs` optimisation steps on one uniform batch of the replay ring once `start_training_after` tuples are there --
the reference's ratio of steps to games.  Sampling, feature building and the step stay on the device
(agz_train_step_from_replay); the host only sees the loss."""
function train(env::GameEnv; num_games::Int = 25000, memory_size::Int = 500000, batch_size::Int = 32, epochs::Int = 1, ckp_freq::Int = 1000,
               readouts::Int = 800, tower_height::Int = 19, model = nothing, start_training_after::Int = 50000, concurrent::Int = 1024,
               seed::Integer = 0, lr::Float32 = 2f-2, momentum::Float32 = 9f-1, on_checkpoint = nn -> nothing, verbose::Bool = true)
  cur_nn = model === nothing ? NeuralNet(env; tower_height = tower_height) : model
  e = B200Engine(env; n_games = min(concurrent, num_games), readouts = readouts, tower_height = cur_nn.tower_height, seed = seed)
  set_option!(e, "replay.capacity", memory_size)
  load_weights!(e, cur_nn)
  check(e.h, ccall((:agz_selfplay_start, libagz), Int32, (Ptr{Cvoid}, Int64), e.h, num_games))
  done = 0; last_ckp = 0; pr = Ref{AgzProgress}(); total = Ref{Int64}(0); nrec = Ref{Int32}(0)
  hd = Vector{AgzGameHeader}(undef, Int(e.cfg.n_games) * 2)
  while done < num_games
    check(e.h, ccall((:agz_selfplay_step, libagz), Int32, (Ptr{Cvoid}, Int32, Ref{AgzProgress}), e.h, 8, pr))
    pr[].error != 0 && error("a game stopped on the device with status $(pr[].error)")
    check(e.h, ccall((:agz_replay_gather, libagz), Int32, (Ptr{Cvoid}, Ref{Int64}), e.h, total))
    GC.@preserve hd check(e.h, ccall((:agz_selfplay_harvest, libagz), Int32,
        (Ptr{Cvoid}, Int32, Ptr{AgzGameHeader}, Ptr{Int16}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ref{Int32}),
        e.h, length(hd), hd, C_NULL, C_NULL, C_NULL, C_NULL, nrec))
    nrec[] == 0 && pr[].games_live == 0 && break
    for r in 1:nrec[]
      done += 1
      if total[] >= start_training_after && total[] >= batch_size
        loss = Ref{Float32}(0); acc = 0f0
        for ep in 1:epochs
          check(e.h, ccall((:agz_train_step_from_replay, libagz), Int32, (Ptr{Cvoid}, Int32, UInt64, Float32, Float32, Ref{Float32}),
                           e.h, batch_size, UInt64(seed) * 1000003 + UInt64(done) * 131 + UInt64(ep), lr, momentum, loss))
          acc += loss[]
        end
        verbose && println("Episode $done over. Loss: $(acc / epochs). Winner: $(result_string(hd[r], env)). Moves: $(hd[r].n_moves).")
      end
      if done ÷ ckp_freq > last_ckp
        last_ckp = done ÷ ckp_freq
        on_checkpoint(read_weights(e, cur_nn))                    # save_model (train.jl:14-35) is the caller's BSON code
        verbose && print("Model saved. ")
      end
    end
  end
  read_weights(e, cur_nn)
  close!(e)
  cur_nn
end

# ---- evaluate (neural_net.jl:103-158): all gating games at once, one engine per player ----------------------------------------
function evaluate(env::GameEnv, black_net::NeuralNet, white_net::NeuralNet; num_games::Int = 400, ro::Int = 800, seed::Integer = 0, verbose::Bool = false)
  G = num_games
  engines = [B200Engine(env; n_games = G, readouts = ro, tower_height = net.tower_height, seed = seed + k - 1, two_player_mode = true)
             for (k, net) in enumerate((black_net, white_net))]
  for (e, net) in zip(engines, (black_net, white_net))
    load_weights!(e, net)
    check(e.h, ccall((:agz_match_start, libagz), Int32, (Ptr{Cvoid}, Ptr{Int64}), e.h, C_NULL))
  end
  alive = trues(G); score = zeros(Float32, G); num_move = 0
  moves = Vector{Int32}(undef, G); res = Vector{Int32}(undef, G); sc = Vector{Float32}(undef, G); done = Vector{Int32}(undef, G)
  while any(alive)
    active, inactive = isodd(num_move) ? (engines[2], engines[1]) : (engines[1], engines[2])
    act = UInt8.(alive)
    GC.@preserve act moves res sc check(active.h, ccall((:agz_match_search, libagz), Int32,
        (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}), active.h, act, moves, res, sc))
    for g in findall(alive .& (res .!= 0)); score[g] = sc[g]; alive[g] = false; end          # forced resignation (:129-133)
    mv = Int32[alive[g] ? moves[g] : Int32(-1) for g in 1:G]
    for p in (active, inactive)
      GC.@preserve mv done sc check(p.h, ccall((:agz_match_play, libagz), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}), p.h, mv, done, sc))
      if p === active
        for g in findall(alive .& (done .!= 0)); score[g] = sc[g]; alive[g] = false; end      # is_done(active) (:140-146)
      end
    end
    num_move += 1
  end
  foreach(close!, engines)
  won = count(>(0f0), score)                                                                 # result(black.root.position) == BLACK (:150)
  verbose && print("Won $won / $G. Win rate: $(won / G). ")
  won / G >= 0.55
end

end # module
