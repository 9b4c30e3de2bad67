# run_reference.jl -- times the UNMODIFIED reference (tejank10/AlphaGo.jl) on its own CPU path: BASELINE config C1/C2-sample.
#
# STATUS: UNVERIFIED HERE (no Julia binary in the build image, no network for its packages).  What it needs on a machine that has
# them: Julia >= 1.0 with the package versions of the reference's Manifest.toml (Flux 0.10.4 / Zygote) OR the Tracker-era Flux the
# sources were written against (the committed code uses `param`, `.data`, `back!`: src/mcts_play.jl:90, src/train.jl:2).
#
#   julia baseline/run_reference.jl /path/to/AlphaGo.jl [board=9] [readouts=400] [tower_height=6] [moves=8]
#
# The committed reference does not run as written: src/selfplay.jl:9 reads `rand() < 0.05 : -1.0 : -0.9` (a range expression, not
# the ternary it means; SURVEY.md section 8c).  The script works on a temporary copy with that one line patched and touches nothing
# else, then plays `moves` moves of one self-play game and prints one JSON line in the format of bench.py --impl reference.
using Printf

ref = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia run_reference.jl <AlphaGo.jl checkout> [board] [readouts] [tower_height] [moves]")
board = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 9
readouts = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 400
tower = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 6
max_moves = length(ARGS) >= 5 ? parse(Int, ARGS[5]) : 8

work = mktempdir()
cp(ref, joinpath(work, "AlphaGo.jl"); force = true)
sp = joinpath(work, "AlphaGo.jl", "src", "selfplay.jl")
chmod(sp, 0o644)
src = read(sp, String)
patched = replace(src, "rand() < 0.05 : -1.0 : -0.9" => "rand() < 0.05 ? -1.0 : -0.9")
patched == src && @warn "src/selfplay.jl:9 did not contain the expected typo; running the file as it is"
write(sp, patched)

push!(LOAD_PATH, joinpath(work, "AlphaGo.jl", "src"))
include(joinpath(work, "AlphaGo.jl", "src", "AlphaGo.jl"))
using .AlphaGo

env = AlphaGo.GoEnv(board)
nn = AlphaGo.NeuralNet(env; tower_height = tower)

# the loop of src/selfplay.jl:11-43 with a move budget and a clock around it (selfplay itself has no hook to stop early)
function timed_moves(env, nn, readouts, max_moves)
  player = AlphaGo.MCTSPlayer(env, nn; num_readouts = readouts, resign_threshold = -1.0)
  AlphaGo.initialize_game!(player)
  first_node = AlphaGo.select_leaf(player.root)
  prob, val = nn(first_node.position)
  AlphaGo.incorporate_results!(first_node, prob.data, val.data, first_node)
  t0 = time(); moves = 0
  while moves < max_moves
    AlphaGo.inject_noise!(player.root)
    current = AlphaGo.N(player.root)
    while AlphaGo.N(player.root) < current + readouts
      AlphaGo.tree_search!(player)
    end
    move = AlphaGo.pick_move(player)
    AlphaGo.play_move!(player, move)
    moves += 1
    AlphaGo.is_done(player.root) && break
  end
  moves, time() - t0
end

timed_moves(env, nn, readouts, 1)                      # compile
moves, el = timed_moves(env, nn, readouts, max_moves)
@printf("{\"impl\": \"reference\", \"kind\": \"reference\", \"metric\": \"self-play moves/sec (%dx%d, %d readouts)\", \"value\": %.6f, \"unit\": \"moves/s\", \"moves\": %d, \"seconds\": %.3f, \"threads\": %d, \"config\": {\"workload\": \"%dx%d Go, 1 game, %d readouts/move, tower_height %d, Flux CPU\"}}\n",
        board, board, readouts, moves / el, moves, el, Threads.nthreads(), board, board, readouts, tower)
