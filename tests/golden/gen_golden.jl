# gen_golden.jl -- pins the network oracle against the reference itself: runs the reference's NeuralNet (Flux) with the shipped
# models/weights/agz_*.bson on the committed fixture positions and writes its (π, v) next to them.
#
# STATUS: UNVERIFIED HERE (no Julia binary / Flux / BSON.jl in the build image).  Until someone runs it, network parity stays
# "unpinned" (oracle/__init__.py, DESIGN.md section 2).  Once tests/golden/agz_shipped_9x9_julia.json exists,
# tests/test_abi_nn.py::test_reference_julia_outputs compares both the oracle and the CUDA engine with it at 1e-3.
#
#   python tests/golden/export_fixture_positions.py          # agz_shipped_9x9.npz -> agz_shipped_9x9_positions.json (committed)
#   julia tests/golden/gen_golden.jl /path/to/AlphaGo.jl      # -> tests/golden/agz_shipped_9x9_julia.json
#
# The input JSON holds, per position, the 8 history boards (flat column-major Int8 lists, current first) and to_play; the feature
# tensor is rebuilt exactly as src/features.jl:3-26 does (planes 2k-1 / 2k = board_k .== to_play / .== -to_play, plane 17 = to_play).
using BSON, Flux

ref = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia gen_golden.jl <AlphaGo.jl checkout>")
here = @__DIR__

# minimal JSON reader / writer for the fixture (arrays of numbers only), to avoid a JSON.jl dependency
parse_fixture(path) = include_string(Main, replace(replace(read(path, String), "{" => "Dict(", ), "}" => ")") |> s -> replace(s, "\":" => "\"=>"))
fix = parse_fixture(joinpath(here, "agz_shipped_9x9_positions.json"))
N = 9; B = length(fix["to_play"])

feats = zeros(Float64, N, N, 17, B)
for b in 1:B
  tp = fix["to_play"][b]
  for k in 1:8
    board = reshape(Int8.(fix["boards_hist"][b][k]), N, N)          # column-major: board[i, j], flat f = N*(j-1) + i
    feats[:, :, 2k - 1, b] = board .== tp
    feats[:, :, 2k, b] = board .== -tp
  end
  feats[:, :, 17, b] .= tp
end

# the shipped net: 9x9, 17 planes, 256 filters, tower_height 0 (SURVEY.md section 2 row 15).  models/agz_*.bson hold the Flux structs.
BSON.@load joinpath(ref, "models", "agz_base.bson") base_net
BSON.@load joinpath(ref, "models", "agz_value.bson") value
BSON.@load joinpath(ref, "models", "agz_policy.bson") policy
Flux.testmode!(base_net); Flux.testmode!(value); Flux.testmode!(policy)
common = base_net(feats)
π = Flux.data(policy(common)); v = Flux.data(value(common))

open(joinpath(here, "agz_shipped_9x9_julia.json"), "w") do io
  print(io, "{\"pi\": [", join(["[" * join(string.(Float64.(π[:, b])), ", ") * "]" for b in 1:B], ", "), "], ")
  print(io, "\"v\": [", join(string.(Float64.(vec(v))), ", "), "], \"flux_version\": \"", string(Flux), "\"}\n")
end
println("wrote agz_shipped_9x9_julia.json: ", size(π), " ", v[1:min(4, B)])
