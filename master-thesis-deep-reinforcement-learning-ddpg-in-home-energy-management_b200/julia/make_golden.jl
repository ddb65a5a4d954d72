# make_golden.jl — golden vectors from the UNMODIFIED reference (RL-SHEMS), for the parity tests of this repository.
#
# The build image of this repository has no Julia, so its CPU oracle is pinned only by hand-derived vectors and by an
# independent numpy restatement.  This script closes that gap wherever Julia 1.6 and the reference's packages are available:
# it `include`s the reference's own files, runs them on the committed inputs under tests/golden/julia_inputs/ and writes
# tests/golden/reference_julia/*.csv.  tests/test_julia_golden.py consumes those files when present (bit-exact for the
# environment, 1e-5 for replay()) and is skipped otherwise.
#
#   JOB_ID=9800 julia --project=<reference>/RL-SHEMS make_golden.jl <reference>/RL-SHEMS <this repo>/tests/golden [env|ddpg|all]
#
# Nothing of the reference is modified or copied: RL_environments/envs/shems_LU1.jl, algorithms/DDPG.jl and
# src/memory_plotting_saving.jl are included where they lie.  What this file restates is only configuration that the reference
# keeps in input.jl (the tuned hyper-parameters of README.md:68-86 and the noise structs of input.jl:190-215).

length(ARGS) >= 2 || error("usage: julia make_golden.jl <reference>/RL-SHEMS <repo>/tests/golden [env|ddpg|all]")
const REF = ARGS[1]
const GOLD = ARGS[2]
const PART = length(ARGS) >= 3 ? ARGS[3] : "all"
const IN = joinpath(GOLD, "julia_inputs")
const OUT = joinpath(GOLD, "reference_julia")
mkpath(OUT)
haskey(ENV, "JOB_ID") || (ENV["JOB_ID"] = "9800")            # charger id = (JOB_ID ÷ 100) % 100 = 98  (shems_LU1.jl:45)

using CSV, DataFrames, Random
using Distributions: Uniform

include(joinpath(REF, "RL_environments", "envs", "shems_LU1.jl"))
using .ShemsEnv_LU1: Shems, reset!, step!, action, finished

const SERIES = joinpath(IN, "Charger98_all_test_fix.csv")
const STATE_FIELDS = [:Soc_b, :Soc_ev, :c_ev, :d_e, :g_e, :p_buy, :h_cos, :h_sin, :season]
const TRACE_NAMES = ["index", "c_ev", "EV_target", "EV", "Soc_ev", "rewards", "profit", "discomfort", "penalty", "PV_DE", "B_DE", "GR_DE",
                     "PV_B", "PV_GR", "PV_EV", "B_EV", "GR_EV", "EX_EV", "GR_B", "B_GR", "B", "B_tar", "Soc_b"]
state64(env) = Float64[Float64(getproperty(env.state, f)) for f in STATE_FIELDS]
write_matrix(path, m, names) = CSV.write(path, DataFrame(m, names))

# ------------------------------------------------------------------------------------------------ environment
function golden_cases()
    cases = CSV.read(joinpath(IN, "lu1_cases.csv"), DataFrame)
    env = Shems(72, SERIES)
    out = Matrix{Float64}(undef, nrow(cases), 2 + 9 + 1 + 23)
    for i in 1:nrow(cases)
        for f in STATE_FIELDS
            setproperty!(env.state, f, Float32(cases[i, f]))
        end
        env.idx = Int(cases[i, :idx])
        a = Float32[cases[i, :a1], cases[i, :a2]]
        tr = cases[i, :track]
        s = copy(env.state)
        if tr == 0
            r, s2 = step!(env, s, a)                                   # learning phase (DDPG.jl:205)
            res = fill(NaN, 1, 23)
        else
            r, s2, res = step!(env, s, a, track=tr)                    # DRL inference (track = 1) / rule-based (track < 0)
        end
        out[i, :] = vcat(Float64(cases[i, :case]), Float64(r), Float64.(s2), Float64(env.idx), vec(Float64.(res)))
    end
    write_matrix(joinpath(OUT, "lu1_cases_out.csv"), out, vcat(["case", "reward"], "s2_" .* string.(STATE_FIELDS), ["idx"], TRACE_NAMES))
end

function golden_rule_episode()
    # inference(env; track = -0.5) (memory_plotting_saving.jl:62-71 -> episode! DDPG.jl:186-242 with rng_ep = -1): the series has
    # 2999 rows, so 2998 steps can read row idx+1
    nsteps = 2998
    env = Shems(nsteps, SERIES)
    reset!(env, rng=-1)
    rows = Matrix{Float64}(undef, nsteps, 23 + 9)
    reward_eps = 0f0                                                   # DDPG.jl:190 starts the sum as a Float32; += Float64 promotes
    for step in 1:nsteps
        s = copy(env.state)
        a = action(env, -0.5)
        r, s2, res = step!(env, s, a, track=-0.5)
        reward_eps += r
        rows[step, :] = vcat(vec(Float64.(res)), Float64.(s2))
    end
    write_matrix(joinpath(OUT, "lu1_rule_episode.csv"), rows, vcat(TRACE_NAMES, "s2_" .* string.(STATE_FIELDS)))
    write_matrix(joinpath(OUT, "lu1_rule_episode_return.csv"), reshape([Float64(reward_eps)], 1, 1), ["reward_eps"])
end

function golden_drl_episode()
    tape = CSV.read(joinpath(IN, "lu1_tape.csv"), DataFrame)
    env = Shems(72, SERIES)
    reset!(env, rng=1234)
    idx0, socb0 = env.idx, Float64(env.state.Soc_b)
    rows = Matrix{Float64}(undef, nrow(tape), 2 + 23 + 9)
    for t in 1:nrow(tape)
        s = copy(env.state)
        r, s2, res = step!(env, s, Float32[tape[t, :a1], tape[t, :a2]], track=1)
        rows[t, :] = vcat(Float64(idx0), socb0, vec(Float64.(res)), Float64.(s2))
    end
    write_matrix(joinpath(OUT, "lu1_drl_episode.csv"), rows, vcat(["idx0", "socb0"], TRACE_NAMES, "s2_" .* string.(STATE_FIELDS)))
end

function golden_resets()
    # reset!(env; rng) (shems_LU1.jl:206-262): the two draws of :224-225 next to where the window-shift loop ends up
    maxsteps = 72
    env = Shems(maxsteps, SERIES)
    nrows = nrow(CSV.read(SERIES, DataFrame))
    b = ShemsEnv_LU1.b
    out = Matrix{Float64}(undef, 300, 4 + 9)
    for rng in 1:300
        idx_draw = rand(MersenneTwister(rng), 1:(nrows - maxsteps))
        socb_draw = Float32(rand(MersenneTwister(rng), Uniform(b.soc_min, b.soc_max)))
        reset!(env, rng=rng)
        out[rng, :] = vcat(Float64(rng), Float64(idx_draw), Float64(socb_draw), Float64(env.idx), state64(env))
    end
    write_matrix(joinpath(OUT, "lu1_resets.csv"), out, vcat(["rng", "idx_draw", "socb_draw", "idx"], string.(STATE_FIELDS)))
    # rng == -1
    reset!(env, rng=-1)
    write_matrix(joinpath(OUT, "lu1_reset_deterministic.csv"), reshape(vcat(Float64(env.idx), state64(env)), 1, 10), vcat(["idx"], string.(STATE_FIELDS)))
end

if PART in ("env", "all")
    golden_cases()
    golden_rule_episode()
    golden_drl_episode()
    golden_resets()
    println("environment golden vectors written to ", OUT)
end

# ------------------------------------------------------------------------------------------------ replay()  (DDPG.jl:121-145)
if PART in ("ddpg", "all")
    using Flux, Zygote, Printf
    using Flux.Optimise: update!
    using Statistics: mean, std, median
    using DataStructures: CircularBuffer
    using Distributions: sample, Normal
    using Dates

    # configuration the reference keeps in input.jl (tuned values: README.md:68-86; input_templates/input09_08_on_01-09_eval.jl)
    global STATE_SIZE, ACTION_SIZE, L1, L2, BATCH_SIZE, MEM_SIZE = 9, 2, 250, 500, 120, 512
    global MIN_EXP_SIZE = MEM_SIZE
    global γ, τ, η_act, η_crit = 0.99f0, 1f-3, 1f-4, 1f-3
    global noise_type = "gn"
    global rng_run = 1231
    global opt_crit = ADAM(η_crit)                                     # input.jl:126-127
    global opt_act = ADAM(η_act)
    global memory = CircularBuffer{Any}(MEM_SIZE)                      # input.jl:140
    global ACTION_BOUND_HI, ACTION_BOUND_LO = (1f0, 1f0), (0f0, 0f0)
    global EP_LENGTH = Dict("train" => 72)
    global NUM_EP, test_every, test_runs, seed_ini, current_episode = 1, 100, 100, 123, 0
    struct OUNoise; μ; σ; θ; dt; X; end                                # input.jl:190-215
    struct GNoise; μ; σ_act; σ_trg; end
    mutable struct EpsNoise; ζ; ξ; ξ_min; end
    mutable struct ParamNoise; μ; σ_current; σ_target; adoption; end
    global gn = GNoise(0f0, 0.1f0, 0.2f0)

    include(joinpath(REF, "algorithms", "DDPG.jl"))
    include(joinpath(REF, "src", "memory_plotting_saving.jl"))

    # the problem of tests/julia_golden_spec.py, generated here with the same counter-based generator
    function splitmix_uniform(seed::Integer, n::Integer)
        u = Vector{Float64}(undef, n)
        for i in 0:n-1
            z = (UInt64(seed) << 32) + UInt64(i)
            z += 0x9E3779B97F4A7C15
            z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
            z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
            z = z ⊻ (z >> 31)
            u[i+1] = Float64(z >> 11) * (1.0 / 9007199254740992.0)
        end
        return u
    end
    tensor_seed(net, layer, is_bias) = 1000 + 10 * net + 2 * layer + (is_bias ? 1 : 0)       # net, layer 0-based
    nets = [actor, critic, actor_target, critic_target]
    for (n, net) in enumerate(nets)
        ps = collect(Flux.params(net))                                  # W1, b1, W2, b2, W3, b3
        for k in 1:3
            W, b = ps[2k-1], ps[2k]
            o, i = size(W)
            u = splitmix_uniform(tensor_seed(n - 1, k - 1, false), i * o)
            w = k < 3 ? (u .- 0.5) .* sqrt(24.0 / (i + o)) : 6e-3 .* u .- 3e-3
            copyto!(W, reshape(Float32.(w), o, i))                      # Flux order: out x in, column-major
            copyto!(b, Float32.((splitmix_uniform(tensor_seed(n - 1, k - 1, true), o) .- 0.5) .* 0.02))
        end
    end
    function states(seed0)
        u = [splitmix_uniform(seed0 + k, MEM_SIZE) for k in 0:8]
        s = hcat(u[1] .* 6.75, u[2], floor.(u[3] .* 42.0) .- 1.0, 0.2 .+ u[4] .* 5.8, u[5] .* 20.0, fill(0.4, MEM_SIZE),
                 2.0 .* u[7] .- 1.0, 2.0 .* u[8] .- 1.0, 1.0 .+ floor.(u[9] .* 4.0))
        return Float32.(permutedims(s))                                 # 9 x n
    end
    S1, S2 = states(2000), states(2100)
    Amat = Float32.(permutedims(hcat(2.0 .* splitmix_uniform(2200, MEM_SIZE) .- 1.0, 2.0 .* splitmix_uniform(2201, MEM_SIZE) .- 1.0)))
    R = -(splitmix_uniform(2300, MEM_SIZE) .* 5.0)                      # Float64, like env.reward
    for j in 1:MEM_SIZE
        remember(S1[:, j], Amat[:, j], R[j], S2[:, j], false)          # memory_plotting_saving.jl:46-47
    end
    global s_min = minimum(S1, dims=2) |> gpu                          # what driver:30 holds, here over the whole memory (no RNG)
    global s_max = maximum(S1, dims=2) |> gpu

    digest_positions(n) = n <= 64 ? collect(1:n) : collect((0:63) .* (n ÷ 64) .+ 1)
    function digest(p)
        x = Float64.(vec(cpu(p)))
        d = vcat(sum(x), sum(abs.(x)), x[digest_positions(length(x))])
        return vcat(d, fill(NaN, 66 - length(d)))
    end
    rows = Matrix{Float64}(undef, 0, 4 + 66)
    idxrows = Matrix{Float64}(undef, 0, 3)
    for u in 1:3
        idx = sample(MersenneTwister(u), 1:length(memory), BATCH_SIZE)  # the draws getData(BATCH_SIZE, rng_dt=u) makes (:33)
        mb = sample(MersenneTwister(u), memory, BATCH_SIZE)
        all(mb[j][3] == memory[idx[j]][3] for j in 1:BATCH_SIZE) || error("index stream of sample(rng, 1:n, B) differs from sample(rng, memory, B)")
        replay(rng_rpl=u)                                               # the unmodified update
        for (n, net) in enumerate(nets), (t, p) in enumerate(collect(Flux.params(net)))
            global rows = vcat(rows, reshape(vcat(Float64(u), Float64(n - 1), Float64(t - 1), Float64(length(p)), digest(p)), 1, :))
        end
        global idxrows = vcat(idxrows, hcat(fill(Float64(u), BATCH_SIZE), collect(1.0:BATCH_SIZE), Float64.(idx)))
    end
    write_matrix(joinpath(OUT, "ddpg_replay.csv"), rows, vcat(["update", "net", "tensor", "len", "sum", "sumabs"], "d" .* string.(0:63)))
    write_matrix(joinpath(OUT, "ddpg_indices.csv"), idxrows, ["update", "j", "idx1"])
    println("replay() golden vectors written to ", OUT)
end
