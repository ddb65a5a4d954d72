# DdpgB200.jl — the DDPG half of the drop-in: RL-SHEMS/DDPG_reinforce_charger_v1.jl runs UNCHANGED on libshems_b200.so.
#
# How the reference driver reaches this file, and why nothing of the reference has to be edited:
#   * the driver includes, in this order, the input file, algorithms/DDPG.jl and src/memory_plotting_saving.jl
#     (DDPG_reinforce_charger_v1.jl:10, :24, :25), then calls populate_memory, min_max_buffer, run_episodes, saveBSON, inference
#     (:28-108).  The INPUT FILE is the configuration the user edits anyway (JOB_ID grid search, paths, algo); it gets four lines
#     (INTEGRATION.md §2): the two env lines point at ShemsB200.jl, `memory = DeviceMemory(MEM_SIZE)`, and `include` of this file
#     as its last line.
#   * `Shems(maxsteps, path)` then builds a `ShemsCuda <: Shems`.  This file defines `populate_memory`, `episode!` and `run_episodes`
#     for `::ShemsCuda`; DDPG.jl and memory_plotting_saving.jl, included LATER by the driver, add their methods for `::Shems` to the
#     same generic functions — the more specific methods here win by dispatch, whatever the include order.  `inference`
#     (memory_plotting_saving.jl:62-89) stays the reference's: it calls `episode!`, which lands here.
#   * the functions memory_plotting_saving.jl defines WITHOUT an env argument — remember, getData, min_max_buffer (:31-53) — are kept
#     too: they act on the global `memory`, which is a `DeviceMemory` (the replay ring on the GPU) with `push!`, `length` and
#     `sample` methods, so `remember` pushes to the device and `min_max_buffer` draws the reference's own MersenneTwister indices.
#   * DDPG.jl still builds its Flux `actor`, `critic` and targets (:30-46).  They are adopted as the learner's INITIAL WEIGHTS on first
#     use — a shimmed run starts from bit-identical weights for identical seeds — and `actor` is refreshed from the device before every
#     saveBSON, so the BSON files the driver writes (:45, DDPG.jl:282-289) are the reference's.  `global actor = ...; inference(...)`
#     (driver :93-101) pushes that actor back to the device.
#   * all random draws are made in Julia exactly where the reference makes them (reset! draws, populate_memory's actions,
#     sample_noise — the reference's own function is called —, the minibatch indices of getData): SHEMS_B200_RNG=julia (default).
#     SHEMS_B200_RNG=philox moves them to the device (one ddpg_episode call per training episode, no host round trip per step).
#
# NOTE: Julia is not installed in the build/CI image of this repository: this file is reviewed, never executed there.  Every entry
# point it calls is exercised through the identical C ABI by the Python ctypes harness, in the driver's own call order
# (tests/test_driver_walk_gpu.py), and tests/test_julia_shim_static.py checks this file against the reference's sources
# (signatures, keyword names, the globals it reads, the ccall argument lists against include/shems_b200.h).

import CUDA
import StatsBase
using Random: MersenneTwister, AbstractRNG
using .ShemsB200: ShemsB200, Shems, ShemsCuda, LIB, check

struct DdpgParams                        # must match include/shems_b200.h
    state_size::Cint; action_size::Cint; l1::Cint; l2::Cint; batch::Cint
    gamma::Cfloat; tau::Cfloat; lr_actor::Cfloat; lr_critic::Cfloat
    adam_beta1::Cdouble; adam_beta2::Cdouble; adam_eps::Cdouble
    act_lo::NTuple{2,Cfloat}; act_hi::NTuple{2,Cfloat}
    use_tensor_cores::Cint; population::Cint
end

struct RolloutArgs                       # ShemsRolloutArgs
    policy::Cint; n_steps::Cint; seed::UInt64; env_id_base::Int64
    tape::CUDA.CuPtr{Cfloat}; ep_return::CUDA.CuPtr{Cdouble}; replay::Ptr{Cvoid}
    trace::CUDA.CuPtr{Cdouble}; obs_traj::CUDA.CuPtr{Cfloat}; reward_traj::CUDA.CuPtr{Cfloat}
    tape_unscaled::Cint
end

const _device = parse(Int, get(ENV, "GPU_ID", "0"))
const _rng_mode = get(ENV, "SHEMS_B200_RNG", "julia")      # "julia": the reference's own draws; "philox": device streams

# ------------------------------------------------------------------ replay memory on the device
# memory = CircularBuffer{Any}(MEM_SIZE) (input.jl:140) becomes  memory = DeviceMemory(MEM_SIZE)
mutable struct DeviceMemory
    handle::Ptr{Cvoid}
    capacity::Int
end
function DeviceMemory(capacity::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:replay_create, LIB), Cint, (Int64, Cint, Ref{Ptr{Cvoid}}), capacity, _device, h))
    m = DeviceMemory(h[], capacity)
    finalizer(x -> ccall((:replay_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), m)
    return m
end
Base.length(m::DeviceMemory) = Int(ccall((:replay_length, LIB), Int64, (Ptr{Cvoid},), m.handle))
# push!(memory, [state, action, reward, next_state, done]) — what remember() does (memory_plotting_saving.jl:46-47)
function Base.push!(m::DeviceMemory, t::AbstractVector)
    s, a = CUDA.CuArray(Float32.(vec(t[1]))), CUDA.CuArray(Float32.(vec(t[2])))
    r, s2, d = CUDA.CuArray(Float32[t[3]]), CUDA.CuArray(Float32.(vec(t[4]))), CUDA.CuArray(Float32[t[5]])
    check(ccall((:replay_push, LIB), Cint,
                (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, Int64),
                m.handle, s, a, r, s2, d, 1))
    return m
end
# logical indices (1 = oldest) -> the transitions, as the [s, a, r, s′, done] vectors the reference's memory holds
function fetch_transitions(m::DeviceMemory, idx::AbstractVector{<:Integer})
    n = length(idx)
    s, a, r = CUDA.zeros(Float32, n, 9), CUDA.zeros(Float32, n, 2), CUDA.zeros(Float32, n)        # [9][n], [2][n], [n] in C order
    s2, d = CUDA.zeros(Float32, n, 9), CUDA.zeros(Float32, n)
    check(ccall((:replay_sample, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Int32}, UInt64, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}),
                m.handle, n, Int32.(idx .- 1), 0, s, a, r, s2, d))
    S, A, R, S2, D = Array(s), Array(a), Array(r), Array(s2), Array(d)
    return Any[Any[S[j, :], A[j, :], Float64(R[j]), S2[j, :], D[j] != 0f0] for j in 1:n]
end
# sample(MersenneTwister(rng_dt), memory, batch_size) of getData (memory_plotting_saving.jl:33): the reference's own index stream
StatsBase.sample(rng::AbstractRNG, m::DeviceMemory, n::Integer) = fetch_transitions(m, StatsBase.sample(rng, 1:length(m), n))

_memory() = (memory isa DeviceMemory) ? memory :
    error("DdpgB200: set `memory = DeviceMemory(MEM_SIZE)` in the input file (input.jl:140 builds a CircularBuffer)")

# ------------------------------------------------------------------ the learner
const _learner = let h = Ref{Ptr{Cvoid}}(C_NULL)
    p = DdpgParams(STATE_SIZE, ACTION_SIZE, L1, L2, BATCH_SIZE, γ, τ, η_act, η_crit, 0.9, 0.999, 1e-8,
                   (ACTION_BOUND_LO[1], ACTION_BOUND_LO[2]), (ACTION_BOUND_HI[1], ACTION_BOUND_HI[2]), 0, 1)
    check(ccall((:ddpg_create, LIB), Cint, (Ref{DdpgParams}, Cint, Ref{Ptr{Cvoid}}), p, _device, h))
    check(ccall((:ddpg_init, LIB), Cint, (Ptr{Cvoid}, UInt64), h[], UInt64(rng_run)))       # replaced by the Flux nets' weights on first use
    h[]
end
# BATCH_SIZE = 120, L1/L2 = 250/500 run replay() as two thread-block-cluster kernels by default; `fused_replay(false)` selects the
# one-launch-per-product sequence (returns the resulting state)
fused_replay(on::Bool=true) = ccall((:ddpg_set_fused, LIB), Cint, (Ptr{Cvoid}, Cint), _learner, on ? 1 : 0) == 1

# weights <-> Flux.  Flux.params(chain) = [W1, b1, W2, b2, W3, b3]; a Dense weight is out×in column-major — the library's layout.
function push_net!(net_id::Integer, chain)
    ps = collect(Flux.params(chain))
    for k in 1:3
        check(ccall((:ddpg_set_layer, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}),
                    _learner, net_id, k - 1, Float32.(vec(cpu(ps[2k-1]))), Float32.(vec(cpu(ps[2k])))))
    end
end
function pull_net!(net_id::Integer, chain)
    ps = collect(Flux.params(chain))
    for k in 1:3
        w, b = zeros(Float32, length(ps[2k-1])), zeros(Float32, length(ps[2k]))
        check(ccall((:ddpg_get_layer, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}), _learner, net_id, k - 1, w, b))
        copyto!(ps[2k-1], reshape(w, size(ps[2k-1]))); copyto!(ps[2k], b)
    end
    return chain
end
push_actor!(chain) = push_net!(0, chain)                 # loadBSON -> device (driver :93-101)
pull_actor!(chain) = pull_net!(0, chain)                 # device -> the Chain saveBSON writes (memory_plotting_saving.jl:263-270)

const _adopted = Ref(false)
const _actor_on_device = Ref{Any}(nothing)               # the Flux `actor` object whose weights the device currently holds
function _sync!(; evaluating::Bool)
    if !_adopted[]                                       # first use: the reference's own initial weights (DDPG.jl:21-46), all four nets
        push_net!(0, actor); push_net!(1, critic); push_net!(2, actor_target); push_net!(3, critic_target)
        _adopted[] = true
        _actor_on_device[] = actor
    end
    # `global actor = deepcopy(ac) |> gpu` (driver :95, :101) binds a NEW Chain: an evaluation run uses that one
    if evaluating && _actor_on_device[] !== actor
        push_actor!(actor)
        _actor_on_device[] = actor
    end
    # s_min, s_max are the driver's globals (:30), frozen for the run; normalize() is fused into the kernels
    check(ccall((:ddpg_set_norm, LIB), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}), _learner,
                Float32.(vec(cpu(s_min))), Float32.(vec(cpu(s_max)))))
    return nothing
end

# ------------------------------------------------------------------ populate_memory  (memory_plotting_saving.jl:9-29)
function populate_memory(env::ShemsCuda; rng=0)
    mem = _memory()
    T = EP_LENGTH["train"]
    while length(mem) < MIN_EXP_SIZE
        reset!(env; rng=rng)
        tape = Array{Float32}(undef, env.n_envs, 2, T)                                        # [T][2][N] in C order
        for step in 1:T
            rng2 = parse(Int, string(rng) * string(step))                                     # :15
            tape[1, :, step] = _rng_mode == "julia" ? Float32.(rand(MersenneTwister(rng2), ACTION_SIZE) .* 2 .- 1) : zeros(Float32, 2)   # :18
        end
        dtape = CUDA.CuArray(tape)
        policy, unscaled = _rng_mode == "julia" ? (2, 1) : (1, 0)                             # the taped draws | Philox on the device
        args = RolloutArgs(policy, T, UInt64(abs(rng)), 0, policy == 2 ? pointer(dtape) : CUDA.CU_NULL, CUDA.CU_NULL, mem.handle,
                           CUDA.CU_NULL, CUDA.CU_NULL, CUDA.CU_NULL, unscaled)
        check(ccall((:shems_rollout, LIB), Cint, (Ptr{Cvoid}, Ref{RolloutArgs}), env.handle, args))   # a stored, scale_action(a) applied (:20-23)
        check(ccall((:shems_sync, LIB), Cint, (Ptr{Cvoid},), env.handle))
        rng += 1                                                                              # :26
    end
    ShemsB200.pull!(env)
    return nothing
end

# ------------------------------------------------------------------ act / replay on the device
# act(normalize(s); train) + scale_action (DDPG.jl:148-184) on the RAW state with the noise vector the reference would draw
function act_b200(s, noise::AbstractVector)
    obs, nz = CUDA.CuArray(Float32.(vec(s))), CUDA.CuArray(Float32.(vec(noise)))
    a, scaled = CUDA.zeros(Float32, ACTION_SIZE), CUDA.zeros(Float32, ACTION_SIZE)
    check(ccall((:ddpg_act, LIB), Cint,
                (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, Int64, Cfloat, UInt64, Int64, Int64, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}),
                _learner, obs, 1, 0f0, 0, 0, 0, nz, a, scaled))
    return Array(a), Array(scaled)
end
# replay(; rng_rpl) (DDPG.jl:121-145): one whole update on the device, on the minibatch getData(BATCH_SIZE, rng_dt = rng_rpl) draws
function replay_b200(; rng_rpl=0)
    mem = _memory()
    if _rng_mode == "julia"
        idx = Int32.(StatsBase.sample(MersenneTwister(rng_rpl), 1:length(mem), BATCH_SIZE) .- 1)
        check(ccall((:ddpg_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, UInt64), _learner, mem.handle, 1, idx, 0))
    else
        check(ccall((:ddpg_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, UInt64), _learner, mem.handle, 1, C_NULL, UInt64(abs(rng_rpl))))
    end
    return nothing
end

# ------------------------------------------------------------------ episode!  (DDPG.jl:186-242)
function episode!(env::ShemsCuda; NUM_STEPS=EP_LENGTH["train"], train=true, render=false, track=0, rng_ep=0)
    _sync!(evaluating=!train)
    reset!(env; rng=rng_ep)
    if track < 0                                         # rule-based benchmark (:209-212): the fused rollout with the 23-column trace
        ret, trace = CUDA.zeros(Float64, env.n_envs), CUDA.zeros(Float64, env.n_envs, 23, NUM_STEPS)
        args = RolloutArgs(0, NUM_STEPS, 0, 0, CUDA.CU_NULL, pointer(ret), C_NULL, pointer(trace), CUDA.CU_NULL, CUDA.CU_NULL, 0)
        check(ccall((:shems_rollout, LIB), Cint, (Ptr{Cvoid}, Ref{RolloutArgs}), env.handle, args))
        ShemsB200.pull!(env)
        return Array(ret)[1], permutedims(Array(trace)[1, :, :])                              # (reward_eps, results [NUM_STEPS x 23])
    end
    if !train                                            # evaluation / DRL inference: ONE kernel for the whole episode (ddpg_rollout)
        ret = CUDA.zeros(Float64, env.n_envs)
        trace = track == 0 ? nothing : CUDA.zeros(Float64, env.n_envs, 23, NUM_STEPS)
        check(ccall((:ddpg_rollout, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cfloat, UInt64, Int64, CUDA.CuPtr{Cdouble}, CUDA.CuPtr{Cdouble}, CUDA.CuPtr{Cfloat}),
                    _learner, env.handle, NUM_STEPS, 0f0, 0, 0, ret, trace === nothing ? CUDA.CU_NULL : pointer(trace), CUDA.CU_NULL))
        ShemsB200.pull!(env)
        track == 0 && return Array(ret)[1], NUM_STEPS, 0f0
        return Array(ret)[1], permutedims(Array(trace)[1, :, :])
    end
    mem = _memory()
    if _rng_mode != "julia" && noise_type == "gn"
        # the whole training loop as ONE native call (ddpg_episode): Philox noise and minibatch draws, no host round trip per step
        ret, nz = CUDA.zeros(Float64, env.n_envs), CUDA.zeros(Float32, env.n_envs)
        mems = Ptr{Cvoid}[mem.handle]
        check(ccall((:ddpg_episode, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cfloat, UInt64, Cint, Int64, CUDA.CuPtr{Cdouble}, CUDA.CuPtr{Cfloat}),
                    _learner, env.handle, mems, NUM_STEPS, 1, gn.σ_act, UInt64(abs(rng_ep)), 1, 0, ret, nz))
        ShemsB200.pull!(env)
        return Array(ret)[1], NUM_STEPS, Array(nz)[1]
    end
    # the reference's loop, statement for statement, with its own random draws; every heavy step is one call into the library
    reward_eps, noise_eps, last_step = 0f0, 0f0, 1
    for step = 1:NUM_STEPS
        rng_step = parse(Int, string(rng_ep) * string(step))                                  # :197
        s = copy(env.state)
        noise = noise_type == "ou" ? sample_noise(ou, rng_step + 1) : sample_noise(gn, rng_step + 1)   # act(): rng_act + i, i = 1 (:157-160); DDPG.jl's own
        a, scaled_action = act_b200(s, noise)                                                 # clamp(actor(normalize(s)) + noise, -1, 1), scale_action
        r, s′ = step!(env, s, scaled_action)                                                  # :205
        reward_eps += r
        noise_eps += sum(noise) / length(noise)                                               # mean(noise) (:172, :224)
        last_step = step
        push!(mem, Any[s, a, r, s′, finished(env, s′)])                                       # remember(...) :229 (the UNSCALED action)
        replay_b200(rng_rpl=rng_step)                                                         # :231
        finished(env, s′) && break
    end
    return reward_eps, last_step, noise_eps
end

# ------------------------------------------------------------------ run_episodes  (DDPG.jl:244-298)
function run_episodes(env_train::ShemsCuda, env_eval::ShemsCuda, total_reward, score_mean, best_run, noise_mean, test_every, render, rng; track=0)
    best_score = -100000
    for i = 1:NUM_EP
        score = 0f0
        score_all = 0f0
        global current_episode = i
        rng_ep = parse(Int, string(rng) * string(i))                                          # :252
        total_reward[i], last_step, noise_mean[i] = episode!(env_train, train=true, render=render, track=track, rng_ep=rng_ep)
        if i % test_every == 1                                                                # :261
            idx = ceil(Int32, i / test_every)
            for test_ep in 1:test_runs
                rng_test = parse(Int, string(seed_ini) * string(test_ep))
                score, noise = episode!(env_eval, train=false, render=false, NUM_STEPS=EP_LENGTH["train"], track=track, rng_ep=rng_test)
                score_all += score
            end
            score_mean[idx] = score_all / test_runs
            if score_mean[idx] > best_score                                                   # :282-289 early-stopping checkpoint
                pull_actor!(actor)
                saveBSON(actor, total_reward, score_mean, best_run, noise_mean, idx=i, path="temp", rng=rng_run)
                best_score = score_mean[idx]
                global best_run = i
            end
        end
    end
    pull_actor!(actor)                                   # the driver saves `actor` right after this call (:45)
    return nothing
end

# ------------------------------------------------------------------ what the reference cannot do: resume
# the whole learner (nets, targets, ADAM moments, β powers, update counter, s_min/s_max) as one binary blob (ddpg_get_state)
function save_learner_state(path::AbstractString)
    n = ccall((:ddpg_state_floats, LIB), Int64, (Ptr{Cvoid},), _learner)
    st, opt = zeros(Float32, n), zeros(Float64, 8)
    check(ccall((:ddpg_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cdouble}), _learner, st, opt))
    open(io -> (write(io, Int64(n)); write(io, opt); write(io, st)), path, "w")
end
function load_learner_state(path::AbstractString)
    open(path, "r") do io
        n = read(io, Int64)
        opt = read!(io, zeros(Float64, 8)); st = read!(io, zeros(Float32, n))
        check(ccall((:ddpg_set_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cdouble}), _learner, st, opt))
    end
    _adopted[] = true
    return nothing
end
