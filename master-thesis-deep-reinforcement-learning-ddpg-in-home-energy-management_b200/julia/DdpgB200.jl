# DdpgB200.jl — thin Julia shim over the DDPG half of libshems_b200.so (include/shems_b200.h).
#
# Drop-in for the functions of RL-SHEMS/algorithms/DDPG.jl and RL-SHEMS/src/memory_plotting_saving.jl:1-57 that sit on the hot
# path.  `include` it from the input file INSTEAD of algorithms/DDPG.jl (DDPG_reinforce_charger_v1.jl:24) after ShemsB200.jl; the
# driver's calls keep their names and argument meaning:
#     populate_memory(env; rng)            memory_plotting_saving.jl:9-29
#     remember(s, a, r, s′, done)          :46-47
#     min_max_buffer(n; rng_mm)            :50-53   (also freezes s_min / s_max inside the learner, driver :30)
#     act(s; train, rng_act)               DDPG.jl:148-176 — takes the RAW state: normalize() is fused into the kernel
#     scale_action(a)                      :178-184
#     replay(; rng_rpl)                    :121-145
#     episode!(env; NUM_STEPS, train, render, track, rng_ep)   :186-242 (track == 0: one native ddpg_episode call)
#     pull_actor!(actor)                   copies the learner's actor into a Flux Chain so saveBSON (:263-270) is unchanged
# Globals read from input.jl exactly like the reference: BATCH_SIZE, MEM_SIZE, L1, L2, γ, τ, η_act, η_crit, EP_LENGTH, noise_type,
# gn / ou, ACTION_BOUND_LO / HI, rng_run.
#
# NOTE: Julia is not installed in the build/CI image of this repository: this file is reviewed, never executed there; every entry
# point it calls is exercised through the identical C ABI by the Python ctypes harness (tests/test_replay_ddpg_gpu.py).
# RNG: Julia's MersenneTwister streams stay on the Julia side where a draw is cheap (the two reset draws, see ShemsB200.reset!);
# minibatch indices, warm-up actions and exploration noise come from the library's Philox streams keyed by the same integer seeds.

import CUDA
using .ShemsB200: ShemsB200, Shems, LIB, check

struct DdpgParams                        # must match include/shems_b200.h
    state_size::Cint; action_size::Cint; l1::Cint; l2::Cint; batch::Cint
    gamma::Cfloat; tau::Cfloat; lr_actor::Cfloat; lr_critic::Cfloat
    adam_beta1::Cdouble; adam_beta2::Cdouble; adam_eps::Cdouble
    act_lo::NTuple{2,Cfloat}; act_hi::NTuple{2,Cfloat}
    use_tensor_cores::Cint; population::Cint
end

const _device = parse(Int, get(ENV, "GPU_ID", "0"))
const _learner = let h = Ref{Ptr{Cvoid}}(C_NULL)
    p = DdpgParams(STATE_SIZE, ACTION_SIZE, L1, L2, BATCH_SIZE, γ, τ, η_act, η_crit, 0.9, 0.999, 1e-8,
                   (ACTION_BOUND_LO[1], ACTION_BOUND_LO[2]), (ACTION_BOUND_HI[1], ACTION_BOUND_HI[2]), 0, 1)
    check(ccall((:ddpg_create, LIB), Cint, (Ref{DdpgParams}, Cint, Ref{Ptr{Cvoid}}), p, _device, h))
    check(ccall((:ddpg_init, LIB), Cint, (Ptr{Cvoid}, UInt64), h[], UInt64(rng_run)))       # glorot / ±3e-3 init, DDPG.jl:21-22
    h[]
end
const _memory = let h = Ref{Ptr{Cvoid}}(C_NULL)                                              # memory = CircularBuffer{Any}(MEM_SIZE), input.jl:140
    check(ccall((:replay_create, LIB), Cint, (Int64, Cint, Ref{Ptr{Cvoid}}), MEM_SIZE, _device, h))
    h[]
end
memory_length() = Int(ccall((:replay_length, LIB), Int64, (Ptr{Cvoid},), _memory))
# BATCH_SIZE = 120, L1/L2 = 250/500 run replay() as two thread-block-cluster kernels by default; `fused_replay(false)` selects the
# one-launch-per-product sequence (returns the resulting state)
fused_replay(on::Bool=true) = ccall((:ddpg_set_fused, LIB), Cint, (Ptr{Cvoid}, Cint), _learner, on ? 1 : 0) == 1

# ------------------------------------------------------------------ replay memory
function remember(state, action, reward, next_state, done)                                   # memory_plotting_saving.jl:46-47
    s, a = CUDA.CuArray(Float32.(vec(state))), CUDA.CuArray(Float32.(vec(action)))
    r, s2, d = CUDA.CuArray(Float32[reward]), CUDA.CuArray(Float32.(vec(next_state))), CUDA.CuArray(Float32[done])
    check(ccall((:replay_push, LIB), Cint,
                (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, Int64),
                _memory, s, a, r, s2, d, 1))
end

struct RolloutArgs                       # ShemsRolloutArgs
    policy::Cint; n_steps::Cint; seed::UInt64; env_id_base::Int64
    tape::CUDA.CuPtr{Cfloat}; ep_return::CUDA.CuPtr{Cdouble}; replay::Ptr{Cvoid}
    trace::CUDA.CuPtr{Cdouble}; obs_traj::CUDA.CuPtr{Cfloat}; reward_traj::CUDA.CuPtr{Cfloat}
end

function populate_memory(env::Shems; rng=0)                                                  # memory_plotting_saving.jl:9-29
    while memory_length() < MEM_SIZE
        reset!(env; rng=rng)
        args = RolloutArgs(1, EP_LENGTH["train"], UInt64(rng), 0, CUDA.CU_NULL, CUDA.CU_NULL, _memory, CUDA.CU_NULL, CUDA.CU_NULL, CUDA.CU_NULL)
        check(ccall((:shems_rollout, LIB), Cint, (Ptr{Cvoid}, Ref{RolloutArgs}), env.handle, args))   # a = 2U-1 stored, scaled to [0,1]² for the env
        rng += 1                                                                              # `rng += 1` per episode (:26)
    end
    return nothing
end

function min_max_buffer(n; rng_mm=0)                                                          # memory_plotting_saving.jl:50-53
    s_min, s_max = zeros(Float32, STATE_SIZE), zeros(Float32, STATE_SIZE)
    check(ccall((:replay_minmax, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, UInt64, Ptr{Cfloat}, Ptr{Cfloat}),
                _memory, n, C_NULL, UInt64(rng_mm), s_min, s_max))
    check(ccall((:ddpg_set_norm, LIB), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}), _learner, s_min, s_max))
    return reshape(s_min, :, 1), reshape(s_max, :, 1)
end

# ------------------------------------------------------------------ act / replay
const _ou_x = CUDA.zeros(Float32, ACTION_SIZE)                                                # OUNoise.X (input.jl:234), never reset

function act(s; train=true, rng_act=0)                                                        # DDPG.jl:148-176 on the raw state
    obs = CUDA.CuArray(Float32.(vec(s)))
    a, scaled = CUDA.zeros(Float32, ACTION_SIZE), CUDA.zeros(Float32, ACTION_SIZE)
    if train && noise_type == "ou"
        check(ccall((:ddpg_act_ou, LIB), Cint,
                    (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, Int64, Cfloat, Cfloat, Cfloat, Cfloat, CUDA.CuPtr{Cfloat}, UInt64, Int64, Int64,
                     CUDA.CuPtr{Cdouble}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}),
                    _learner, obs, 1, ou.θ, ou.μ, ou.σ, ou.dt, _ou_x, UInt64(rng_act), 0, 0, CUDA.CU_NULL, a, scaled))
    else
        σ_now = (train && noise_type == "gn") ? gn.σ_act : 0f0
        check(ccall((:ddpg_act, LIB), Cint,
                    (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, Int64, Cfloat, UInt64, Int64, Int64, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}),
                    _learner, obs, 1, σ_now, UInt64(rng_act), 0, 0, CUDA.CU_NULL, a, scaled))
    end
    return Array(a), Array(scaled)                                                            # (action in [-1,1]², scale_action(action))
end

scale_action(action) = Float32.(ACTION_BOUND_LO .+ (action .+ ones(ACTION_SIZE)) .* 0.5 .* (ACTION_BOUND_HI .- ACTION_BOUND_LO))   # :178-184

function replay(; rng_rpl=0)                                                                  # DDPG.jl:121-145, one whole update on the device
    check(ccall((:ddpg_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, UInt64), _learner, _memory, 1, C_NULL, UInt64(rng_rpl)))
end

# ------------------------------------------------------------------ episode!  (DDPG.jl:186-242)
function episode!(env::Shems; NUM_STEPS=EP_LENGTH["train"], train=true, render=false, track=0, rng_ep=0)
    reset!(env; rng=rng_ep)
    if track == 0 && noise_type == "gn"
        # the whole loop below as ONE native call (ddpg_episode): no host round trip per step, same per-step seeds
        ret = CUDA.zeros(Float64, env.n_envs)
        mems = Ptr{Cvoid}[_memory]
        check(ccall((:ddpg_episode, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cfloat, UInt64, Cint, Int64, CUDA.CuPtr{Cdouble}),
                    _learner, env.handle, train ? mems : C_NULL, NUM_STEPS, train ? 1 : 0, gn.σ_act, UInt64(abs(rng_ep)), 1, 0, ret))
        ShemsB200.pull!(env)                          # refresh env.state / env.idx / env.step mirrors from the device
        return Array(ret)[1], NUM_STEPS, 0f0
    end
    reward_eps, noise_eps, last_step = 0.0, 0f0, 1
    results = Matrix{Float64}(undef, 0, 23)
    for step = 1:NUM_STEPS
        rng_step = parse(Int, string(abs(rng_ep)) * string(step))                             # :197
        s = copy(env.state)
        if track < 0
            a = action(env, track)                                                            # rule-based controller, :209-212
            r, s′, row = step!(env, s, a; track=track)
            results = vcat(results, row)
        else
            a, scaled = act(s; train=train, rng_act=rng_step)
            if track == 0
                r, s′ = step!(env, s, scaled)
            else
                r, s′, row = step!(env, s, scaled; track=track)
                results = vcat(results, row)
            end
        end
        reward_eps += r
        last_step = step
        if train
            remember(s, a, r, s′, finished(env, s′))                                          # the UNSCALED action is stored (:229)
            replay(rng_rpl=rng_step)                                                          # :231
        end
        finished(env, s′) && break
    end
    return track == 0 ? (reward_eps, last_step, noise_eps) : (reward_eps, results)
end

# ------------------------------------------------------------------ weights <-> Flux (checkpoints stay BSON files written by the driver)
function pull_actor!(chain)                                                                   # chain = Chain(Dense, Dense, Dense) on the CPU
    for (k, layer) in enumerate(chain.layers)
        w, b = zeros(Float32, length(layer.W)), zeros(Float32, length(layer.b))
        check(ccall((:ddpg_get_layer, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}), _learner, 0, k - 1, w, b))
        layer.W .= reshape(w, size(layer.W)); layer.b .= b                                    # same layout as Dense.W (out×in, column-major)
    end
    return chain
end
function push_actor!(chain)                                                                   # loadBSON -> library (inference runs, driver :93-101)
    for (k, layer) in enumerate(chain.layers)
        check(ccall((:ddpg_set_layer, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}),
                    _learner, 0, k - 1, Float32.(vec(layer.W)), Float32.(vec(layer.b))))
    end
end
