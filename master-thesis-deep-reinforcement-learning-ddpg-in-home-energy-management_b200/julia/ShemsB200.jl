# ShemsB200.jl — thin Julia shim over libshems_b200.so (include/shems_b200.h).
#
# Drop-in for RL-SHEMS/RL_environments/envs/shems_LU1.jl: it exports the same names
# (Shems, reset!, step!, action, finished, state, track) and subtypes Reinforce.AbstractEnvironment,
# so algorithms/DDPG.jl (episode!, populate_memory, inference) keeps calling
#     reset!(env; rng), step!(env, s, a; track), action(env, track), finished(env, s′),
#     env.state, env.a, env.reward, env.step, env.idx, env.maxsteps, env.path
# unchanged.  All arithmetic happens in CUDA kernels behind the C ABI; this file only marshals.
#
# `Shems` is an ABSTRACT type here and `Shems(maxsteps, path)` returns the concrete `ShemsCuda <: Shems`: every reference method
# written for `env::Shems` (episode!, run_episodes, populate_memory, inference) still applies, and DdpgB200.jl overrides the hot
# ones with methods for `env::ShemsCuda`, which win by dispatch whatever the order in which the driver includes the files.
#
# NOTE: Julia is not installed in the build/CI image of this repository, so this file is reviewed but
# never executed there; every code path it calls is exercised through the identical C ABI by the
# Python ctypes harness (tests/test_env_gpu.py).  It is written against Julia 1.6 / CUDA.jl 2.6 like
# the reference (RL-SHEMS/Manifest.toml).
module ShemsB200

using Reinforce: AbstractEnvironment
import Reinforce: reset!, action, finished, step!, state, actions
using CSV, DataFrames

export Shems, ShemsCuda, reset!, step!, action, finished, state, actions, track, ShemsState, ShemsAction

const LIB = get(ENV, "SHEMS_B200_LIB", "libshems_b200")

# ---------------------------------------------------------------- ABI structs (must match include/shems_b200.h)
struct ShemsParams
    pv_eta::Cfloat; b_eta::Cfloat; b_soc_min::Cfloat; b_soc_max::Cfloat
    b_rate_max::Cdouble; b_loss::Cfloat; ev_soc_min::Cfloat; ev_soc_max::Cfloat
    ev_rate_max::Cfloat; penalty_weight::Cfloat; sell_discount::Cdouble
    discomfort_weight_ev::Cdouble; disc_pot::Cdouble
    penalty_weight_f64::Cdouble; penalty_in_f64::Cint; reward_form::Cint   # sibling envs (shems_LU7 / shems_LU1_input0607), zero for LU1
end

last_error() = unsafe_string(ccall((:shems_last_error, LIB), Cstring, ()))

# status -> the exception the reference would have raised
function check(st::Integer)
    st == 0 && return nothing
    msg = last_error()
    st == -3 && throw(BoundsError(msg))          # next_state! row idx+1 > nrow (shems_LU1.jl:266-268)
    st == -4 && throw(KeyError(msg))             # capacities[charger_id]       (shems_LU1.jl:95)
    error("libshems_b200 [$st]: $msg")
end

# ---------------------------------------------------------------- module-level constants (shems_LU1.jl:17, 45-59)
const Job_ID = ENV["JOB_ID"]
const charger_id = ((parse(Int, Job_ID) ÷ 100) % 100)
const params = let p = Ref{ShemsParams}()
    check(ccall((:shems_params_for_charger, LIB), Cint, (Cint, Ref{ShemsParams}), charger_id, p))
    p[]
end

# ---------------------------------------------------------------- state / action views (shems_LU1.jl:101-167)
mutable struct ShemsState{T<:AbstractFloat} <: AbstractVector{T}
    Soc_b::T; Soc_ev::T; c_ev::T; d_e::T; g_e::T; p_buy::T; h_cos::T; h_sin::T; season::T
end
ShemsState() = ShemsState(0f0, 0f0, -1f0, 0f0, 0f0, 0f0, 1f0, 0f0, 1f0)
ShemsState(v::AbstractVector) = ShemsState(Float32.(v)...)
Base.size(::ShemsState) = (9,)
Base.getindex(s::ShemsState, i::Int) = getfield(s, i)

mutable struct ShemsAction{T<:AbstractFloat} <: AbstractVector{T}
    B::T; EV::T
end
ShemsAction() = ShemsAction(0.7f0, 1f0)
Base.size(::ShemsAction) = (2,)
Base.minimum(::ShemsAction) = (0f0, 0f0)
Base.maximum(::ShemsAction) = (1f0, 1f0)
Base.getindex(a::ShemsAction, i::Int) = getfield(a, i)

# ---------------------------------------------------------------- the environment (shems_LU1.jl:169-203)
abstract type Shems <: AbstractEnvironment end
mutable struct ShemsCuda <: Shems
    handle::Ptr{Cvoid}
    state::ShemsState{Float32}
    reward::Float64
    a::ShemsAction{Float32}
    step::Int
    maxsteps::Int
    idx::Int
    path::String
    n_envs::Int
end

const SERIES_COLS = (:soc_ev, :h_countdown, :electkwh, :PV_generation, :p_buy, :hour_cos, :hour_sin, :season)

# Shems(maxsteps, path): the CSV is parsed ONCE here (the reference re-parses it on every reset/step, :217, :265)
Shems(maxsteps, path; kw...) = ShemsCuda(maxsteps, path; kw...)
function ShemsCuda(maxsteps, path; n_envs::Int=1, device::Int=parse(Int, get(ENV, "GPU_ID", "0")))
    df = CSV.read(path, DataFrame)
    nrows = nrow(df)
    series = Matrix{Float32}(undef, nrows, 8)            # column-major: [8][nrows] in C order
    for (k, c) in enumerate(SERIES_COLS)
        series[:, k] = Float32.(df[!, c])
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:shems_create, LIB), Cint,
                (Ref{ShemsParams}, Ptr{Cfloat}, Cint, Cint, Int64, Cint, Ref{Ptr{Cvoid}}),
                params, series, nrows, maxsteps, n_envs, device, h))
    env = ShemsCuda(h[], ShemsState(), 0.0, ShemsAction(), 0, maxsteps, 1, path, n_envs)
    finalizer(e -> ccall((:shems_destroy, LIB), Cint, (Ptr{Cvoid},), e.handle), env)
    return env
end

# copy instance 1 of the device state into the Julia-visible fields (n_envs == 1 is the reference's use)
function pull!(env::ShemsCuda)
    obs = Matrix{Float32}(undef, env.n_envs, 9)
    idx = Vector{Int32}(undef, env.n_envs)
    check(ccall((:shems_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Int32}), env.handle, obs, idx))
    env.state = ShemsState(obs[1, :])
    env.idx = idx[1]
    st = Ref{Cint}(0)
    check(ccall((:shems_get_step, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}), env.handle, st))
    env.step = st[]
    return obs, idx
end

# reset!(env; rng) (shems_LU1.jl:206-262).  rng == -1: deterministic start.  Otherwise the two draws the
# reference takes from MersenneTwister(rng) (:224-225) are made HERE, in Julia, and handed to the library
# (mode 1 = SHEMS_RESET_HOST_DRAWS), so the start row and Soc_b are bit-identical to the reference run.
function reset!(env::ShemsCuda; rng=0)
    if rng == -1
        check(ccall((:shems_reset, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Cfloat}, UInt64, Int64),
                    env.handle, 0, C_NULL, C_NULL, 0, 0))
    else
        nrows = Int(ccall((:shems_num_rows, LIB), Cint, (Ptr{Cvoid},), env.handle))
        socb0 = Float32[rand(MersenneTwister(rng + i - 1), Uniform(params.b_soc_min, params.b_soc_max)) for i in 1:env.n_envs]
        idx0 = Int32[rand(MersenneTwister(rng + i - 1), 1:(nrows - env.maxsteps)) for i in 1:env.n_envs]
        check(ccall((:shems_reset, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Cfloat}, UInt64, Int64),
                    env.handle, 1, idx0, socb0, 0, 0))
    end
    env.reward = 0.0
    env.a = ShemsAction()
    pull!(env)
    return env
end

# step!(env, s, a; track=0) (shems_LU1.jl:343-485): a = targets (track >= 0) or (B, EV) (track < 0)
function step!(env::ShemsCuda, s, a; track=0)
    n = env.n_envs
    act = CUDA.CuArray(reshape(Float32.(collect(a)), n, 2))          # [2][N] in C order
    rew = CUDA.zeros(Float64, n)                                     # env.reward::Float64 (shems_LU1.jl:171)
    trace = track == 0 ? nothing : CUDA.zeros(Float64, n, 23)
    check(ccall((:shems_step, LIB), Cint,
                (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, Cint, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cdouble}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cdouble}),
                env.handle, act, track < 0 ? -1 : (track > 0 ? 1 : 0), CUDA.CU_NULL, rew, CUDA.CU_NULL,
                trace === nothing ? CUDA.CU_NULL : trace))
    check(ccall((:shems_sync, LIB), Cint, (Ptr{Cvoid},), env.handle))
    pull!(env)
    if track >= 0
        env.a = ShemsAction(Float32(a[1]), Float32(a[2]))
    else
        env.a = ShemsAction(0f0, 0f0)
    end
    env.reward = Array(rew)[1]                                       # the Float64 reward of the reference (:467-470)
    if track == 0
        return env.reward, Vector{Float32}(env.state)
    else
        return env.reward, Vector{Float32}(env.state), Matrix{Float64}(Array(trace)[1:1, :])
    end
end

# action(env, a::ShemsAction) (:283-316) and action(env, track) (:318-340)
function action(env::ShemsCuda, a::ShemsAction)
    tgt = CUDA.CuArray(Float32[a.B, a.EV]); out = CUDA.zeros(Float32, 2 * env.n_envs)
    check(ccall((:shems_action_drl, LIB), Cint, (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, CUDA.CuPtr{Cfloat}), env.handle, tgt, out))
    return Array(out)[1:env.n_envs:end]
end
function action(env::ShemsCuda, track=-1)
    out = CUDA.zeros(Float32, 2 * env.n_envs)
    check(ccall((:shems_action_rule, LIB), Cint, (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}), env.handle, out))
    return Array(out)[1:env.n_envs:end]
end

finished(env::Shems, s′) = false                                  # shems_LU1.jl:487-502
state(env::Shems) = env.state
actions(env::Shems, s) = (minimum(env.a), maximum(env.a))
track = 0

import CUDA
using Random: MersenneTwister
using Distributions: Uniform

end # module
