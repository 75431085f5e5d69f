# InteriorPointDDPB200.jl -- Julia binding of libipddp_b200.so (C ABI: include/ipddp_b200.h).
#
# NOT EXECUTED in the build image (no Julia there; tests/test_abi.py checks every ccall signature and struct mirror of this
# file against include/ipddp_b200.h).  It keeps the exported names of reference src/InteriorPointDDP.jl:29-45
# (Dynamics / Objective / Constraint / Bound / Solver / Options / solve! / get_trajectory) on top of a batched layer
# (BatchProblem, solve_queue!).  Model code generation: the reference's `Symbolics.build_function` call sites
# (src/dynamics.jl:26-34, src/objectives.jl:23-28, src/constraints.jl:25-39) are retargeted to
# `target = Symbolics.CTarget()` in codegen.jl; the emitted C bodies are wrapped into the `Model_<name>` struct layout of
# interiorpointddp.jl_b200/csrc/models_gen/*.cuh and compiled with nvcc into a plugin (see INTEGRATION.md).
module InteriorPointDDPB200

# the reference's exported names (src/InteriorPointDDP.jl:29-45) ...
export Objective, Constraint, Dynamics, Bound, Solver, Options, solve!, get_trajectory
# ... and the batched layer underneath them
export BatchProblem, Stats, Queue, solve_queue!, solve_many!, get_duals, get_trace, get_stats, set_tuning!, set_stream!,
       load_model!

const LIB = get(ENV, "IPDDP_B200_LIB", "libipddp_b200.so")

include("codegen.jl")

"""
    Options{T}(; kwargs...)

The reference's `Options{T}` (src/options.jl:1-38): same field names, types and defaults.
"""
Base.@kwdef mutable struct Options{T}
    quasi_newton::Bool = false
    optimality_tolerance::T = 1.0e-8
    max_iterations::Int = 1000
    reset_cache::Bool = true
    verbose = false
    print_frequency = 10
    μ_init::T = 1.0
    ineq_dual_init::T = 1.0
    κ_1::T = 0.01
    κ_2::T = 0.01
    reg_1::T = 1e-4
    reg_min::T = 1e-20
    reg_max::T = 1e40
    κ_̄w_p::T = 100.0
    κ_w_p::T = 8.0
    κ_w_m::T = 1.0 / 3.0
    κ_c::T = 0.25
    δ_c::T = 1e-8
    κ_ϵ::T = 10.0
    κ_μ::T = 0.2
    θ_μ::T = 1.2
    τ_min::T = 0.99
    s_max::T = 100.0
    η_L::T = 1e-4
    s_L::T = 2.3
    δ::T = 1.0
    s_θ::T = 1.1
    γ_α::T = 0.05
    γ_θ::T = 1e-5
    γ_L::T = 1e-5
    κ_Σ::T = 1e10
end

# C mirror of `ipddp_options` (include/ipddp_b200.h), same field order as Options{T}; passed by reference through ccall
mutable struct COptions
    quasi_newton::Cint
    optimality_tolerance::Cdouble
    max_iterations::Cint
    reset_cache::Cint
    verbose::Cint
    print_frequency::Cint
    μ_init::Cdouble
    ineq_dual_init::Cdouble
    κ_1::Cdouble
    κ_2::Cdouble
    reg_1::Cdouble
    reg_min::Cdouble
    reg_max::Cdouble
    κ_̄w_p::Cdouble
    κ_w_p::Cdouble
    κ_w_m::Cdouble
    κ_c::Cdouble
    δ_c::Cdouble
    κ_ϵ::Cdouble
    κ_μ::Cdouble
    θ_μ::Cdouble
    τ_min::Cdouble
    s_max::Cdouble
    η_L::Cdouble
    s_L::Cdouble
    δ::Cdouble
    s_θ::Cdouble
    γ_α::Cdouble
    γ_θ::Cdouble
    γ_L::Cdouble
    κ_Σ::Cdouble
end
COptions(o::Options) = COptions((getfield(o, f) for f in fieldnames(Options))...)   # Bool / Int / T convert field by field

last_error() = unsafe_string(ccall((:ipddp_last_error, LIB), Cstring, ()))
check(rc, what) = rc == 0 || error("$what failed: $(last_error())")

load_model!(plugin_path::AbstractString) =
    check(ccall((:ipddp_model_load, LIB), Cint, (Cstring,), plugin_path), "ipddp_model_load")

mutable struct BatchProblem
    handle::Ptr{Cvoid}
    model::String
    B::Int
    N::Int
    nx::Int
    nu::Int
    nc::Int
    np::Int
    ns::Int          # stride of the state outputs: the largest state of any knot (= nx for a plain model)
    nstage::Int      # stage types of the model (1 for a plain model)
    # per-instance mirror of the reference's SolverData (src/data/solver.jl:8-33), filled by solve!
    status::Vector{Cint}
    k::Vector{Cint}
    j::Vector{Cint}
    l::Vector{Cint}
    objective::Vector{Cdouble}
    primal_inf::Vector{Cdouble}
    dual_inf::Vector{Cdouble}
    cs_inf::Vector{Cdouble}
    μ::Vector{Cdouble}
    reg_last::Vector{Cdouble}
    step_size::Vector{Cdouble}
end

# mirror of `ipddp_stats` (timers of src/data/solver.jl:16-18 split per kernel, work counters)
Base.@kwdef mutable struct Stats
    iterations::Cint = 0
    launches::Clonglong = 0
    ms_total::Cdouble = 0.0
    ms_init::Cdouble = 0.0
    ms_derivs::Cdouble = 0.0
    ms_backward::Cdouble = 0.0
    ms_check::Cdouble = 0.0
    ms_forward::Cdouble = 0.0
    sum_backward::Clonglong = 0
    sum_sweeps::Clonglong = 0
    sum_kkt::Clonglong = 0
    sum_rollouts::Clonglong = 0
    sum_deriv_stages::Clonglong = 0
    n_converged::Clonglong = 0
    n_active_rounds::Clonglong = 0
    sum_active_sq::Cdouble = 0.0
end

"""
    BatchProblem(model, B, N; options=Options{Float64}(), device=0, indices_compl=Cint[])

Batched counterpart of `Solver(T, dynamics, objectives, constraints, bounds; options)` (reference src/solver.jl:11-26).
"""
function BatchProblem(model::String, B::Int, N::Int; options::Options=Options{Float64}(), device::Int=0,
                      indices_compl::Vector{Cint}=Cint[], trace_capacity::Int=0)
    nx = Ref{Cint}(0); nu = Ref{Cint}(0); nc = Ref{Cint}(0); np = Ref{Cint}(0); slots = Ref{Cint}(0)
    check(ccall((:ipddp_model_dims, LIB), Cint, (Cstring, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}),
                model, nx, nu, nc, np, slots), "ipddp_model_dims")      # (ccall takes no splatted arguments)
    nst = Ref{Cint}(0); nxt = Ref{Cint}(0)
    snx = zeros(Cint, 4); snu = zeros(Cint, 4); snc = zeros(Cint, 4); snxn = zeros(Cint, 4)
    check(ccall((:ipddp_model_stages, LIB), Cint,
                (Cstring, Ref{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ref{Cint}), model, nst, snx, snu, snc, snxn, nxt),
          "ipddp_model_stages")
    ns = max(nxt[], maximum(snx[1:nst[]]), maximum(snxn[1:nst[]]))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ipddp_problem_create, LIB), Cint,
                (Cstring, Cint, Cint, Ptr{Cint}, Cint, Ref{COptions}, Cint, Cint, Ref{Ptr{Cvoid}}),
                model, B, N, isempty(indices_compl) ? C_NULL : indices_compl, length(indices_compl),
                Ref(COptions(options)), device, trace_capacity, h), "ipddp_problem_create")
    p = BatchProblem(h[], model, B, N, nx[], nu[], nc[], np[], Int(ns), Int(nst[]), zeros(Cint, B), zeros(Cint, B),
                     zeros(Cint, B), zeros(Cint, B), zeros(B), zeros(B), zeros(B), zeros(B), zeros(B), zeros(B), zeros(B))
    finalizer(q -> ccall((:ipddp_problem_destroy, LIB), Cint, (Ptr{Cvoid},), q.handle), p)
    return p
end

"""
    solve!(prob, x1, controls; params, lower, upper, horizons)

Batched `solve!(solver, x1, controls)` (reference src/solve.jl:1-4).  `x1` is nx x B, `controls` is nu x (N-1) x B
(column-major, i.e. instance-major in memory as the C ABI expects), `params` np x B, `lower`/`upper` nu x B.
"""
function solve!(p::BatchProblem, x1::Matrix{Float64}, controls::Array{Float64,3}; params=nothing,
                lower::Matrix{Float64}, upper::Matrix{Float64}, horizons=nothing)
    size(x1, 2) == p.B && size(controls, 3) == p.B || error("x1 must be nx x B and controls nu x (N-1) x B with B = $(p.B)")
    hz = horizons === nothing ? nothing : convert(Vector{Cint}, horizons)
    check(ccall((:ipddp_set_inputs, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                p.handle, x1, controls, params === nothing ? C_NULL : params, lower, upper,
                hz === nothing ? C_NULL : hz), "ipddp_set_inputs")
    check(ccall((:ipddp_solve, LIB), Cint, (Ptr{Cvoid}, Cint), p.handle, 0), "ipddp_solve")
    return fetch_results!(p)
end

function fetch_results!(p::BatchProblem)
    check(ccall((:ipddp_get_results, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                p.handle, p.status, p.k, p.j, p.l, p.objective, p.primal_inf, p.dual_inf, p.cs_inf, p.μ, p.reg_last,
                p.step_size), "ipddp_get_results")
    return p
end

"solve!(solver) of the reference (src/solve.jl:6-17): warm start from the stored nominal trajectory."
solve!(p::BatchProblem) = (check(ccall((:ipddp_solve, LIB), Cint, (Ptr{Cvoid}, Cint), p.handle, 1), "ipddp_solve");
                           fetch_results!(p))

"""
    solve_many!(problems; total_solves=length(problems)) -> (elapsed_ms, Stats)

Several batches in flight on one GPU (`ipddp_solve_many`; no reference counterpart): the lock-step tail of one batch
overlaps the bulk rounds of the others.  Inputs must have been set with `set_inputs!` / a previous `solve!`.
"""
function solve_many!(ps::Vector{BatchProblem}; total_solves::Int=length(ps), warm_start::Bool=false)
    hs = [p.handle for p in ps]
    ms = Ref{Cdouble}(0.0)
    st = Ref(Stats())
    check(ccall((:ipddp_solve_many, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Ref{Cdouble}, Ref{Stats}),
                hs, length(hs), total_solves, warm_start ? 1 : 0, ms, st), "ipddp_solve_many")
    foreach(fetch_results!, ps)
    return ms[], st[]
end

"Execution tuning that never changes results (`ipddp_set_tuning`): \"fw_spec_max\", \"bw_spec_max\", \"list_sort\", \"bulk_slots\"."
set_tuning!(p::Union{BatchProblem,Nothing}, key::AbstractString, value::Integer) =
    check(ccall((:ipddp_set_tuning, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint),
                p === nothing ? C_NULL : p.handle, key, value), "ipddp_set_tuning")

function get_stats(p::BatchProblem)
    st = Ref(Stats())
    check(ccall((:ipddp_get_stats, LIB), Cint, (Ptr{Cvoid}, Ref{Stats}), p.handle, st), "ipddp_get_stats")
    return st[]
end

"nominal duals (ϕ nc x (N-1) x B, zl and zu nu x (N-1) x B, λ nx x N x B): reference problem.nominal_*_duals"
function get_duals(p::BatchProblem)
    ϕ = zeros(p.nc, p.N - 1, p.B); zl = zeros(p.nu, p.N - 1, p.B); zu = zeros(p.nu, p.N - 1, p.B)
    λ = zeros(p.ns, p.N, p.B)
    check(ccall((:ipddp_get_duals, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                p.handle, ϕ, zl, zu, λ), "ipddp_get_duals")
    return ϕ, zl, zu, λ
end

"""
    get_trace(p, b) -> 12 x nrows matrix

The numeric columns the reference prints per iteration (src/print.jl:13-29) for instance `b` (1-based): k, j, objective,
primal_inf, dual_inf, cs_inf, μ, reg_last, step_size, l, θ, barrier Lagrangian.  Needs `trace_capacity > 0`.
"""
function get_trace(p::BatchProblem, b::Integer)
    n = Ref{Cint}(0)
    # (the argument types of a ccall must be a literal tuple)
    check(ccall((:ipddp_get_trace, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Cint}), p.handle, b - 1, C_NULL, n),
          "ipddp_get_trace")                                                                           # size query
    rows = zeros(12, n[])
    check(ccall((:ipddp_get_trace, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Cint}), p.handle, b - 1, rows, n),
          "ipddp_get_trace")
    return rows
end

"get_trajectory(solver) -> (states ns x N x B, controls nu x (N-1) x B)  (reference src/solver.jl:46-48); zero-padded for stage chains"
function get_trajectory(p::BatchProblem)
    x = zeros(p.ns, p.N, p.B)
    u = zeros(p.nu, p.N - 1, p.B)
    check(ccall((:ipddp_get_trajectory, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), p.handle, x, u),
          "ipddp_get_trajectory")
    return x, u
end


"Launch on the caller's CUDA stream (a `CUstream` / `cudaStream_t` as Ptr{Cvoid}, e.g. from CUDA.jl); C_NULL = own stream."
set_stream!(p::BatchProblem, stream::Ptr{Cvoid}) =
    check(ccall((:ipddp_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), p.handle, stream), "ipddp_set_stream")

# C mirror of `ipddp_queue` (include/ipddp_b200.h): Q queued instances, input and output arrays
mutable struct Queue
    Q::Cint
    x1::Ptr{Cdouble}
    ubar::Ptr{Cdouble}
    params::Ptr{Cdouble}
    lower::Ptr{Cdouble}
    upper::Ptr{Cdouble}
    horizons::Ptr{Cint}
    inputs_on_device::Cint
    status::Ptr{Cint}
    k::Ptr{Cint}
    j::Ptr{Cint}
    l::Ptr{Cint}
    objective::Ptr{Cdouble}
    primal_inf::Ptr{Cdouble}
    dual_inf::Ptr{Cdouble}
    cs_inf::Ptr{Cdouble}
    mu::Ptr{Cdouble}
    reg_last::Ptr{Cdouble}
    step_size::Ptr{Cdouble}
    n_backward::Ptr{Cint}
    n_sweeps::Ptr{Cint}
    n_kkt::Ptr{Cint}
    n_rollouts::Ptr{Cint}
    x::Ptr{Cdouble}
    u::Ptr{Cdouble}
    outputs_on_device::Cint
end

"""
    solve_queue!(prob, x1, controls; params, lower, upper, horizons) -> NamedTuple

Streaming solve (`ipddp_solve_queue`): the Q = size(x1, 2) queued instances flow through the B resident slots of `prob`;
Q may be much larger than B.  Arrays as in `solve!` with Q in place of B.  Returns the SolverData vectors, the work counters
and the trajectories (states nx x N x Q, controls nu x (N-1) x Q).
"""
function solve_queue!(p::BatchProblem, x1::Matrix{Float64}, controls::Array{Float64,3}; params=nothing,
                      lower::Matrix{Float64}, upper::Matrix{Float64}, horizons=nothing)
    Q = size(x1, 2)
    horizons = horizons === nothing ? nothing : convert(Vector{Cint}, horizons)
    ints = [zeros(Cint, Q) for _ in 1:8]
    dbls = [zeros(Cdouble, Q) for _ in 1:7]
    x = zeros(p.ns, p.N, Q); u = zeros(p.nu, p.N - 1, Q)
    GC.@preserve x1 controls params lower upper horizons ints dbls x u begin
        q = Queue(Q, pointer(x1), pointer(controls), params === nothing ? C_NULL : pointer(params), pointer(lower),
                  pointer(upper), horizons === nothing ? C_NULL : pointer(horizons), 0,
                  pointer(ints[1]), pointer(ints[2]), pointer(ints[3]), pointer(ints[4]),
                  pointer(dbls[1]), pointer(dbls[2]), pointer(dbls[3]), pointer(dbls[4]), pointer(dbls[5]), pointer(dbls[6]),
                  pointer(dbls[7]), pointer(ints[5]), pointer(ints[6]), pointer(ints[7]), pointer(ints[8]),
                  pointer(x), pointer(u), 0)
        check(ccall((:ipddp_solve_queue, LIB), Cint, (Ptr{Cvoid}, Ref{Queue}), p.handle, Ref(q)), "ipddp_solve_queue")
    end
    return (status=ints[1], k=ints[2], j=ints[3], l=ints[4], objective=dbls[1], primal_inf=dbls[2], dual_inf=dbls[3],
            cs_inf=dbls[4], μ=dbls[5], reg_last=dbls[6], step_size=dbls[7], n_backward=ints[5], n_sweeps=ints[6],
            n_kkt=ints[7], n_rollouts=ints[8], states=x, controls=u)
end

# =====================================================================================================================
# The reference's exported API (src/InteriorPointDDP.jl:29-45) on top of BatchProblem: same constructors, same call
# sequence (`Solver(T, dynamics, objectives, constraints, bounds; options)`, `solve!(solver, x1, controls)`,
# `get_trajectory(solver)`, results in `solver.data`).  The closures are traced symbolically exactly where the reference
# traces them; instead of `eval`-ing Julia closures the derivatives are emitted through Symbolics' C target (codegen.jl),
# compiled into a model plugin and registered with the library.  Additive extensions for batching: `batch`,
# `num_parameter` (closures may take a trailing parameter vector `p`), per-instance `params` / bounds in `solve!`.
# =====================================================================================================================

"Dynamics(f, num_state, num_control; quasi_newton=false)  (reference src/dynamics.jl:15)"
struct Dynamics
    f::Function
    num_next_state::Int
    num_state::Int
    num_control::Int
    quasi_newton::Bool
    user::Dict{String,Function}      # user-provided derivative closures (src/dynamics.jl:58-61)
    inplace::Bool                    # closures in the reference's in-place form `f!(out, x, u)` (user-derivative form)
end
Dynamics(f::Function, num_state::Int, num_control::Int; quasi_newton::Bool=false) =
    Dynamics(f, num_state, num_state, num_control, quasi_newton, Dict{String,Function}(), false)
"""
    Dynamics(f!, fx!, fu!, num_next_state, num_state, num_control; vfxx=nothing, vfux=nothing, vfuu=nothing, inplace=true)

User-provided dynamics and derivatives (reference src/dynamics.jl:58-61).  As in the reference the closures are in-place:
`f!(y, x, u)`, `fx!(J, x, u)`, `vfxx!(H, x, u, v)` (each optionally with a trailing parameter vector `p`); they are called
once with symbolic arguments and compiled.  `inplace=false` takes closures that return their value instead.
"""
function Dynamics(f::Function, fx::Function, fu::Function, num_next_state::Int, num_state::Int, num_control::Int;
                  vfxx=nothing, vfux=nothing, vfuu=nothing, inplace::Bool=true)
    user = Dict{String,Function}("fx" => fx, "fu" => fu)
    for (k, g) in (("vfxx", vfxx), ("vfux", vfux), ("vfuu", vfuu))
        g === nothing || (user[k] = g)
    end
    return Dynamics(f, num_next_state, num_state, num_control, false, user, inplace)
end

"Objective(f, num_state, num_control)  (reference src/objectives.jl:12)"
struct Objective
    f::Function
    num_state::Int
    num_control::Int
end

"Constraint(c, num_state, num_control; quasi_newton=false, indices_compl=nothing) | Constraint(num_state, num_control)  (reference src/constraints.jl:16,52)"
struct Constraint
    c::Union{Function,Nothing}
    num_constraint::Int              # known up front only in the user-derivative form; otherwise length(c(x, u))
    num_state::Int
    num_control::Int
    quasi_newton::Bool
    indices_compl::Vector{Int}       # 1-based, as in the reference
    user::Dict{String,Function}
    inplace::Bool                    # closures in the reference's in-place form `c!(out, x, u)` (user-derivative form)
end
Constraint(c::Function, num_state::Int, num_control::Int; quasi_newton::Bool=false, indices_compl=nothing) =
    Constraint(c, -1, num_state, num_control, quasi_newton, indices_compl === nothing ? Int[] : collect(Int, indices_compl),
               Dict{String,Function}(), false)
Constraint(num_state::Int, num_control::Int) =
    Constraint(nothing, 0, num_state, num_control, false, Int[], Dict{String,Function}(), false)
"""
    Constraint(c!, cx!, cu!, num_constraint, num_state, num_control; indices_compl, vcxx, vcux, vcuu, inplace=true)

User-provided constraints and derivatives (reference src/constraints.jl:60-64), in-place closures as in the reference
(`c!(out, x, u)`, `cx!(J, x, u)`, `vcxx!(H, x, u, v)`); `inplace=false` takes closures that return their value.
"""
function Constraint(c::Function, cx::Function, cu::Function, num_constraint::Int, num_state::Int, num_control::Int;
                    indices_compl=nothing, vcxx=nothing, vcux=nothing, vcuu=nothing, inplace::Bool=true)
    user = Dict{String,Function}("cx" => cx, "cu" => cu)
    for (k, g) in (("vcxx", vcxx), ("vcux", vcux), ("vcuu", vcuu))
        g === nothing || (user[k] = g)
    end
    return Constraint(c, num_constraint, num_state, num_control, false,
                      indices_compl === nothing ? Int[] : collect(Int, indices_compl), user, inplace)
end

"Bound(lower, upper) | Bound(T, num_control) | Bound(num_control, lower, upper)  (reference src/bounds.jl:12-26)"
struct Bound{T}
    lower::Vector{T}
    upper::Vector{T}
    indices_lower::Vector{Int}
    indices_upper::Vector{Int}
    num_lower::Int
    num_upper::Int
end
function Bound(lower::Vector{T}, upper::Vector{T}) where T
    il = [i for (i, b) in enumerate(lower) if !isinf(b)]
    iu = [i for (i, b) in enumerate(upper) if !isinf(b)]
    return Bound{T}(lower, upper, il, iu, length(il), length(iu))
end
Bound(T, num_control::Int) = Bound(-T(Inf) .* ones(T, num_control), T(Inf) .* ones(T, num_control))
Bound(num_control::Int, lower::T, upper::T) where T = Bound(lower .* ones(T, num_control), upper .* ones(T, num_control))

# What makes two stage objects "the same" when stages are grouped into types: the identity of their closures and their
# sizes / flags -- so `[Dynamics(f, nx, nu) for k = 1:N-1]` (reference experiments/ipddp2/pushing_1_obs.jl:98) is ONE stage
# type exactly like `d = Dynamics(f, nx, nu); [d for k = 1:N-1]` (cartpole_friction.jl:53).
user_ident(user) = Tuple(sort!([(k, objectid(g)) for (k, g) in user]))
ident(d::Dynamics) = (objectid(d.f), d.num_next_state, d.num_state, d.num_control, d.quasi_newton, d.inplace, user_ident(d.user))
ident(o::Objective) = (objectid(o.f), o.num_state, o.num_control)
ident(c::Constraint) = (c.c === nothing ? UInt(0) : objectid(c.c), c.num_constraint, c.num_state, c.num_control,
                        c.quasi_newton, c.inplace, Tuple(c.indices_compl), user_ident(c.user))

"Mirror of the reference's SolverData fields users read (src/data/solver.jl:8-33); vectors of length `batch` (scalars for batch = 1)."
Base.@kwdef mutable struct SolverData
    status = 0
    k = 0
    j = 0
    l = 0
    objective = 0.0
    primal_inf = 0.0
    dual_inf = 0.0
    cs_inf = 0.0
    μ = 0.0
    reg_last = 0.0
    step_size = 0.0
    wall_time::Float64 = 0.0        # seconds: device time of the whole (batched) solve
    solver_time::Float64 = 0.0      # wall_time minus derivative evaluation, as in src/solve.jl:86-87
    fn_eval_time::Float64 = 0.0
end

mutable struct Solver{T}
    problem::BatchProblem
    data::SolverData
    options::Options{T}
    bounds::Vector{Bound{T}}       # one per stage type
    stage_type::Vector{Int}        # 0-based stage type of every running stage
    stage_nx::Vector{Int}          # per-knot sizes (ipddp_stage_layout)
    stage_nu::Vector{Int}
    N::Int
    batch::Int
    num_parameter::Int
end

"""
    Solver(T, dynamics, objectives, constraints, bounds=nothing; options=nothing, batch=1, num_parameter=0, name=nothing)

Reference src/solver.jl:11-26: one Dynamics / Objective / Constraint / Bound per stage.  Stages built from the same objects
form one stage TYPE (one set of compiled device functions); a horizon may use up to 4 types, and the state and control
sizes may change from type to type (reference README.md:18).  The terminal stage has num_control = 0 and no constraints.
"""
function Solver(T, dynamics::Vector{Dynamics}, objectives::Vector{Objective}, constraints::Vector{Constraint},
                bounds=nothing; options=nothing, batch::Int=1, num_parameter::Int=0, name=nothing, device::Int=0,
                trace_capacity::Int=0)
    T === Float64 || error("FP64 only (all reference experiments are Float64)")
    N = length(objectives)
    length(dynamics) + 1 == N == length(constraints) || error("need N-1 dynamics, N objectives, N constraints")
    oN = objectives[end]
    (oN.num_control == 0 && constraints[end].c === nothing) || error("terminal stage: num_control = 0, no constraints")
    bnds = bounds === nothing ? [Bound(T, d.num_control) for d in dynamics] : bounds[1:N-1]
    # ---- stage types: running stages built from the same objects (and equal bounds) share one type
    keys = Any[]; stage_type = Int[]
    for t in 1:N-1
        key = (ident(dynamics[t]), ident(objectives[t]), ident(constraints[t]), bnds[t].lower, bnds[t].upper)
        k = findfirst(==(key), keys)
        k === nothing && (push!(keys, key); k = length(keys))
        push!(stage_type, k)
    end
    length(keys) <= 4 || error("at most 4 distinct stage types per horizon")
    last = stage_type[end]                      # the last running stage's type carries the terminal cost: it goes last
    order = vcat([k for k in 1:length(keys) if k != last], [last])
    rep = [findfirst(==(k), stage_type) for k in order]          # a representative stage of each type
    stage_type = [findfirst(==(k), order) - 1 for k in stage_type]
    opts = options === nothing ? Options{T}() : deepcopy(options)
    lN = (x, p) -> call_with_p(oN.f, (x, T[]), p)       # the terminal objective is called with an empty control vector
    zeroN = (x, p) -> 0 * x[1]
    sms = StageModel[]
    for (k, t) in enumerate(rep)
        d, o, c = dynamics[t], objectives[t], constraints[t]
        islast = k == length(rep)
        push!(sms, trace("pending_s$(k - 1)", d.f, o.f, islast ? lN : zeroN, c.c, d.num_state, d.num_control;
                         num_parameter=num_parameter, qn_dynamics=d.quasi_newton, qn_constraint=c.quasi_newton,
                         indices_compl=c.indices_compl, user=merge(d.user, c.user),
                         nx_term=islast ? oN.num_state : d.num_state, inplace_dynamics=d.inplace,
                         inplace_constraint=c.inplace, num_next_state=d.num_next_state,
                         num_constraint=max(c.num_constraint, 0)))
    end
    for sm in sms      # the kernels' limits, before any compiler is started
        sm.nu >= 1 || error("a running stage needs at least one control")
        sm.nu + sm.nc <= 64 || error("num_control + num_constraint = $(sm.nu + sm.nc) > 64: the warp-level KKT factorisation holds at most 64 rows")
    end
    chain = length(sms) > 1 || sms[1].nxn != sms[1].nx || sms[1].nxt != sms[1].nx
    # the model's identity is its traced source
    tag = name === nothing ? "user_" * bytes2hex(sha256(join(emit_device(sm) for sm in sms)))[1:12] : String(name)
    if chain
        sms = [renamed(sm, "$(tag)_s$(k - 1)") for (k, sm) in enumerate(sms)]
        load_model!(build_plugin(tag, emit_chain(tag, sms)))
    else
        sms = [renamed(sms[1], tag)]
        load_model!(build_plugin(sms[1]))
    end
    prob = BatchProblem(tag, batch, N; options=opts, device=device, indices_compl=Cint.(sms[1].indices_compl .- 1),
                        trace_capacity=trace_capacity)
    check(ccall((:ipddp_set_stage_types, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), prob.handle, Cint.(stage_type)),
          "ipddp_set_stage_types")
    for (k, sm) in enumerate(sms)
        ic = Cint.(sm.indices_compl .- 1)
        check(ccall((:ipddp_set_stage_compl, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Cint), prob.handle, k - 1,
                    isempty(ic) ? C_NULL : ic, length(ic)), "ipddp_set_stage_compl")
    end
    snx = zeros(Cint, N); snu = zeros(Cint, N); snc = zeros(Cint, N)
    check(ccall((:ipddp_stage_layout, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), prob.handle, snx, snu, snc),
          "ipddp_stage_layout")
    return Solver{T}(prob, SolverData(), opts, [bnds[t] for t in rep], stage_type, Int.(snx), Int.(snu), N, batch,
                     num_parameter)
end

unbatch(v, batch) = batch == 1 ? v[1] : v

function fill_data!(s::Solver)
    p = s.problem
    st = get_stats(p)
    s.data = SolverData(status=unbatch(p.status, s.batch), k=unbatch(p.k, s.batch), j=unbatch(p.j, s.batch),
                        l=unbatch(p.l, s.batch), objective=unbatch(p.objective, s.batch),
                        primal_inf=unbatch(p.primal_inf, s.batch), dual_inf=unbatch(p.dual_inf, s.batch),
                        cs_inf=unbatch(p.cs_inf, s.batch), μ=unbatch(p.μ, s.batch), reg_last=unbatch(p.reg_last, s.batch),
                        step_size=unbatch(p.step_size, s.batch), wall_time=st.ms_total * 1e-3,
                        solver_time=(st.ms_total - st.ms_derivs) * 1e-3, fn_eval_time=st.ms_derivs * 1e-3)
    return s.data
end

"""
    solve!(solver, x1, controls; params=nothing, lower=nothing, upper=nothing, horizons=nothing)

Reference src/solve.jl:1-4.  Single instance: `x1::Vector{T}`, `controls::Vector{Vector{T}}` (N entries, the last one
empty), as in the reference.  Batched: `x1` nx x B, `controls` nu x (N-1) x B, `params` np x B, bounds nu x B.
"""
function solve!(s::Solver{T}, x1::Vector{T}, controls::Vector{Vector{T}}; params=nothing, kw...) where T
    p = s.problem
    x1p = zeros(T, p.nx); x1p[1:length(x1)] .= x1              # a chain's first stage may have fewer states than the largest
    u = zeros(T, p.nu, s.N - 1)                                 # per-stage control vectors padded to the largest size
    for t in 1:s.N-1
        u[1:length(controls[t]), t] .= controls[t]
    end
    return solve!(s, repeat(reshape(x1p, :, 1), 1, s.batch), repeat(reshape(u, size(u, 1), size(u, 2), 1), 1, 1, s.batch);
                  params=params === nothing ? nothing : repeat(reshape(params, :, 1), 1, s.batch), kw...)
end
function solve!(s::Solver{T}, x1::Matrix{T}, controls::Array{T,3}; params=nothing, lower=nothing, upper=nothing,
                horizons=nothing) where T
    p = s.problem
    function padded(get, fill)       # one bound vector per stage type, padded to the largest control size: nu*ntypes x B
        m = fill .* ones(T, p.nu, length(s.bounds))
        for (k, b) in enumerate(s.bounds)
            v = get(b)
            m[1:length(v), k] .= v
        end
        return repeat(reshape(m, :, 1), 1, s.batch)
    end
    lo = lower === nothing ? padded(b -> b.lower, -T(Inf)) : lower
    up = upper === nothing ? padded(b -> b.upper, T(Inf)) : upper
    solve!(p, x1, controls; params=params, lower=lo, upper=up, horizons=horizons)
    return fill_data!(s)
end
"solve!(solver): warm start from the stored nominal trajectory (reference src/solve.jl:6-17)"
solve!(s::Solver) = (solve!(s.problem); fill_data!(s))

"get_trajectory(solver) -> (nominal_states, nominal_controls)  (reference src/solver.jl:46-48); batch = 1: vectors of per-stage vectors"
function get_trajectory(s::Solver)
    x, u = get_trajectory(s.problem)
    s.batch == 1 || return x, u
    return [x[1:s.stage_nx[t], t, 1] for t in 1:s.N], vcat([u[1:s.stage_nu[t], t, 1] for t in 1:s.N-1], [Float64[]])
end

end # module
