# InteriorPointDDPB200.jl -- Julia binding of libipddp_b200.so (C ABI: include/ipddp_b200.h).
#
# NOT EXECUTED in the build image (no Julia there); it documents exactly what a maintainer of
# mingu6/InteriorPointDDP.jl adds to route `solve!` through the B200 library.  It keeps the exported names of
# reference src/InteriorPointDDP.jl:29-45.  Model code generation: the reference's `Symbolics.build_function`
# call sites (src/dynamics.jl:26-34, src/objectives.jl:23-28, src/constraints.jl:25-39) are retargeted to
# `target = Symbolics.CTarget()`; the emitted C bodies are wrapped into the `Model_<name>` struct layout of
# interiorpointddp.jl_b200/csrc/models_gen/*.cuh and compiled with nvcc into a plugin (see INTEGRATION.md).
module InteriorPointDDPB200

export Options, BatchProblem, Stats, solve!, solve_many!, get_trajectory, get_duals, get_trace, get_stats,
       set_tuning!, load_model!

const LIB = get(ENV, "IPDDP_B200_LIB", "libipddp_b200.so")

# mirror of `ipddp_options` == reference Options{T} (src/options.jl:1-38), same field order
Base.@kwdef mutable struct Options
    quasi_newton::Cint = 0
    optimality_tolerance::Cdouble = 1.0e-8
    max_iterations::Cint = 1000
    reset_cache::Cint = 1
    verbose::Cint = 0
    print_frequency::Cint = 10
    μ_init::Cdouble = 1.0
    ineq_dual_init::Cdouble = 1.0
    κ_1::Cdouble = 0.01
    κ_2::Cdouble = 0.01
    reg_1::Cdouble = 1e-4
    reg_min::Cdouble = 1e-20
    reg_max::Cdouble = 1e40
    κ_̄w_p::Cdouble = 100.0
    κ_w_p::Cdouble = 8.0
    κ_w_m::Cdouble = 1.0 / 3.0
    κ_c::Cdouble = 0.25
    δ_c::Cdouble = 1e-8
    κ_ϵ::Cdouble = 10.0
    κ_μ::Cdouble = 0.2
    θ_μ::Cdouble = 1.2
    τ_min::Cdouble = 0.99
    s_max::Cdouble = 100.0
    η_L::Cdouble = 1e-4
    s_L::Cdouble = 2.3
    δ::Cdouble = 1.0
    s_θ::Cdouble = 1.1
    γ_α::Cdouble = 0.05
    γ_θ::Cdouble = 1e-5
    γ_L::Cdouble = 1e-5
    κ_Σ::Cdouble = 1e10
end

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
end

"""
    BatchProblem(model, B, N; options=Options(), device=0, indices_compl=Cint[])

Batched counterpart of `Solver(T, dynamics, objectives, constraints, bounds; options)` (reference src/solver.jl:11-26).
"""
function BatchProblem(model::String, B::Int, N::Int; options::Options=Options(), device::Int=0,
                      indices_compl::Vector{Cint}=Cint[], trace_capacity::Int=0)
    dims = [Ref{Cint}(0) for _ in 1:5]
    check(ccall((:ipddp_model_dims, LIB), Cint, (Cstring, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}),
                model, dims...), "ipddp_model_dims")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ipddp_problem_create, LIB), Cint,
                (Cstring, Cint, Cint, Ptr{Cint}, Cint, Ref{Options}, Cint, Cint, Ref{Ptr{Cvoid}}),
                model, B, N, isempty(indices_compl) ? C_NULL : pointer(indices_compl), length(indices_compl),
                Ref(options), device, trace_capacity, h), "ipddp_problem_create")
    p = BatchProblem(h[], model, B, N, dims[1][], dims[2][], dims[3][], dims[4][], zeros(Cint, B), zeros(Cint, B),
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
    check(ccall((:ipddp_set_inputs, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                p.handle, x1, controls, params === nothing ? C_NULL : params, lower, upper,
                horizons === nothing ? C_NULL : horizons), "ipddp_set_inputs")
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

"Execution tuning that never changes results (`ipddp_set_tuning`): \"fw_spec_max\", \"bw_spec_max\", \"bulk_slots\"."
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
    λ = zeros(p.nx, p.N, p.B)
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
    sig = (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Cint})
    check(ccall((:ipddp_get_trace, LIB), Cint, sig, p.handle, b - 1, C_NULL, n), "ipddp_get_trace")   # size query
    rows = zeros(12, n[])
    check(ccall((:ipddp_get_trace, LIB), Cint, sig, p.handle, b - 1, rows, n), "ipddp_get_trace")
    return rows
end

"get_trajectory(solver) -> (states nx x N x B, controls nu x (N-1) x B)  (reference src/solver.jl:46-48)"
function get_trajectory(p::BatchProblem)
    x = zeros(p.nx, p.N, p.B)
    u = zeros(p.nu, p.N - 1, p.B)
    check(ccall((:ipddp_get_trajectory, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), p.handle, x, u),
          "ipddp_get_trajectory")
    return x, u
end

end # module
