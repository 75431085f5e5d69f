# codegen.jl -- Symbolics.jl C-target emitter for libipddp_b200.so model plugins.
#
# Retargets the reference's `Symbolics.build_function` call sites (src/dynamics.jl:25-34, src/objectives.jl:23-28,
# src/constraints.jl:25-39) from Julia closures to `target = Symbolics.CTarget()`: the same symbolic Jacobians / Hessians /
# tensor contractions the reference builds are classified entry by entry (structural zero / compile-time constant /
# dynamic), the dynamic entries are emitted as ONE C function body per group, and everything is wrapped into the
# `Model_<name>` struct that the CUDA kernel templates are instantiated with -- the same header layout as
# interiorpointddp.jl_b200/csrc/models_gen/*.cuh, which interiorpointddp.jl_b200/codegen/generate.py writes from SymPy.
# `build_plugin` compiles the header with nvcc into `<name>_<hash>.so` exporting `ipddp_plugin_vtable`; `ipddp_model_load`
# registers it.
#
# NOT EXECUTED in the build image (no Julia, no Symbolics there).  tests/test_abi.py checks what can be checked without
# Julia: every `ccall` of this package against the C prototypes, the struct mirrors, and that the table / constant names
# this emitter writes are exactly the ones the kernel sources read.

using Symbolics
using SHA

# one structurally non-zero entry of a derivative matrix: slot >= 0 indexes the compact tile, slot < 0 the constant table
struct TileEntry
    slot::Int
    i::Int      # 0-based row
    j::Int      # 0-based column
    mat::String
end

# matrices that only feed the upper triangle of the KKT block H (strictly lower entries are dropped)
const UPPER_ONLY = ("luu", "vcuu", "vfuu")

"Symbolic bundle of one stage type: everything the reference's constructors differentiate (18 functions, SURVEY 8 a15)."
struct StageModel
    name::String
    nx::Int
    nu::Int
    nc::Int
    np::Int
    nxn::Int                  # size of the next state f maps to (may differ from nx inside a stage chain)
    nxt::Int                  # size of the state the terminal cost is evaluated on
    f::Vector{Num}            # x+ = f(x, u[, p])
    l::Num                    # stage cost
    lN::Num                   # terminal cost (x[, p])
    c::Vector{Num}            # equality constraints
    mats::Dict{String,Matrix{Num}}     # fx fu lx lu lxx luu lux cx cu vcxx vcux vcuu | vfxx vfux vfuu | Nlx Nlxx
    x::Vector{Num}
    u::Vector{Num}
    v::Vector{Num}            # multipliers of the contractions: phi (nc) for vc*, lambda+ (nx) for vf*
    lam::Vector{Num}
    p::Vector{Num}
    indices_compl::Vector{Int}   # 1-based, as in the reference
    xt::Vector{Num}              # the terminal cost's state symbols
end

"the same stage under another name (the model's name is derived from its emitted source)"
renamed(sm::StageModel, name::String) = StageModel(name, sm.nx, sm.nu, sm.nc, sm.np, sm.nxn, sm.nxt, sm.f, sm.l, sm.lN, sm.c,
                                                   sm.mats, sm.x, sm.u, sm.v, sm.lam, sm.p, sm.indices_compl, sm.xt)

call_with_p(fn, args, p) = applicable(fn, args..., p) ? fn(args..., p) : fn(args...)

"""
Value of a traced closure: `fn(args...[, p])`; for a closure in the reference's in-place form `fn!(out, args...[, p])`
(what the reference's user-derivative constructors take, src/dynamics.jl:49-61, src/constraints.jl:60-64) the symbolic
buffer `out` of the given shape after the call.
"""
function traced(fn, args, p, shape)
    shape === nothing && return call_with_p(fn, args, p)
    out = zeros(Num, shape...)
    call_with_p(fn, (out, args...), p)
    return out
end

"""
    trace(name, f, l, lN, c, nx, nu; num_parameter=0, qn_dynamics=false, qn_constraint=false, indices_compl=Int[],
          user=Dict(), inplace_dynamics=false, inplace_constraint=false, num_next_state=nx, num_constraint=0)

The reference's symbolic differentiation (src/dynamics.jl:15-47, src/objectives.jl:12-33, src/constraints.jl:16-50) for one
stage type.  `user` may hold user-provided derivative closures ("fx", "fu", "vfxx", ... as in src/dynamics.jl:58-61,
src/constraints.jl:60-64): they are traced instead of differentiated and missing contractions stay zero.  With
`inplace_dynamics` / `inplace_constraint` the object's closures (f, fx, ... / c, cx, ...) follow the reference's in-place
convention `f!(out, x, u)` and write into symbolic buffers sized by `num_next_state` / `num_constraint`.
"""
function trace(name, f, l, lN, c, nx::Int, nu::Int; num_parameter::Int=0, qn_dynamics::Bool=false,
               qn_constraint::Bool=false, indices_compl=Int[], user=Dict{String,Function}(), nx_term::Int=nx,
               inplace_dynamics::Bool=false, inplace_constraint::Bool=false, num_next_state::Int=nx,
               num_constraint::Int=0)
    x = Symbolics.variables(:x, 1:nx)
    u = Symbolics.variables(:u, 1:nu)
    p = Symbolics.variables(:p, 1:max(num_parameter, 1))
    y = collect(Num, traced(f, (x, u), p, inplace_dynamics ? (num_next_state,) : nothing))
    nxn = length(y)          # a stage of a chain may map onto a state of another size
    cv = c === nothing ? Num[] : collect(Num, traced(c, (x, u), p, inplace_constraint ? (num_constraint,) : nothing))
    nc = length(cv)
    v = Symbolics.variables(:v, 1:max(nc, 1))
    lam = Symbolics.variables(:lam, 1:nxn)
    lval = call_with_p(l, (x, u), p)
    xt = nx_term == nx ? x : Symbolics.variables(:x, 1:nx_term)     # the terminal cost may see a state of another size
    lNval = call_with_p(lN, (xt,), p)
    m = Dict{String,Matrix{Num}}()
    # a user-provided closure is traced (into a rows x cols buffer if its object is in-place); of a user-derivative object
    # the contractions that were not provided stay zero (the reference leaves those caches untouched); anything else is
    # differentiated as the reference's symbolic constructors do
    inplace_of(key) = (key[1] == 'c' || startswith(key, "vc")) ? inplace_constraint : inplace_dynamics
    pick(key, auto, args...; group_user=false, rows, cols) = begin
        haskey(user, key) && return Matrix{Num}(reshape(collect(Num, traced(user[key], args, p, inplace_of(key) ? (rows, cols) : nothing)), rows, cols))
        (group_user && startswith(key, "v")) ? zeros(Num, rows, cols) : auto()
    end
    udyn = haskey(user, "fx"); ucon = haskey(user, "cx")
    m["fx"] = pick("fx", () -> Symbolics.jacobian(y, x), x, u; rows=nxn, cols=nx)
    m["fu"] = pick("fu", () -> Symbolics.jacobian(y, u), x, u; rows=nxn, cols=nu)
    lx = Symbolics.gradient(lval, x); lu = Symbolics.gradient(lval, u)
    m["lx"] = reshape(lx, :, 1); m["lu"] = reshape(lu, :, 1)
    m["lxx"] = Symbolics.jacobian(lx, x); m["luu"] = Symbolics.jacobian(lu, u); m["lux"] = Symbolics.jacobian(lu, x)
    m["cx"] = nc == 0 ? zeros(Num, 0, nx) : pick("cx", () -> Symbolics.jacobian(cv, x), x, u; rows=nc, cols=nx)
    m["cu"] = nc == 0 ? zeros(Num, 0, nu) : pick("cu", () -> Symbolics.jacobian(cv, u), x, u; rows=nc, cols=nu)
    vv = v[1:nc]
    if qn_constraint || nc == 0      # quasi-Newton constraint object: its contraction caches stay zero
        m["vcxx"] = zeros(Num, nx, nx); m["vcux"] = zeros(Num, nu, nx); m["vcuu"] = zeros(Num, nu, nu)
    else                             # src/constraints.jl:34-36
        m["vcxx"] = pick("vcxx", () -> Symbolics.jacobian(m["cx"]' * vv, x), x, u, vv; group_user=ucon, rows=nx, cols=nx)
        m["vcux"] = pick("vcux", () -> Symbolics.jacobian(m["cu"]' * vv, x), x, u, vv; group_user=ucon, rows=nu, cols=nx)
        m["vcuu"] = pick("vcuu", () -> Symbolics.jacobian(m["cu"]' * vv, u), x, u, vv; group_user=ucon, rows=nu, cols=nu)
    end
    if qn_dynamics                   # src/dynamics.jl:35-38
        m["vfxx"] = zeros(Num, nx, nx); m["vfux"] = zeros(Num, nu, nx); m["vfuu"] = zeros(Num, nu, nu)
    else                             # src/dynamics.jl:28-30
        m["vfxx"] = pick("vfxx", () -> sum(lam[t] .* Symbolics.hessian(y[t], x) for t in 1:nxn), x, u, lam; group_user=udyn, rows=nx, cols=nx)
        m["vfux"] = pick("vfux", () -> sum(lam[t] .* Symbolics.jacobian(Symbolics.gradient(y[t], u), x) for t in 1:nxn), x, u, lam; group_user=udyn, rows=nu, cols=nx)
        m["vfuu"] = pick("vfuu", () -> sum(lam[t] .* Symbolics.hessian(y[t], u) for t in 1:nxn), x, u, lam; group_user=udyn, rows=nu, cols=nu)
    end
    lNx = Symbolics.gradient(lNval, xt)
    m["Nlx"] = reshape(lNx, :, 1); m["Nlxx"] = Symbolics.jacobian(lNx, xt)
    return StageModel(String(name), nx, nu, nc, num_parameter, nxn, nx_term, y, lval, lNval, cv, m, x, u, v, lam, p,
                      collect(Int, indices_compl), xt)
end

isconst_entry(e::Num) = isempty(Symbolics.get_variables(e))
constval(e::Num) = Float64(Symbolics.value(Symbolics.substitute(e, Dict())))

"Classifies the entries of the named matrices (column-major order); returns (entries, constants, dynamic expressions)."
function classify(sm::StageModel, names::Vector{String}, consts::Vector{Float64})
    entries = TileEntry[]
    dyn = Num[]
    for nm in names
        M = sm.mats[nm]
        shown = startswith(nm, "N") ? nm[2:end] : nm
        for j in 1:size(M, 2), i in 1:size(M, 1)
            e = M[i, j]
            iszero(e) && continue                      # structural zero (Base.iszero(::Num))
            (shown in UPPER_ONLY && i > j) && continue
            if isconst_entry(e)
                push!(consts, constval(e))
                push!(entries, TileEntry(-length(consts), i - 1, j - 1, shown))
            else
                push!(dyn, e)
                push!(entries, TileEntry(length(dyn) - 1, i - 1, j - 1, shown))
            end
        end
    end
    return entries, dyn
end

"""
C statements `du[k] = <expr>;` for the expressions `exprs`, emitted by Symbolics' C target
(`build_function(...; target = Symbolics.CTarget())`) with the given argument names, body only.
"""
function c_body(exprs::Vector{Num}, args::Vector, argnames::Vector{Symbol})
    isempty(exprs) && return ""
    src = Symbolics.build_function(exprs, args...; target=Symbolics.CTarget(), fname=:ipddp_body, lhsname=:du,
                                   rhsnames=argnames, expression=Val{true})
    src = String(src)
    body = src[findfirst('{', src)+1:findlast('}', src)-1]
    return expand_literal_pows(body)
end

"""
`pow(<expr>, 2)` -> `((<expr>)*(<expr>))`, `pow(<expr>, 3)` -> `((<expr>)*(<expr>)*(<expr>))` for arbitrary (nested)
base expressions: the reference's generated Julia code evaluates literal powers 2 and 3 by multiplication
(`Base.literal_pow`), which `pow` would not reproduce to the last bit.  Other exponents stay `pow` (the deterministic
`dm::pow`).  The base is a pure expression, so repeating it is safe; nvcc evaluates it once.
"""
function expand_literal_pows(s::AbstractString)
    cs = collect(s)
    n = length(cs)
    out = Char[]
    i = 1
    while i <= n
        ident_before = i > 1 && (isletter(cs[i-1]) || isdigit(cs[i-1]) || cs[i-1] == '_')
        if !ident_before && i + 3 <= n && cs[i] == 'p' && cs[i+1] == 'o' && cs[i+2] == 'w' && cs[i+3] == '('
            depth = 1
            j = i + 4
            comma = 0
            while j <= n
                c = cs[j]
                if c == '('
                    depth += 1
                elseif c == ')'
                    depth -= 1
                    depth == 0 && break
                elseif c == ',' && depth == 1
                    comma = j
                end
                j += 1
            end
            if j <= n && comma > 0                      # cs[j] is the parenthesis that closes this pow(
                base = expand_literal_pows(String(cs[i+4:comma-1]))
                ex = strip(String(cs[comma+1:j-1]))
                if ex == "2"
                    append!(out, collect("(($base)*($base))"))
                elseif ex == "3"
                    append!(out, collect("(($base)*($base)*($base))"))
                else
                    append!(out, collect("pow($base, $ex)"))
                end
                i = j + 1
                continue
            end
        end
        push!(out, cs[i])
        i += 1
    end
    return String(out)
end

store_calls(body::String) = replace(body, r"du\[(\d+)\]\s*=\s*(.*?);" => s"st(\1, \2);")

function void_args(io, names)
    for n in names
        println(io, "    (void)$n;")
    end
end

"""
    emit_device(sm::StageModel) -> String

The generated header: tables `TBL` / `CONSTS` / `FUCOL` / `FUIDX` in namespace `gen_<name>`, `struct Model_<name>` with the
dimension constants, the table offsets `D_<mat>_OFF/_N`, `VF_*`, `DN_*`, and the device functions dyn, cost, costN, con,
derivs, vf, derivsN.  Elementary functions resolve to the deterministic `dm::` implementations (csrc/detmath.cuh).
"""
function emit_device(sm::StageModel)
    n = sm.name
    consts = Float64[]
    d_names = ["fx", "fu", "lx", "lu", "lxx", "luu", "lux", "cx", "cu", "vcxx", "vcux", "vcuu"]
    vf_names = ["vfxx", "vfux", "vfuu"]
    dn_names = ["Nlx", "Nlxx"]
    d_entries, d_dyn = classify(sm, d_names, consts)
    vf_entries, vf_dyn = classify(sm, vf_names, consts)
    dn_entries, dn_dyn = classify(sm, dn_names, consts)
    all_entries = vcat(d_entries, vf_entries, dn_entries)
    fu = sm.mats["fu"]
    fucol = [j - 1 for j in 1:sm.nu if any(!iszero(fu[i, j]) for i in 1:sm.nxn)]
    fuidx = fill(255, sm.nu)
    for (c, j) in enumerate(fucol)
        fuidx[j+1] = c - 1
    end
    io = IOBuffer()
    println(io, "// GENERATED by julia/codegen.jl (Symbolics C target) -- do not edit.")
    println(io, "// Device model functions for '$n'.")
    println(io, "#pragma once\n#include \"model_common.cuh\"")
    println(io, "#define sin(a) dm::sin(a)\n#define cos(a) dm::cos(a)\n#define tan(a) dm::tan(a)\n#define log(a) dm::log(a)")
    println(io, "#define exp(a) dm::exp(a)\n#define pow(a, b) dm::pow((a), (double)(b))\n#define sqrt(a) ::sqrt(a)")
    println(io, "namespace gen_$n {")
    println(io, "IPDDP_TABLE MEntry TBL[$(max(length(all_entries), 1))] = {")
    for e in all_entries
        println(io, "  {$(e.slot), $(e.i), $(e.j)},  // $(e.mat)")
    end
    isempty(all_entries) && println(io, "  {0, 0, 0},")
    println(io, "};")
    println(io, "IPDDP_TABLE double CONSTS[$(max(length(consts), 1))] = {", join((repr(cst) for cst in (isempty(consts) ? [0.0] : consts)), ", "), "};")
    println(io, "IPDDP_TABLE unsigned char FUCOL[$(max(length(fucol), 1))] = {", join(isempty(fucol) ? [0] : fucol, ", "), "};")
    println(io, "IPDDP_TABLE unsigned char FUIDX[$(max(sm.nu, 1))] = {", join(sm.nu == 0 ? [255] : fuidx, ", "), "};")
    println(io, "}")
    println(io, "struct Model_$n {")
    println(io, "  static constexpr const char* NAME = \"$n\";")
    println(io, "  static constexpr int NX = $(sm.nx), NU = $(sm.nu), NC = $(sm.nc), NXN = $(sm.nxn), NP = $(sm.np);")
    println(io, "  static constexpr int NXT = $(sm.nxt);   // state size the terminal cost is evaluated on")
    println(io, "  // a plain model is a chain of one stage type (ipk::for_stage); chains of several types: generate.py")
    println(io, "  static constexpr int NSTAGE = 1;")
    println(io, "  template <int I> using Stage = Model_$n;")
    println(io, "  using Terminal = Model_$n;")
    println(io, "  static constexpr int NTBL = $(length(all_entries)), NCONST = $(length(consts));")
    println(io, "  static constexpr int D_NSLOT = $(length(d_dyn)), VF_NSLOT = $(length(vf_dyn)), DN_NSLOT = $(length(dn_dyn));")
    off = 0
    for (prefix, names, entries) in (("D", d_names, d_entries), ("VF", vf_names, vf_entries), ("DN", dn_names, dn_entries))
        for nm in names
            shown = startswith(nm, "N") ? nm[2:end] : nm
            cnt = count(e -> e.mat == shown, entries)
            println(io, "  static constexpr int $(prefix)_$(shown)_OFF = $off, $(prefix)_$(shown)_N = $cnt;")
            off += cnt
        end
    end
    println(io, "  static constexpr int FU_NC = $(length(fucol));")
    println(io, "  static IPDDP_D int fu_col(int c) { return gen_$n::FUCOL[c]; }")
    println(io, "  static IPDDP_D int fu_idx(int u) { return gen_$n::FUIDX[u]; }")
    println(io, "  static IPDDP_D const MEntry* tbl() { return gen_$n::TBL; }")
    println(io, "  static IPDDP_D const double* consts() { return gen_$n::CONSTS; }")
    xa, ua, pa = sm.x, sm.u, sm.p
    # plain functions: the C target writes du[k] = ...; the wrappers name the output array `du`
    println(io, "  static IPDDP_D void dyn(const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ p, double* __restrict__ du) {")
    void_args(io, ("x", "u", "p")); print(io, c_body(sm.f, [xa, ua, pa], [:x, :u, :p])); println(io, "  }")
    println(io, "  static IPDDP_D void cost(const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ p, double* __restrict__ du) {")
    void_args(io, ("x", "u", "p")); print(io, c_body([sm.l], [xa, ua, pa], [:x, :u, :p])); println(io, "  }")
    println(io, "  static IPDDP_D void costN(const double* __restrict__ x, const double* __restrict__ p, double* __restrict__ du) {")
    void_args(io, ("x", "p")); print(io, c_body([sm.lN], [sm.xt, pa], [:x, :p])); println(io, "  }")
    println(io, "  static IPDDP_D void con(const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ p, double* __restrict__ du) {")
    void_args(io, ("x", "u", "p", "du")); print(io, c_body(sm.c, [xa, ua, pa], [:x, :u, :p])); println(io, "  }")
    # tile functions: du[k] = e  ->  st(k, e)
    va = sm.v
    println(io, "  template <class Store> static IPDDP_D void derivs(const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ p, Store st) {")
    void_args(io, ("x", "u", "v", "p")); print(io, store_calls(c_body(d_dyn, [xa, ua, va, pa], [:x, :u, :v, :p]))); println(io, "  }")
    println(io, "  template <class Store> static IPDDP_D void vf(const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ p, Store st) {")
    void_args(io, ("x", "u", "v", "p")); print(io, store_calls(c_body(vf_dyn, [xa, ua, sm.lam, pa], [:x, :u, :v, :p]))); println(io, "  }")
    println(io, "  template <class Store> static IPDDP_D void derivsN(const double* __restrict__ x, const double* __restrict__ p, Store st) {")
    void_args(io, ("x", "p")); print(io, store_calls(c_body(dn_dyn, [sm.xt, pa], [:x, :p]))); println(io, "  }")
    println(io, "};")
    println(io, "#undef sin\n#undef cos\n#undef tan\n#undef log\n#undef exp\n#undef pow\n#undef sqrt")
    return String(take!(io))
end

"""
    emit_chain(name, stages::Vector{StageModel}) -> String

Header of a stage chain (state / control sizes that change along the horizon; reference README.md:18): the stage structs
plus the composite `Model_<name>` the kernels are instantiated with -- NSTAGE, `Stage<I>`, `Terminal` (the last stage type)
and the maxima over the stage types under the usual names, which size every buffer and stride (the same layout as
interiorpointddp.jl_b200/codegen/generate.py:emit_device_chain).
"""
function emit_chain(name::String, stages::Vector{StageModel})
    io = IOBuffer()
    println(io, "// GENERATED by julia/codegen.jl -- stage chain '$name'.\n#pragma once")
    for sm in stages
        print(io, replace(emit_device(sm), "#pragma once\n" => ""))
    end
    S = ["Model_$(sm.name)" for sm in stages]
    mx(f) = "ipk::cmax(" * join(("$s::$f" for s in S), ", ") * ")"
    println(io, "struct Model_$name {")
    println(io, "  static constexpr const char* NAME = \"$name\";")
    println(io, "  static constexpr int NSTAGE = $(length(S));")
    println(io, "  template <int I> using Stage = typename ipk::TypeAt<I, $(join(S, ", "))>::type;")
    println(io, "  using Terminal = $(S[end]);")
    for f in ("NX", "NU", "NC", "NXN", "NP", "D_NSLOT", "VF_NSLOT", "FU_NC")
        println(io, "  static constexpr int $f = $(mx(f));")
    end
    println(io, "  static constexpr int NXT = Terminal::NXT, DN_NSLOT = Terminal::DN_NSLOT;")
    println(io, "};")
    return String(take!(io))
end

"""
    build_plugin(sm; csrc, outdir, nvcc) -> path of the plugin .so

`csrc` is the directory of the kernel templates (interiorpointddp.jl_b200/csrc).  The plugin's name carries a hash of the
emitted source and of the kernel headers it embeds, so a changed closure or a changed library version never reuses a
stale plugin.  Flags as in interiorpointddp.jl_b200/build.py (sm_100a, no implicit FMA contraction).
"""
build_plugin(sm::StageModel; kw...) = build_plugin(sm.name, emit_device(sm); kw...)
function build_plugin(name::AbstractString, src::AbstractString; csrc::AbstractString=get(ENV, "IPDDP_B200_CSRC", ""),
                      outdir::AbstractString=mktempdir(), nvcc::AbstractString=get(ENV, "NVCC", "nvcc"))
    isdir(csrc) || error("IPDDP_B200_CSRC must point at interiorpointddp.jl_b200/csrc (kernel templates)")
    hdrs = join((read(joinpath(csrc, f), String) for f in sort(readdir(csrc)) if endswith(f, ".cuh") || endswith(f, ".h")))
    tag = bytes2hex(sha256(src * hdrs))[1:12]
    so = joinpath(outdir, "$(name)_$tag.so")
    isfile(so) && return so
    cuh = joinpath(outdir, "$(name)_$tag.cuh")
    cu = joinpath(outdir, "$(name)_$tag.cu")
    write(cuh, src)
    write(cu, "#include \"$cuh\"\n#include \"model_register.cuh\"\nIPDDP_REGISTER_MODEL(Model_$(name), ipddp_plugin_vtable)\n")
    run(`$nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC
         -Xcompiler -fno-gnu-unique -I$csrc -shared $cu -o $so`)
    return so
end
