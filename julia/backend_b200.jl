# backend_b200.jl — Julia shim that puts liblwfb90 (B200-native batched integrator, C ABI of include/lwfb90.h) behind
# LWFBrook90.jl's own `simulate!`.
#
# How a maintainer of LWFBrook90.jl uses it:  `include("backend_b200.jl")` at the end of src/LWFBrook90.jl (inside the
# module), set ENV["LWFB_LIB"] to liblwfb90.so, and call
#
#     simulate!(s; backend = :b200)                       # one DiscretizedSPAC
#     simulate!(v; backend = :b200)                       # a Vector{DiscretizedSPAC}: all members in one GPU run
#
# Everything else (loadSPAC, setup, remakeSPAC, get_soil_, get_water_partitioning, plotting) keeps working, because the
# daily records come back in the reference's own state order (src/func_DiffEq_definition_u0.jl:42-73) and are wrapped in
# ComponentVectors with the problem's own axes.
#
# What replaces what:  lwfb_create / lwfb_set_sites / lwfb_set_members / lwfb_run  <->  solve(s.ODEProblem; ...) at
# src/LWFBrook90.jl:643;  B200Solution  <->  the fields of ODESolution that src/func_postprocess.jl reads (:1027,
# :1058-1060, :1119-1120, :1264, :1271, :1280-1287, :1319-1322): .t, .u, .prob, .retcode, .destats and sol(t).
#
# STATUS: written against include/lwfb90.h without a Julia toolchain in the build image - it has not been executed.
# The tested twin of every step below is the Python binding lwfbrook90.jl_b200/{capi,setup,api}.py.
# julia/lwfb90_enums.jl is generated from the C headers (tools/gen_julia_enums.py).

include(joinpath(@__DIR__, "lwfb90_enums.jl"))

using ComponentArrays: ComponentVector, getaxes
using Dates: Day

const liblwfb90 = get(ENV, "LWFB_LIB", "liblwfb90.so")

struct LwfbOpts                      # must match lwfb_opts in include/lwfb90.h
    n_layers::Int32; simulate_isotopes::Int32; t0_day::Int32; n_days::Int32
    reltol::Float64; abstol::Float64; dt_init::Float64; dtmax::Float64
    maxiters::Int64; device::Int32; record_daily::Int32
end

lwfb_error(h) = unsafe_string(ccall((:lwfb_last_error, liblwfb90), Cstring, (Ptr{Cvoid},), h))
check(h, rc) = rc == 0 || error("liblwfb90 error $rc: " * lwfb_error(h))

# ---------------------------------------------------------------------------------------------------------------------
# p (NamedTuple of define_LWFB90_p, src/func_DiffEq_definition_p.jl:421-582) and p.p_soil (KPT_SOILPAR_Mvg1d,
# src/module_KPT.jl:310-330)  ->  scal [LWFB_PS], lay [N, LWFB_PL]  (column-major [N, PL] = the ABI's row-major [PL][N])
# ---------------------------------------------------------------------------------------------------------------------
"params of the parametrized SPAC (param.csv keys); only needed for what `p` does not carry"
_par(s) = s.parametrizedSPAC.pars.params

function flatten_p!(scal::AbstractVector{Float64}, lay::AbstractMatrix{Float64}, s, t0::Real, t1::Real)
    p = s.ODEProblem.p
    soil = p.p_soil
    N = soil.NLAYER
    fill!(scal, 0.0); fill!(lay, 0.0)
    S(k, v) = (scal[k+1] = Float64(v))                 # enums are 0-based
    # solver / isotope parameters
    S(LWFB_S_DPSIMAX, p.p_DPSIMAX); S(LWFB_S_DSWMAX, p.p_DSWMAX); S(LWFB_S_DTIMAX, p.p_DTIMAX)
    S(LWFB_S_VXYLEM, p.p_VXYLEM); S(LWFB_S_DISPERSIVITY, p.p_DISPERSIVITY)
    # location, snow, canopy, interception, plant, soil-surface parameters: same names in `p`
    for (k, v) in ((LWFB_S_LAT, p.p_LAT), (LWFB_S_ESLOPE, p.p_ESLOPE), (LWFB_S_L1, p.p_L1), (LWFB_S_L2, p.p_L2),
                   (LWFB_S_MELFAC, p.p_MELFAC), (LWFB_S_RSTEMP, p.p_RSTEMP), (LWFB_S_FETCH, p.p_FETCH),
                   (LWFB_S_WNDRAT, p.p_WNDRAT), (LWFB_S_Z0G, p.p_Z0G), (LWFB_S_Z0S, p.p_Z0S), (LWFB_S_Z0W, p.p_Z0W),
                   (LWFB_S_ZMINH, p.p_ZMINH), (LWFB_S_ZW, p.p_ZW), (LWFB_S_HS, p.p_HS), (LWFB_S_HR, p.p_HR),
                   (LWFB_S_CZS, p.p_CZS), (LWFB_S_CZR, p.p_CZR), (LWFB_S_C1, p.p_C1), (LWFB_S_C2, p.p_C2),
                   (LWFB_S_C3, p.p_C3), (LWFB_S_SNODEN, p.p_SNODEN), (LWFB_S_CCFAC, p.p_CCFAC),
                   (LWFB_S_GRDMLT, p.p_GRDMLT), (LWFB_S_LAIMLT, p.p_LAIMLT), (LWFB_S_SAIMLT, p.p_SAIMLT),
                   (LWFB_S_MAXLQF, p.p_MAXLQF), (LWFB_S_CR, p.p_CR), (LWFB_S_LWIDTH, p.p_LWIDTH),
                   (LWFB_S_RHOTP, p.p_RHOTP), (LWFB_S_LPC, p.p_LPC), (LWFB_S_KSNVP, p.p_KSNVP),
                   (LWFB_S_ALBSN, p.p_ALBSN), (LWFB_S_ALB, p.p_ALB), (LWFB_S_NN, p.p_NN), (LWFB_S_RM, p.p_RM),
                   (LWFB_S_CINTRL, p.p_CINTRL), (LWFB_S_CINTRS, p.p_CINTRS), (LWFB_S_CINTSL, p.p_CINTSL),
                   (LWFB_S_CINTSS, p.p_CINTSS), (LWFB_S_FRINTL, p.p_FRINTL), (LWFB_S_FRINTS, p.p_FRINTS),
                   (LWFB_S_FSINTL, p.p_FSINTL), (LWFB_S_FSINTS, p.p_FSINTS), (LWFB_S_MXKPL, p.p_MXKPL),
                   (LWFB_S_FXYLEM, p.p_FXYLEM), (LWFB_S_GLMAX, p.p_GLMAX), (LWFB_S_GLMIN, p.p_GLMIN),
                   (LWFB_S_TH, p.p_TH), (LWFB_S_T1, p.p_T1), (LWFB_S_T2, p.p_T2), (LWFB_S_TL, p.p_TL),
                   (LWFB_S_PSICR, p.p_PSICR), (LWFB_S_CVPD, p.p_CVPD), (LWFB_S_R5, p.p_R5),
                   (LWFB_S_RTRAD, p.p_RTRAD), (LWFB_S_MXRTLN, p.p_MXRTLN), (LWFB_S_RSSA, p.p_RSSA),
                   (LWFB_S_RSSB, p.p_RSSB), (LWFB_S_IMPERV, p.p_IMPERV), (LWFB_S_QLAYER, p.p_QLAYER),
                   (LWFB_S_BYPAR, p.p_BYPAR), (LWFB_S_DRAIN, p.p_DRAIN), (LWFB_S_DSLOPE, p.p_DSLOPE),
                   (LWFB_S_GSC, p.p_GSC), (LWFB_S_GSP, p.p_GSP), (LWFB_S_LENGTH_SLOPE, p.p_LENGTH_SLOPE),
                   (LWFB_S_QFFC, p.p_QFFC), (LWFB_S_QFPAR, p.p_QFPAR), (LWFB_S_SWATQX, p.p_SWATQX),
                   (LWFB_S_SWATQF, p.p_SWATQF))
        S(k, v)
    end
    S(LWFB_S_NOOUTF, p.NOOUTF ? 1.0 : 0.0)
    # canopy: the forcing table built by forcing_table() carries the ABSOLUTE values p_DENSEF(t) ... p_SAI(t); the kernel
    # multiplies each by its "baseline" scalar, which is therefore 1
    S(LWFB_S_MAXLAI, 1.0); S(LWFB_S_SAI_BASELINE, 1.0); S(LWFB_S_DENSEF_BASELINE, 1.0); S(LWFB_S_HEIGHT_BASELINE, 1.0)
    # root density: constant over the simulated period (the usual case: stand age beyond the root growth period)
    r0, r1 = p.p_fT_RELDEN(float(t0)), p.p_fT_RELDEN(float(t1))
    r0 == r1 || error("backend_b200: relative root density changes during the simulation (root growth); flatten the " *
                      "growth parameters as lwfbrook90.jl_b200/setup.py:flatten_parameters does (LWFB_S_AGE0, " *
                      "LWFB_S_RGROPER, LWFB_S_RL0, LWFB_L_TINI)")
    S(LWFB_S_RELDEN_CONST, 1.0)
    L(k, v) = (lay[:, k+1] .= v)
    L(LWFB_L_THICK, soil.p_THICK); L(LWFB_L_STONEF, soil.p_STONEF); L(LWFB_L_THSAT, soil.p_THSAT)
    L(LWFB_L_KSAT, soil.p_KSAT); L(LWFB_L_ALPHA, soil.p_MvGα); L(LWFB_L_NPAR, soil.p_MvGn); L(LWFB_L_MPAR, soil.p_MvGm)
    L(LWFB_L_TORT, soil.p_MvGl); L(LWFB_L_THR, soil.p_θr); L(LWFB_L_PSIF, soil.p_PSIF); L(LWFB_L_THETAF, soil.p_THETAF)
    L(LWFB_L_PSIG, soil.p_PSIG); L(LWFB_L_SWATMAX, soil.p_SWATMAX); L(LWFB_L_WETF, soil.p_WETF)
    L(LWFB_L_INFRAC, p.p_INFRAC); L(LWFB_L_RELDEN, r0); L(LWFB_L_TINI, 0.0)
    lay[1:N-1, LWFB_L_DZ+1] .= soil.dz_m[1:N-1]
    return nothing
end

"forcing table [LWFB_F, n_rows]: what the closures p_GLOBRAD(t) ... p_δ2H_PREC(t) return at integer t (LWFB_F_* order)"
function forcing_table(p, n_rows::Integer)
    F = zeros(LWFB_F, n_rows)
    iso = p.simulate_isotopes
    for d in 0:n_rows-1
        t = float(d)
        c = d + 1
        F[LWFB_F_GLOBRAD+1, c] = p.p_GLOBRAD(t); F[LWFB_F_TMAX+1, c] = p.p_TMAX(t); F[LWFB_F_TMIN+1, c] = p.p_TMIN(t)
        F[LWFB_F_VAPPRES+1, c] = p.p_VAPPRES(t); F[LWFB_F_WIND+1, c] = p.p_WIND(t); F[LWFB_F_PREC+1, c] = p.p_PREC(t)
        F[LWFB_F_DENSEF_REL+1, c] = p.p_DENSEF(t); F[LWFB_F_HEIGHT_REL+1, c] = p.p_HEIGHT(t)
        F[LWFB_F_LAI_REL+1, c] = p.p_LAI(t); F[LWFB_F_SAI_REL+1, c] = p.p_SAI(t)
        F[LWFB_F_DOY+1, c] = p.p_DOY(t)
        F[LWFB_F_IHD+1, c] = floor((p.p_DURATN[p.p_MONTHN(t)] + 0.1) / 2)          # IHD of INTER24, src/module_EVP.jl:645
        if iso
            # the value of row d holds for t in (d - 0.5, d + 0.5] (Constant{Next} on sample dates + 0.5 d)
            F[LWFB_F_D18O+1, c] = p.p_δ18O_PREC(t); F[LWFB_F_D2H+1, c] = p.p_δ2H_PREC(t)
        end
    end
    return F
end

# ---------------------------------------------------------------------------------------------------------------------
# the solution object
# ---------------------------------------------------------------------------------------------------------------------
struct B200Solution{U,P}
    t::Vector{Float64}
    u::Vector{U}                 # ComponentVectors with the axes of prob.u0
    prob::P
    retcode::Symbol
    destats::NamedTuple{(:naccept, :nreject, :nf),NTuple{3,Int64}}
end
Base.getproperty(sol::B200Solution, f::Symbol) = f === :stats ? getfield(sol, :destats) : getfield(sol, f)
"sol(t): the saved vector at a saved t, linear interpolation between daily records otherwise (saveat is set in the
reference too, so its solution is not dense either)"
function (sol::B200Solution)(t::Real)
    tt = getfield(sol, :t); u = getfield(sol, :u)
    (t < tt[1] || t > tt[end]) && error("t = $t outside the solution time span")
    j = searchsortedfirst(tt, t)
    tt[j] == t && return u[j]
    w = (t - tt[j-1]) / (tt[j] - tt[j-1])
    return u[j-1] .+ w .* (u[j] .- u[j-1])
end
(sol::B200Solution)(ts::AbstractVector) = [sol(t) for t in ts]
const RETCODES = (:Success, :Unstable, :DtLessThanMin, :MaxIters, :RSSError, :InvalidInput)
successful(sol::B200Solution) = getfield(sol, :retcode) === :Success

# ---------------------------------------------------------------------------------------------------------------------
# simulate!
# ---------------------------------------------------------------------------------------------------------------------
"""
    simulate!(v::Vector{DiscretizedSPAC}; backend = :b200, device = 0, reltol = 1e-5, abstol = 1e-6, dt = 1e-3,
              dtmax = 60/60/24, maxiters = 1_000_000, assert_retcode = true)

Integrates all members in one run on one B200.  The members share grid (N), isotope switch and tspan; they may differ
in every parameter that `remakeSPAC` can change.  Members with the same `parametrizedSPAC.forcing` object share a site.
"""
function simulate!(v::AbstractVector{<:DiscretizedSPAC}; backend = :b200, device = 0, reltol = 1e-5, abstol = 1e-6,
                   dt = 1e-3, dtmax = 60 / 60 / 24, maxiters = 1_000_000, assert_retcode = true)
    backend === :b200 || return foreach(s -> simulate!(s; assert_retcode), v)
    s1 = first(v); p1 = s1.ODEProblem.p
    N = p1.p_soil.NLAYER; iso = p1.simulate_isotopes
    t0, t1 = Int.(s1.ODEProblem.tspan); nd = t1 - t0
    L = length(s1.ODEProblem.u0); M = length(v)
    L == LWFB_STATE_LEN(N, iso) || error("unexpected state length")
    o = Ref(LwfbOpts(N, iso, t0, nd, reltol, abstol, dt, dtmax, maxiters, device, 1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:lwfb_create, liblwfb90), Cint, (Ref{Ptr{Cvoid}}, Ref{LwfbOpts}), h, o)
    rc == 0 || error("lwfb_create: $rc (is a CUDA device available?)")
    try
        # sites: one forcing table per distinct forcing object
        forcings = unique(objectid(s.parametrizedSPAC.forcing) for s in v)
        site = Int32[findfirst(==(objectid(s.parametrizedSPAC.forcing)), forcings) - 1 for s in v]
        n_rows = t1 + 1
        F = Array{Float64}(undef, LWFB_F, n_rows, length(forcings))
        for (k, id) in enumerate(forcings)
            F[:, :, k] .= forcing_table(v[findfirst(s -> objectid(s.parametrizedSPAC.forcing) == id, v)].ODEProblem.p, n_rows)
        end
        check(h[], ccall((:lwfb_set_sites, liblwfb90), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), h[], length(forcings), n_rows, F))
        scal = Matrix{Float64}(undef, LWFB_PS, M); lay = Array{Float64}(undef, N, LWFB_PL, M); u0 = Matrix{Float64}(undef, L, M)
        for (m, s) in enumerate(v)
            flatten_p!(view(scal, :, m), view(lay, :, :, m), s, t0, t1)
            u0[:, m] .= s.ODEProblem.u0                       # ComponentVector memory order = the ABI's order
        end
        check(h[], ccall((:lwfb_set_members, liblwfb90), Cint,
                         (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h[], M, site, scal, lay, u0))
        check(h[], ccall((:lwfb_run, liblwfb90), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h[], C_NULL))
        status = zeros(Int32, M); counters = zeros(Int64, LWFB_NCOUNTERS, M)
        check(h[], ccall((:lwfb_get_status, liblwfb90), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}), h[], status, counters))
        rec = Array{Float64}(undef, L, nd + 1, M)             # column-major [L, day, member] = the ABI's [member][day][L]
        check(h[], ccall((:lwfb_get_daily, liblwfb90), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), h[], 0, M, rec))
        tt = collect(float(t0):float(t1))
        for (m, s) in enumerate(v)
            ax = getaxes(s.ODEProblem.u0)
            s.ODESolution = B200Solution(tt, [ComponentVector(rec[:, k, m], ax) for k in 1:nd+1], s.ODEProblem,
                                         RETCODES[status[m]+1],
                                         (naccept = counters[LWFB_C_NACCEPT+1, m], nreject = counters[LWFB_C_NREJECT+1, m],
                                          nf = counters[LWFB_C_NF+1, m]))
            s.ODESolution_datetime = s.parametrizedSPAC.reference_date .+ Day.(t0:t1)
            assert_retcode && @assert successful(s.ODESolution) "Problem with simulation: Return code of simulation was '$(s.ODESolution.retcode)'"
        end
    finally
        ccall((:lwfb_destroy, liblwfb90), Cvoid, (Ptr{Cvoid},), h[])
    end
    return nothing
end

"`simulate!(s; backend = :b200)`: one member on the GPU; without the keyword the reference's CPU solve runs"
simulate!(s::DiscretizedSPAC, ::Val{:b200}; kwargs...) = simulate!([s]; backend = :b200, kwargs...)

# ---------------------------------------------------------------------------------------------------------------------
# parameter estimation over many members without building DiscretizedSPACs one by one: the batched setup on the GPU
# ---------------------------------------------------------------------------------------------------------------------
"""
    run_knob_ensemble(base::DiscretizedSPAC, knobs::Matrix{Float64}, over_idx::Vector{Int32}; obs..., devices = [0])

`knobs` is [LWFB_K + length(over_idx), M] (one column per member, NaN = keep the base value; see include/lwfb90_setup.h):
the member blocks are built on the device(s) by lwfb_multi_setup_members, integrated, scored against the observations
during the run, and the scores of the whole ensemble come back ([LWFB_NSCORES, M]).  With several devices the members are
sharded and the scores all-gathered peer to peer inside the library (no NCCL.jl needed).
"""
function run_knob_ensemble(base::DiscretizedSPAC, knobs::Matrix{Float64}, over_idx::Vector{Int32};
                           obs_day::Vector{Int32}, obs_layer::Vector{Int32}, obs_theta::Union{Nothing,Matrix{Float64}} = nothing,
                           obs_d18o::Union{Nothing,Matrix{Float64}} = nothing, devices::Vector{Int32} = Int32[0],
                           profiles::Matrix{Float64} = zeros(0, 0), psim_base::Vector{Float64}, ilayer::Integer,
                           cost::Union{Nothing,Vector{Float64}} = nothing)
    p = base.ODEProblem.p
    N = p.p_soil.NLAYER; iso = p.simulate_isotopes
    t0, t1 = Int.(base.ODEProblem.tspan); nd = t1 - t0
    M = size(knobs, 2)
    scal = Vector{Float64}(undef, LWFB_PS); lay = Matrix{Float64}(undef, N, LWFB_PL)
    flatten_p!(scal, lay, base, t0, t1)
    u0 = collect(Float64, base.ODEProblem.u0)
    o = Ref(LwfbOpts(N, iso, t0, nd, 1e-5, 1e-6, 1e-3, 60 / 60 / 24, 1_000_000, 0, 0))     # record_daily = 0
    mh = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:lwfb_multi_create, liblwfb90), Cint, (Ref{Ptr{Cvoid}}, Ref{LwfbOpts}, Int32, Ptr{Int32}), mh, o, length(devices), devices)
    rc == 0 || error("lwfb_multi_create: $rc")
    mcheck(rc) = rc == 0 || error("liblwfb90 error $rc: " * unsafe_string(ccall((:lwfb_multi_last_error, liblwfb90), Cstring, (Ptr{Cvoid},), mh[])))
    scores = Matrix{Float64}(undef, LWFB_NSCORES, M)
    try
        F = forcing_table(p, t1 + 1)
        mcheck(ccall((:lwfb_multi_set_sites, liblwfb90), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), mh[], 1, t1 + 1, F))
        mcheck(ccall((:lwfb_multi_setup_members, liblwfb90), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Int32}, Int32, Ptr{Float64},
                      Ptr{Float64}, Int32, Ptr{Float64}),
                     mh[], M, scal, lay, u0, knobs, length(over_idx), over_idx, size(profiles, 2), profiles, psim_base, ilayer,
                     cost === nothing ? C_NULL : cost))
        mcheck(ccall((:lwfb_multi_set_observations, liblwfb90), Cint,
                     (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                     mh[], length(obs_day), obs_day, length(obs_layer), obs_layer,
                     obs_theta === nothing ? C_NULL : permutedims(obs_theta), obs_d18o === nothing ? C_NULL : permutedims(obs_d18o)))
        mcheck(ccall((:lwfb_multi_run, liblwfb90), Cint, (Ptr{Cvoid},), mh[]))
        mcheck(ccall((:lwfb_multi_allgather_scores, liblwfb90), Cint,
                     (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ptr{Float64}),
                     mh[], 0, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, scores))
    finally
        ccall((:lwfb_multi_destroy, liblwfb90), Cvoid, (Ptr{Cvoid},), mh[])
    end
    return scores
end
