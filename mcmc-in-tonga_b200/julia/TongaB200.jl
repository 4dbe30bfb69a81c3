# TongaB200.jl -- drop-in shim: the reference's hot-path functions re-defined over `ccall` into libtonga_b200.so.
#
# Usage (in main_inversion.jl, AFTER the reference's own includes so that these methods overwrite them):
#
#     @everywhere include("DefStruct.jl"); include("define_TDstructure.jl"); include("MCsub.jl")
#     include("load_data_Tonga.jl"); include("TD_inversion_function.jl")
#     include("/path/to/mcmc-in-tonga_b200/julia/TongaB200.jl")          # <- this file
#     models = TongaB200.run_chains(TD_parameters, dataStruct, 1:TD_parameters.n_chains)   # replaces the pmap line (:15)
#
# Signatures kept exactly (file:line of what each replaces):
#   evaluate(model::Model, dataStruct::DataStruct, TD_parameters::parameters) -> (model, dataStruct, valid)   MCsub.jl:123-185
#   Interpolation(TD_parameters::parameters, model::Model, X, Y, Z) -> Vector{Float64}                        MCsub.jl:306-336
#   v_nearest(x, y, z, mx, my, mz, mv) -> Float64                                                             MCsub.jl:247-263
#   TD_inversion_function(TD_parameters, dataStruct1, chain::Int64) -> model_hist                             TD_inversion_function.jl:7-305
#
# NOTE: Julia is not installed in the build image (SURVEY.md F1), so this file has not been executed there; the Python
# mirror `tonga_b200/api.py` makes the same C-ABI calls with the same argument order and is what the tests drive.
# Host arrays stay owned by Julia's GC: every call is wrapped in GC.@preserve and the library never retains a pointer.

module TongaB200

const LIB = get(ENV, "TONGA_B200_LIB", joinpath(@__DIR__, "..", "lib", "libtonga_b200.so"))

# mirror of `tonga_params` (include/tonga_b200.h)
struct TongaParams
    xmin::Cdouble; xmax::Cdouble; ymin::Cdouble; ymax::Cdouble; zmin::Cdouble; zmax::Cdouble
    sig::Cdouble; zeta_scale::Cdouble; max_sig::Cdouble
    n_iter::Cdouble; burn_in::Cdouble; keep_each::Cdouble
    min_cells::Int32; max_cells::Int32; prior::Int32; debug_prior::Int32; interp_style::Int32; n_actions::Int32
end

last_error() = unsafe_string(ccall((:tonga_last_error, LIB), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error("libtonga_b200 error $rc: $(last_error())")

function make_params(p, ds; n_actions = 4)
    TongaParams(min(ds.xVec...), max(ds.xVec...), min(ds.yVec...), max(ds.yVec...), min(ds.zVec...), max(ds.zVec...),
                p.sig, p.zeta_scale, p.max_sig, p.n_iter, p.burn_in, p.keep_each,
                p.min_cells, p.max_cells, p.prior, p.debug_prior, p.interp_style, n_actions)
end

# one context per dataStruct (the reference passes the same dataStruct to every evaluate call)
const CTX = IdDict{Any,Ptr{Cvoid}}()

function context(ds, p; device = 0)
    get!(CTX, ds) do
        m, R = size(ds.rayX)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        prm = Ref(make_params(p, ds))
        GC.@preserve ds begin
            check(ccall((:tonga_create, LIB), Cint,
                        (Ref{Ptr{Cvoid}}, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                         Ptr{Cdouble}, Ptr{Cdouble}, Ref{TongaParams}, Int32),
                        h, m, R, ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig, prm, device))
        end
        h[]
    end
end

const UTIL = Ref{Ptr{Cvoid}}(C_NULL)
function util_context()
    if UTIL[] == C_NULL
        z = zeros(1, 0)
        prm = Ref(TongaParams(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 1, 4))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:tonga_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                     Ptr{Cdouble}, Ptr{Cdouble}, Ref{TongaParams}, Int32), h, 1, 0, z, z, z, z, z, z, z, prm, 0))
        UTIL[] = h[]
    end
    UTIL[]
end

# ---- evaluate, MCsub.jl:123-185 -------------------------------------------------------------------------------------
function evaluate(model, dataStruct, TD_parameters)
    ctx = context(dataStruct, TD_parameters)
    n = length(dataStruct.tS)
    ptS = Vector{Float64}(undef, n)
    phi = Ref{Cdouble}(0); like = Ref{Cdouble}(0); lg = Ref{Cdouble}(0)
    K = Int32(length(model.xCell))
    GC.@preserve model ptS begin
        check(ccall((:tonga_evaluate, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble},
                     Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
                    ctx, K, model.xCell, model.yCell, model.zCell, model.zeta, 1.0, ptS, phi, like, lg))
    end
    model.phi = phi[]
    model.likelihood = like[]
    if TD_parameters.debug_prior != 1          # MCsub.jl:134-136 returns before ptS / tS are assigned
        model.ptS = ptS
        model.tS = dataStruct.tS               # alias, as MCsub.jl:175
    end
    return model, dataStruct, 1
end

# ---- the misfit alone, MCsub.jl:169-173: phi of given t* vectors (columns of ptS, one per model) against dataStruct.tS / allSig,
# computed by the device kernel every evaluate / proposal ends in.  With the models of a stored reference run (`model.jld`) this
# checks the library against numbers the reference itself produced: misfit(hcat((m.ptS for m in models)...), ds, p) ≈ [m.phi ...].
function misfit(ptS::AbstractMatrix{Float64}, dataStruct, TD_parameters)
    ctx = context(dataStruct, TD_parameters)
    size(ptS, 1) == length(dataStruct.tS) || error("misfit: ptS must have one row per datum")
    pts = Matrix{Float64}(ptS)                 # column-major R x n == the library's [n][R]
    phi = Vector{Float64}(undef, size(pts, 2))
    GC.@preserve pts phi begin
        check(ccall((:tonga_misfit, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx, Int32(size(pts, 2)), pts, C_NULL, phi))
    end
    return phi
end

# ---- Interpolation / v_nearest, MCsub.jl:306-336, 247-263 -----------------------------------------------------------
function interpolate(mx, my, mz, mv, X, Y, Z)
    Xv = collect(Float64, X); Yv = collect(Float64, Y); Zv = collect(Float64, Z)
    out = Vector{Float64}(undef, max(length(Xv), 1))
    np = Ref{Int32}(0)
    GC.@preserve mx my mz mv Xv Yv Zv out begin
        check(ccall((:tonga_interpolate, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ptr{Cdouble}, Int32,
                     Ptr{Cdouble}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ref{Int32}),
                    util_context(), length(mx), mx, my, mz, mv, length(Xv), Xv, length(Yv), Yv, length(Zv), Zv, out, C_NULL, np))
    end
    return out[1:np[]]
end

function Interpolation(TD_parameters, model, X, Y, Z)
    TD_parameters.interp_style == 1 || error("interp_style 2 (IDW) is broken in the reference (MCsub.jl:332)")
    interpolate(model.xCell, model.yCell, model.zCell, model.zeta, X, Y, Z)
end

v_nearest(x, y, z, mx, my, mz, mv) = interpolate(mx, my, mz, mv, [x], [y], [z])[1]

# ---- the chain farm: replaces pmap(x -> TD_inversion_function(TD_parameters, dataStruct, x), 1:n_chains) --------------
# Returns Vector{Vector{Model}} exactly as plot_model_hist consumes it (MCsub.jl:762-767).
# `sampler`: 0 = automatic (shared-memory sampler when the ray set and max_cells fit, else the HBM-streamed one, else the wide one),
# 1 / 2 / 3 = force resident / wide / streamed (TONGA_SAMPLER_* in include/tonga_b200.h).  All give the same chains.
function run_chains(TD_parameters, dataStruct, chains::AbstractUnitRange; seed::UInt64 = UInt64(20260000), device = 0, Model = Main.Model,
                    sampler::Integer = 0)
    ctx = context(dataStruct, TD_parameters; device = device)
    n = length(chains)
    H = Int((TD_parameters.n_iter - TD_parameters.burn_in) / TD_parameters.keep_each) + 1   # TD_inversion_function.jl:25
    ch = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tonga_chains_create_ex, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Int32, Int64, UInt64, Int32, Int32),
                ctx, ch, n, first(chains), seed, H, sampler))
    try
        check(ccall((:tonga_chains_build_starting, LIB), Cint, (Ptr{Cvoid},), ch[]))          # TD_inversion_function.jl:43-45
        check(ccall((:tonga_chains_run, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}, Ptr{Int8}, Ptr{Cdouble}, Ptr{Int32}),
                    ch[], Int64(TD_parameters.n_iter), 0, C_NULL, C_NULL, C_NULL, C_NULL))     # :70-302
        KC = Int(ccall((:tonga_chains_kcap, LIB), Cint, (Ptr{Cvoid},), ch[]))
        R = length(dataStruct.tS)
        nh = Vector{Int32}(undef, n); hK = Matrix{Int32}(undef, H, n)
        cells = Array{Float64,4}(undef, KC, 4, H, n)          # C layout [n][H][4][KC] == Julia (KC, 4, H, n)
        hphi = Matrix{Float64}(undef, H, n); hptS = Array{Float64,3}(undef, R, H, n)
        hiter = Matrix{Int64}(undef, H, n); hact = Matrix{Int32}(undef, H, n); hacc = Matrix{Int32}(undef, H, n); hnext = Matrix{Int32}(undef, H, n)
        GC.@preserve nh hK cells hphi hptS hiter hact hacc hnext begin
            check(ccall((:tonga_chains_get_history, LIB), Cint,
                        (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int32},
                         Ptr{Int32}, Ptr{Int32}), ch[], KC, nh, hK, cells, hphi, hptS, hiter, hact, hacc, hnext))
        end
        like = sum(-log.(dataStruct.allSig * sqrt(2 * pi)) * length(dataStruct.tS))            # MCsub.jl:179 (a constant)
        return [[Model(Float64(hK[j, c]), cells[1:hK[j, c], 1, j, c], cells[1:hK[j, c], 2, j, c], cells[1:hK[j, c], 3, j, c],
                       cells[1:hK[j, c], 4, j, c], hphi[j, c], hptS[:, j, c], dataStruct.tS, like, Int64(hact[j, c]),
                       Int64(hacc[j, c]), -1.0, -1.0) for j in 1:min(nh[c], H)] for c in 1:n]
    finally
        ccall((:tonga_chains_destroy, LIB), Cvoid, (Ptr{Cvoid},), ch[])
    end
end

TD_inversion_function(TD_parameters, dataStruct1, chain::Int64) = run_chains(TD_parameters, dataStruct1, chain:chain)[1]

# ---- one batch over several GPUs (one Julia process per GPU): ray sharding of a STREAMED batch ---------------------------------
# Call on every rank after tonga_chains_create_ex(..., sampler = 3) with the same arguments; `allgather64(bytes)` is the host-side
# exchange of one 64-byte CUDA IPC handle per rank (e.g. MPI.Allgather) returning a Vector of W handles in rank order.
# Returns the mapped peer pointers (close them with shard_disconnect before destroying the batch).
function shard_connect(ch::Ptr{Cvoid}, rank::Integer, world::Integer, device::Integer, allgather64)
    base = Ref{Ptr{Cvoid}}(C_NULL); nbytes = Ref{UInt64}(0)
    check(ccall((:tonga_chains_shard_init, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{Ptr{Cvoid}}, Ref{UInt64}), ch, rank, world, base, nbytes))
    handle = Vector{UInt8}(undef, 64)
    check(ccall((:tonga_ipc_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), base[], handle))
    handles = allgather64(handle)
    peers = Vector{Ptr{Cvoid}}(undef, world)
    for w in 0:world-1
        if w == rank
            peers[w+1] = base[]
        else
            q = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:tonga_ipc_open, LIB), Cint, (Int32, Ptr{UInt8}, Ref{Ptr{Cvoid}}), device, handles[w+1], q))
            peers[w+1] = q[]
        end
    end
    check(ccall((:tonga_chains_shard_connect, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), ch, peers))
    return peers
end
function shard_disconnect(peers::Vector{Ptr{Cvoid}}, rank::Integer, device::Integer)
    for w in 0:length(peers)-1
        w == rank || check(ccall((:tonga_ipc_close, LIB), Cint, (Int32, Ptr{Cvoid}), device, peers[w+1]))
    end
end

end # module

# Method overwrite: make the reference's global names forward to the GPU path.
evaluate(model::Model, dataStruct::DataStruct, TD_parameters::parameters) = TongaB200.evaluate(model, dataStruct, TD_parameters)
Interpolation(TD_parameters::parameters, model::Model, X, Y, Z) = TongaB200.Interpolation(TD_parameters, model, X, Y, Z)
v_nearest(x, y, z, mx, my, mz, mv) = TongaB200.v_nearest(x, y, z, mx, my, mz, mv)
TD_inversion_function(TD_parameters::parameters, dataStruct1::DataStruct, chain::Int64) = TongaB200.TD_inversion_function(TD_parameters, dataStruct1, chain)
