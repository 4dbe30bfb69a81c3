# model_hist_to_jld.jl -- turn the flat binary written by tonga_b200.checkpoint.export_model_hist into the reference's
# `model.jld` (main_inversion.jl:18: save("model.jld", "model", models), models::Vector{Vector{Model}}), so that the
# reference's plotting scripts (loadnplot.jl:29-32, plot_model_hist MCsub.jl:753) consume GPU results unchanged.
#
#   julia model_hist_to_jld.jl model_hist.bin model.jld      (run inside the reference directory: needs DefStruct.jl)
#
# NOTE: Julia is not installed in the build image, so this script has not been executed there; the Python reader
# `tonga_b200.checkpoint.import_model_hist` parses the same layout and is what the tests drive.
using JLD
include("DefStruct.jl")

function read_model_hist(path::String, tS::Vector{Float64} = Float64[])
    open(path, "r") do io
        String(read(io, 8)) == "TONGAMH1" || error("not a tonga model_hist file")
        n = read(io, Int64); R = read(io, Int64)
        models = Vector{Vector{Model}}()
        for c in 1:n
            nm = read(io, Int64)
            hist = Vector{Model}()
            for j in 1:nm
                K = read(io, Int64); action = read(io, Int64); accept = read(io, Int64)
                phi = read(io, Float64); like = read(io, Float64)
                cells = Array{Float64}(undef, K, 4)            # file order: x[K] y[K] z[K] zeta[K]
                read!(io, cells)
                ptS = Vector{Float64}(undef, R); read!(io, ptS)
                push!(hist, Model(Float64(K), cells[:, 1], cells[:, 2], cells[:, 3], cells[:, 4], phi, ptS,
                                  isempty(tS) ? zeros(R) : tS, like, action, accept, -1.0, -1.0))   # DefStruct.jl:32-48
            end
            push!(models, hist)
        end
        return models
    end
end

if abspath(PROGRAM_FILE) == @__FILE__
    models = read_model_hist(ARGS[1])
    save(length(ARGS) > 1 ? ARGS[2] : "model.jld", "model", models)
end
