# VoronoiRTB200.jl — the Julia-side binding a VoronoiRT maintainer adds to route the irregular-grid hot path
# through libvrt.so.  Same function names and signatures as the reference (src/voronoi_utils.jl:36,
# src/irregular_ray_tracing.jl:15,96, src/lambda_iteration.jl:60,207, src/rates.jl:154, src/populations.jl:191,
# src/characteristics.jl:19,110, src/lambda_continuum.jl:1,58), so the
# entry scripts (compare_searchlight.jl, compare_continuum.jl, compare_line.jl) run unchanged after
#     include("VoronoiRTB200.jl"); using .VoronoiRTB200
# NOTE: no Julia toolchain exists in the build image, so this file has not been executed here; it is the
# documented ccall stub for include/vrt.h (see INTEGRATION.md).  The Python mirror voronoirt_b200/api.py makes exactly
# the same calls and is what the parity tests drive.
module VoronoiRTB200

using Unitful

const libvrt = get(ENV, "LIBVRT", joinpath(@__DIR__, "..", "voronoirt_b200", "libvrt.so"))

struct VRTError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint) = rc == 0 ? nothing : throw(VRTError(rc, unsafe_string(ccall((:vrt_last_error, libvrt), Cstring, ()))))

# ---------------------------------------------------------------- structs of include/vrt.h
struct vrt_line
    nlam::Int64
    lidx::NTuple{4,Int64}
    lambda0::Float64
    Aji::Float64; Bji::Float64; Bij::Float64
    chi_i::Float64; chi_j::Float64; chi_inf::Float64
    gi::Int64; gj::Int64; Z::Int64
    atom_weight::Float64
    c_unsold::Float64; gamma_natural::Float64; c_linear_stark::Float64; c_quadratic_stark::Float64
end
struct vrt_site_data
    temperature::Ptr{Float64}; electron_density::Ptr{Float64}; hydrogen_density::Ptr{Float64}
    velocity_z::Ptr{Float64}; velocity_x::Ptr{Float64}; velocity_y::Ptr{Float64}; doppler_width::Ptr{Float64}
    alpha_cont::Ptr{Float64}; destruction::Ptr{Float64}; C::Ptr{Float64}; lte_pops::Ptr{Float64}
end
struct vrt_quadrature
    n_dirs::Int64
    weights::Ptr{Float64}; theta::Ptr{Float64}; phi::Ptr{Float64}
end
struct vrt_config
    n_sweeps::Int32; dir_begin::Int32
    p::Float64
    lam_begin::Int64; lam_end::Int64; lam_chunk::Int64
    prune::Int32; dir_end::Int32
    cell_shard_rank::Int32; cell_shard_count::Int32
end
struct vrt_result
    iterations::Int32; converged::Int32
    diff::Float64; seconds::Float64
end

# ---------------------------------------------------------------- read_cell (src/voronoi_utils.jl:36-85)
mutable struct Grid
    h::Ptr{Cvoid}
    n::Int
end
const GRIDS = IdDict{Any,Grid}()      # NeighbourMatrix -> device grid built from it

function read_cell(fname::String, n_sites::Int, positions::Matrix{<:Unitful.Length},
                   x_min, x_max, y_min, y_max)
    ld = Ref{Int64}(0)
    check(ccall((:vrt_read_neighbours, libvrt), Cint, (Cstring, Int64, Ptr{Int64}, Int64, Ref{Int64}), fname, n_sites, C_NULL, 0, ld))
    NeighbourMatrix = zeros(Int64, n_sites, ld[])
    check(ccall((:vrt_read_neighbours, libvrt), Cint, (Cstring, Int64, Ptr{Int64}, Int64, Ref{Int64}), fname, n_sites, NeighbourMatrix, ld[], ld))
    pos = ustrip.(u"m", positions)
    bounds = Float64[0, 0, ustrip(u"m", x_min), ustrip(u"m", x_max), ustrip(u"m", y_min), ustrip(u"m", y_max)]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vrt_grid_create, libvrt), Cint, (Int64, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                n_sites, pos, NeighbourMatrix, ld[], bounds, h))
    g = Grid(h[], n_sites)
    finalizer(x -> ccall((:vrt_grid_destroy, libvrt), Cvoid, (Ptr{Cvoid},), x.h), g)
    GRIDS[NeighbourMatrix] = g
    layers(down) = begin
        L = Ref{Int64}(0)
        check(ccall((:vrt_grid_num_layers, libvrt), Cint, (Ptr{Cvoid}, Int32, Ref{Int64}), g.h, down, L))
        perm = Vector{Int64}(undef, n_sites); off = Vector{Int64}(undef, L[] + 1)
        check(ccall((:vrt_grid_get_layers, libvrt), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}), g.h, down, perm, off))
        perm, off
    end
    perm_up, layers_up = layers(0)
    perm_down, layers_down = layers(1)
    Delaunay_lines = Array{Float64,3}(undef, 3, ld[] - 1, n_sites)
    check(ccall((:vrt_grid_get_delaunay_lines, libvrt), Cint, (Ptr{Cvoid}, Ptr{Float64}), g.h, Delaunay_lines))
    return positions, NeighbourMatrix, Delaunay_lines, layers_up, layers_down, perm_up, perm_down
end

grid_of(sites) = GRIDS[sites.neighbours]

# ---------------------------------------------------------------- Delaunay_upII / Delaunay_downII (src/irregular_ray_tracing.jl)
function _formal(k, S, I_0, α, sites, n_sweeps, p, down)
    Sv = ustrip.(u"kW*m^-2*nm^-1", S); αv = ustrip.(u"m^-1", α); I0v = ustrip.(u"kW*m^-2*nm^-1", I_0)
    nlam = ndims(Sv) == 1 ? 1 : size(Sv, 1)
    I = similar(Sv)
    check(ccall((:vrt_formal_solve, libvrt), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                grid_of(sites).h, Float64.(k), down, p, n_sweeps, nlam, Sv, αv, I0v, I))
    return I * u"kW*m^-2*nm^-1"
end
Delaunay_upII(k, S, I_0, α, sites, n_sweeps::Int, p::Float64=7.0) = _formal(k, S, I_0, α, sites, n_sweeps, p, 0)
Delaunay_downII(k, S, I_0, α, sites, n_sweeps::Int, p::Float64=7.0) = _formal(k, S, I_0, α, sites, n_sweeps, p, 1)

# ---------------------------------------------------------------- Λ_voronoi (src/lambda_iteration.jl:207-297)
# The host computes the Transparency.jl quantities exactly as the reference does before its loop (:216-247) and hands
# them over; the loop itself (J_λ_voronoi, S update, calculate_R, get_revised_populations, criterion) runs on the GPU.
function Λ_voronoi(ϵ::AbstractFloat, maxiter::Integer, sites, line, quadrature::String, DATA::String;
                   LTE_pops, α_cont, ελ, C, line_struct::vrt_line, on_iteration=nothing)
    tab = readdlm_quadrature(quadrature)                       # weights θ ϕ (functions.jl:33-63; the table, not the path, crosses the ABI)
    w, th, ph = tab[:, 1], tab[:, 2], tab[:, 3]
    q = Ref(vrt_quadrature(length(w), pointer(w), pointer(th), pointer(ph)))
    cfg = Ref(vrt_config(3, 0, 7.0, 0, 0, 0, 1, 0, 0, 0))
    vec(x, u) = Float64.(ustrip.(u, x))
    T = vec(sites.temperature, u"K"); ne = vec(sites.electron_density, u"m^-3"); NH = vec(sites.hydrogen_populations, u"m^-3")
    vz = vec(sites.velocity_z, u"m/s"); vx = vec(sites.velocity_x, u"m/s"); vy = vec(sites.velocity_y, u"m/s")
    dD = vec(line.ΔD, u"nm"); ac = vec(α_cont, u"m^-1"); el = Float64.(ελ); Cv = vec(C, u"s^-1"); lte = vec(LTE_pops, u"m^-3")
    lam = vec(line.λ, u"nm")
    sd = Ref(vrt_site_data(pointer(T), pointer(ne), pointer(NH), pointer(vz), pointer(vx), pointer(vy), pointer(dD),
                           pointer(ac), pointer(el), pointer(Cv), pointer(lte)))
    s = Ref{Ptr{Cvoid}}(C_NULL)
    res = Ref(vrt_result(0, 0, 0.0, 0.0))
    n, nλ = sites.n, length(lam)
    S = Matrix{Float64}(undef, nλ, n); J = similar(S); pops = Matrix{Float64}(undef, n, 3)
    GC.@preserve w th ph T ne NH vz vx vy dD ac el Cv lte lam begin
        check(ccall((:vrt_solver_create_line, libvrt), Cint,
                    (Ptr{Cvoid}, Ref{vrt_line}, Ptr{Float64}, Ref{vrt_site_data}, Ref{vrt_quadrature}, Ref{vrt_config}, Ref{Ptr{Cvoid}}),
                    grid_of(sites).h, Ref(line_struct), lam, sd, q, cfg, s))
        # per-iteration hook: the reference writes populations and S to its HDF5 file every iteration (:280-281)
        cb = on_iteration === nothing ? C_NULL : @cfunction($on_iteration, Cint, (Ptr{Cvoid}, Ptr{Cvoid}))
        check(ccall((:vrt_lambda_iterate, libvrt), Cint, (Ptr{Cvoid}, Float64, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ref{vrt_result}),
                    s[], ϵ, maxiter, cb, C_NULL, res))
        check(ccall((:vrt_get_state, libvrt), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s[], S, J, pops))
        ccall((:vrt_solver_destroy, libvrt), Cvoid, (Ptr{Cvoid},), s[])
    end
    return J * u"kW*m^-2*nm^-1", S * u"kW*m^-2*nm^-1", α_cont, pops * u"m^-3"
end

# get_revised_populations (src/populations.jl:191-221)
function get_revised_populations(R::Array{<:Unitful.Frequency,3}, C::Array{<:Unitful.Frequency,3}, atom_density::Vector)
    n = length(atom_density)
    pops = Matrix{Float64}(undef, n, 3)
    check(ccall((:vrt_get_revised_populations, libvrt), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                n, ustrip.(u"s^-1", R), ustrip.(u"s^-1", C), ustrip.(u"m^-3", atom_density), pops))
    return pops * u"m^-3"
end

# ---------------------------------------------------------------- regular grid (src/characteristics.jl:19-95, :110-180)
# atmos is the reference's Atmosphere (src/atmosphere.jl:22-31): z, x, y with the periodic ghost columns in x and y.
function _short_characteristics(k::Vector, S_0::Array{<:Any,3}, I_0::Matrix, α::Array{<:Any,3}, atmos, n_sweeps::Int, down::Int)
    z = Float64.(ustrip.(u"m", atmos.z)); x = Float64.(ustrip.(u"m", atmos.x)); y = Float64.(ustrip.(u"m", atmos.y))
    S = Float64.(ustrip.(u"kW*m^-2*nm^-1", S_0)); I0 = Float64.(ustrip.(u"kW*m^-2*nm^-1", I_0)); a = Float64.(ustrip.(u"m^-1", α))
    I = similar(S)
    kk = Float64.(k)
    check(ccall((:vrt_regular_formal_solve, libvrt), Cint,
                (Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int64,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                length(z), length(x), length(y), z, x, y, kk, down, n_sweeps, 1, S, a, I0, I, C_NULL))
    return I * u"kW*m^-2*nm^-1"
end
short_characteristics_up(k, S_0, I_0, α, atmos; n_sweeps=3) = _short_characteristics(k, S_0, I_0, α, atmos, n_sweeps, 0)
short_characteristics_down(k, S_0, I_0, α, atmos; n_sweeps=3) = _short_characteristics(k, S_0, I_0, α, atmos, n_sweeps, 1)

# J_λ_regular, continuum form (src/lambda_continuum.jl:1-24); I_0 = blackbody_λ.(500u"nm", atmos.temperature[1,:,:]) (:16)
function J_λ_regular(S_λ::AbstractArray, α_cont::AbstractArray, atmos, quadrature::String; I_0)
    tab = readdlm_quadrature(quadrature)
    w, th, ph = tab[:, 1], tab[:, 2], tab[:, 3]
    q = Ref(vrt_quadrature(length(w), pointer(w), pointer(th), pointer(ph)))
    z = Float64.(ustrip.(u"m", atmos.z)); x = Float64.(ustrip.(u"m", atmos.x)); y = Float64.(ustrip.(u"m", atmos.y))
    S = Float64.(ustrip.(u"kW*m^-2*nm^-1", S_λ)); a = Float64.(ustrip.(u"m^-1", α_cont)); I0 = Float64.(ustrip.(u"kW*m^-2*nm^-1", I_0))
    J = similar(S)
    GC.@preserve w th ph check(ccall((:vrt_regular_mean_intensity, libvrt), Cint,
                (Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{vrt_quadrature}, Int32, Int64,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                length(z), length(x), length(y), z, x, y, q, 3, 1, S, a, I0, C_NULL, J))
    return J * u"kW*m^-2*nm^-1"
end

# Λ_regular, continuum (src/lambda_continuum.jl:58-107): α_cont, ε_λ, B_0 as computed at :66-85
function Λ_regular(ϵ::AbstractFloat, maxiter::Integer, atmos, quadrature::String; α_cont, ε_λ, B_0)
    tab = readdlm_quadrature(quadrature)
    w, th, ph = tab[:, 1], tab[:, 2], tab[:, 3]
    q = Ref(vrt_quadrature(length(w), pointer(w), pointer(th), pointer(ph)))
    z = Float64.(ustrip.(u"m", atmos.z)); x = Float64.(ustrip.(u"m", atmos.x)); y = Float64.(ustrip.(u"m", atmos.y))
    a = Float64.(ustrip.(u"m^-1", α_cont)); e = Float64.(ε_λ); B = Float64.(ustrip.(u"kW*m^-2*nm^-1", B_0))
    S = similar(B); J = similar(B)
    res = Ref(vrt_result(0, 0, 0.0, 0.0))
    GC.@preserve w th ph check(ccall((:vrt_regular_lambda_iterate, libvrt), Cint,
                (Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{vrt_quadrature}, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Float64, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{vrt_result}),
                length(z), length(x), length(y), z, x, y, q, 3, a, e, B, ϵ, maxiter, C_NULL, C_NULL, S, J, res))
    res[].converged == 1 ? println("Converged in $(res[].iterations) iterations") : println("Did not converge inside scope")
    return J * u"kW*m^-2*nm^-1", S * u"kW*m^-2*nm^-1", α_cont
end

# Λ_regular, NLTE line (src/lambda_iteration.jl:116-205): a regular-grid handle feeds the same engine as Λ_voronoi.  Every
# per-site array is the (nz, nx, ny) Julia array as it lies in memory; S, J are (nλ, nz, nx, ny), populations (nz, nx, ny, 3).
function regular_grid(atmos)
    z = Float64.(ustrip.(u"m", atmos.z)); x = Float64.(ustrip.(u"m", atmos.x)); y = Float64.(ustrip.(u"m", atmos.y))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vrt_regular_grid_create, libvrt), Cint, (Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                length(z), length(x), length(y), z, x, y, h))
    return Grid(h[], length(z) * length(x) * length(y))
end
# Λ_regular(ϵ, maxiter, atmos, line, quadrature, DATA; …) is Λ_voronoi above with `grid_of(sites)` replaced by
# `regular_grid(atmos)`, `sites.*` by `atmos.*` and the results reshaped to (nλ, nz, nx, ny) / (nz, nx, ny, 3).

# ---------------------------------------------------------------- site set-up and resampling (optional replacements)
# NeighbourMatrix straight from the positions instead of write_arrays + voro++ + the parser of read_cell
# (src/io.jl:8-40, rt_preprocessing/output_sites.cc, src/voronoi_utils.jl:42-70).  Same neighbour sets; the order inside a
# row differs from voro++'s, and smallest_angle is order-dependent (see DESIGN.md §2).
function voronoi_neighbours(positions::Matrix{<:Unitful.Length}, z_min, z_max, x_min, x_max, y_min, y_max)
    pos = Float64.(ustrip.(u"m", positions))
    n = size(pos, 2)
    b = Float64.(ustrip.(u"m", [z_min, z_max, x_min, x_max, y_min, y_max]))
    nbr = zeros(Int64, n, 64)
    need = Ref{Int64}(0)
    check(ccall((:vrt_voronoi_neighbours, libvrt), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int64, Ref{Int64}),
                n, pos, b, nbr, 64, need))
    return nbr[:, 1:need[]]
end

# trilinear over all sites (src/functions.jl:207-248); initialise (src/voronoi_utils.jl:687-708) is six of these
function trilinear(p_vec::Matrix{<:Unitful.Length}, atmos, vals::Array{Float64,3})
    pos = Float64.(ustrip.(u"m", p_vec))
    z = Float64.(ustrip.(u"m", atmos.z)); x = Float64.(ustrip.(u"m", atmos.x)); y = Float64.(ustrip.(u"m", atmos.y))
    out = Vector{Float64}(undef, size(pos, 2))
    check(ccall((:vrt_trilinear, libvrt), Cint,
                (Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                length(z), length(x), length(y), z, x, y, vals, size(pos, 2), pos, out))
    return out
end

# rejection_sampling(n_sites, atmos, quantity) (src/functions.jl:79-121) with a reproducible Philox stream
function rejection_sampling(n_sites::Int, atmos, quantity::Array{Float64,3}; seed::UInt64=UInt64(2022))
    z = Float64.(ustrip.(u"m", atmos.z)); x = Float64.(ustrip.(u"m", atmos.x)); y = Float64.(ustrip.(u"m", atmos.y))
    p_vec = Matrix{Float64}(undef, 3, n_sites)
    check(ccall((:vrt_rejection_sampling, libvrt), Cint,
                (Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}),
                n_sites, length(z), length(x), length(y), z, x, y, quantity, seed, p_vec, C_NULL))
    return p_vec * u"m"
end

# idx, dist = nn(KDTree(positions), p) for all raster points at once (src/voronoi_utils.jl:441-444)
function nearest_site(positions::Matrix{Float64}, bounds::Vector{Float64}, points::Matrix{Float64})
    m = size(points, 2)
    idx = Vector{Int64}(undef, m); dist = Vector{Float64}(undef, m)
    check(ccall((:vrt_nearest_site, libvrt), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                size(positions, 2), positions, bounds, m, points, idx, dist))
    return idx, dist
end

# ---------------------------------------------------------------- multi-GPU: collectives inside the library
# One Julia process per GPU (e.g. under MPI.jl or Distributed).  The host only ferries the 128-byte unique id: rank 0 of a group
# calls nccl_unique_id(), broadcasts the bytes with whatever it has (MPI.Bcast!, a socket, a file) and every member calls
# comm_init! on its solver; libvrt.so then runs ncclCommInitRank and all collectives of the Λ-iteration itself.
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:vrt_nccl_unique_id, libvrt), Cint, (Ptr{UInt8},), id))
    return id
end
function comm_init!(solver::Ptr{Cvoid}; dir_id=nothing, dir_rank=0, dir_size=1, lam_id=nothing, lam_rank=0, lam_size=1)
    check(ccall((:vrt_solver_comm_init, libvrt), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Ptr{UInt8}, Int32, Int32), solver,
                dir_id === nothing ? C_NULL : pointer(dir_id), dir_rank, dir_size, lam_id === nothing ? C_NULL : pointer(lam_id), lam_rank, lam_size))
end
# J reduced through peer memory (one node): every process exports the 64-byte CUDA IPC handle of its J buffer, the host gathers
# them in rank order (MPI.Allgather) and attaches them; vrt_lambda_iterate then fuses the reduction into the source update
function peer_handle(solver::Ptr{Cvoid})
    h = Vector{UInt8}(undef, 64)
    check(ccall((:vrt_solver_peer_handle, libvrt), Cint, (Ptr{Cvoid}, Ptr{UInt8}), solver, h))
    return h
end
peer_attach!(solver::Ptr{Cvoid}, handles::Vector{UInt8}) =
    check(ccall((:vrt_solver_peer_attach, libvrt), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32), solver, handles, length(handles) ÷ 64))
peer_detach!(solver::Ptr{Cvoid}) = check(ccall((:vrt_solver_peer_detach, libvrt), Cint, (Ptr{Cvoid},), solver))   # all ranks, then a barrier, then destroy
# direction `d` (0-based index in the solver's own table) on the local wavelengths [lo, hi) only: a direction shared with another
# process that takes the rest (20 directions on 8 GPUs: two whole directions and half of a ninth each)
set_direction_lambda!(solver::Ptr{Cvoid}, d::Integer, lo::Integer, hi::Integer) =
    check(ccall((:vrt_solver_set_direction_lambda, libvrt), Cint, (Ptr{Cvoid}, Int32, Int64, Int64), solver, d, lo, hi))
# this process's cells [first, last) in internal order (internal cell c is site perm_up[c+1]) and its slice of the state
function cell_slice(solver::Ptr{Cvoid})
    a = Ref{Int64}(0); b = Ref{Int64}(0)
    check(ccall((:vrt_solver_cell_slice, libvrt), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), solver, a, b))
    return a[], b[]
end
function get_state_slice!(solver::Ptr{Cvoid}, S::Matrix{Float64}, J::Matrix{Float64}, populations::Matrix{Float64})
    check(ccall((:vrt_get_state_slice, libvrt), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), solver, S, J, populations))
end
function set_state_slice!(solver::Ptr{Cvoid}, S::Matrix{Float64}, populations::Matrix{Float64})
    check(ccall((:vrt_set_state_slice, libvrt), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), solver, S, populations))
end
function state_checksum(solver::Ptr{Cvoid})
    out = zeros(Float64, 4)
    check(ccall((:vrt_state_checksum, libvrt), Cint, (Ptr{Cvoid}, Ptr{Float64}), solver, out))
    return out
end

# ---------------------------------------------------------------- output file (src/io.jl:57-225) without HDF5.jl
# create_output_file(output_path, nλ, n_sites, maxiter) -> handle; the write_to_file methods become write_dataset! calls and
# the per-iteration writes of S and the populations (src/lambda_iteration.jl:280-281) one write_state! from the device.
function create_output_file(output_path::String, nλ::Int, n_sites::Int, maxiter::Int)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vrt_output_create, libvrt), Cint, (Cstring, Int64, Int64, Int64, Ref{Ptr{Cvoid}}), output_path, nλ, n_sites, maxiter, h))
    return h[]
end
function create_output_file(output_path::String, nλ::Int, atmosphere_size::Tuple, maxiter::Int)
    nz, nx, ny = atmosphere_size
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vrt_output_create_regular, libvrt), Cint, (Cstring, Int64, Int64, Int64, Int64, Int64, Ref{Ptr{Cvoid}}),
                output_path, nλ, nz, nx, ny, maxiter, h))
    return h[]
end
write_dataset!(out::Ptr{Cvoid}, name::String, a::Array{Float64}) =
    check(ccall((:vrt_output_write, libvrt), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cvoid}, Int64), out, name, a, sizeof(a)))
write_dataset!(out::Ptr{Cvoid}, name::String, a::Array{Int64}) =
    check(ccall((:vrt_output_write, libvrt), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cvoid}, Int64), out, name, a, sizeof(a)))
write_convergence!(out::Ptr{Cvoid}, iteration::Int, difference::Float64) =
    check(ccall((:vrt_output_write_convergence, libvrt), Cint, (Ptr{Cvoid}, Int64, Float64), out, iteration, difference))
write_state!(out::Ptr{Cvoid}, solver::Ptr{Cvoid}) = check(ccall((:vrt_output_write_state, libvrt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), out, solver))
close_output(out::Ptr{Cvoid}) = check(ccall((:vrt_output_close, libvrt), Cint, (Ptr{Cvoid},), out))

function readdlm_quadrature(fname)
    rows = [parse.(Float64, split(l)) for l in eachline(fname) if !isempty(strip(l))]
    return permutedims(hcat(rows...))
end

end # module
