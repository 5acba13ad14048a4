# RayTraceGRCUDA.jl -- Julia host side of libraytracegr_cuda.
#
# Keeps the scene API of eschnett/RayTraceGR.jl (src/RayTraceGR.jl): the metric choice
# `minkowski` / `kerr_schild` (src:262, :274; M and a are keyword-selectable here instead of
# the locals at src:275-276), the objects `Plane` / `Sphere` (src:394-413), `Pixel` / `Canvas` /
# `make_canvas` (src:446-478), `trace_rays(metric, objs, canvas)` (src:483-536) and the entry
# points `example1()` / `example2()` (src:542-612) which write scenes/sphere.png and
# scenes/sphere2.png.  Every computation is done by the CUDA library through `ccall`; there is
# no CPU fallback in this file -- if the library or a GPU is missing the calls throw.
#
# STATUS: written against include/raytracegr_cuda.h, NOT EXECUTED -- the build image of this
# repository has no Julia toolchain (see DESIGN.md "Host language").  The identical call
# sequence is exercised through ctypes by raytracegr.jl_b200/host.py and the GPU test-suite.
#
# Usage:
#   ENV["RAYTRACEGR_CUDA_LIB"] = "/path/to/libraytracegr_cuda.so"   # or put it on the loader path
#   include("RayTraceGRCUDA.jl"); using .RayTraceGRCUDA
#   RayTraceGRCUDA.example2()
module RayTraceGRCUDA

export minkowski, kerr_schild, Object, Plane, Sphere, Pixel, Canvas, make_canvas, screen_widths, trace_rays, trace_rays!, pin!, unpin!,
       user_metric, render, Frame, render!, participants!, example1, example2

const D = 4
const libpath = get(ENV, "RAYTRACEGR_CUDA_LIB", "libraytracegr_cuda")

# ---- C ABI mirrors (include/raytracegr_cuda.h) ------------------------------------------------
# All are isbits with the C layout, so Ref(x) / Vector{...} can be handed to ccall directly.

struct CObject            # rtgr_object, 88 bytes
    kind::Int32
    _pad::Int32
    time::Float64
    pos::NTuple{4,Float64}
    vel::NTuple{4,Float64}
    radius::Float64
end

struct CParams            # rtgr_params, 72 bytes
    metric::Int32
    r_formula::Int32
    M::Float64
    a::Float64
    lambda0::Float64
    lambda1::Float64
    reltol::Float64
    abstol::Float64
    hit_threshold::Float64
    interp_points::Int32
    maxiters::Int32
end

struct CCamera            # rtgr_camera, 136 bytes
    pos::NTuple{4,Float64}
    widthx::NTuple{4,Float64}
    widthy::NTuple{4,Float64}
    normal::NTuple{4,Float64}
    ni::Int32
    nj::Int32
end

struct Stats              # rtgr_stats, 56 bytes
    rays::UInt64
    rhs_evals::UInt64
    steps_accepted::UInt64
    steps_rejected::UInt64
    kernel_ms::Float64
    total_ms::Float64
    drain_ms::Float64
end

struct RtgrError <: Exception
    msg::String
end
Base.showerror(io::IO, e::RtgrError) = print(io, "libraytracegr_cuda: ", e.msg)

last_error() = unsafe_string(ccall((:rtgr_last_error, libpath), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : throw(RtgrError(last_error()))

# ---- context: device buffers and streams live behind an opaque handle -------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(devices::Vector{<:Integer}=Int[])
        h = Ref{Ptr{Cvoid}}(C_NULL)
        if isempty(devices)
            check(ccall((:rtgr_create, libpath), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), h, C_NULL, 0))
        else
            ids = Cint.(devices)
            check(ccall((:rtgr_create, libpath), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), h, ids, length(ids)))
        end
        ctx = new(h[])
        finalizer(close, ctx)
        ctx
    end
end
function Base.close(ctx::Context)
    if ctx.handle != C_NULL
        ccall((:rtgr_destroy, libpath), Cvoid, (Ptr{Cvoid},), ctx.handle)
        ctx.handle = C_NULL
    end
    nothing
end
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
"All visible GPUs of the box share one context; the tiles of a frame are dealt to them round-robin, each GPU feeds its warps from a dynamic queue."
context() = (default_ctx[] === nothing && (default_ctx[] = Context()); default_ctx[]::Context)

# ---- metrics ----------------------------------------------------------------------------------
# The reference passes the metric as a Julia function.  The two built-in ones are hand-specialised
# kernels, so here they are tag values; calling one with keywords picks M / a / the radius formula.
# Any OTHER metric is supplied as CUDA C++ source (the body of the reference's `metric(x)` written
# against a dual-number type with the same operator set) and compiled at run time: `user_metric`.
struct MetricTag
    kind::Int32       # 0 minkowski, 1 kerr_schild, >= 16 a metric compiled by user_metric
    M::Float64
    a::Float64
    r_formula::Int32  # 0 = radius line exactly as written at src:284 (parity), 1 = textbook Kerr-Schild
end
(m::MetricTag)(; M=m.M, a=m.a, r_formula=m.r_formula) = MetricTag(m.kind, M, a, r_formula)
const minkowski = MetricTag(0, 1.0, 0.0, 0)
const kerr_schild = MetricTag(1, 1.0, 0.0, 0)     # reference values M = 1, a = 0 (src:275-276)

"""
    user_metric(source::String; par=Float64[], ctx=context()) -> MetricTag

Compile `source` -- CUDA C++ defining

    template <class T> __device__ void rtgr_user_metric(const T x[4], T g[4][4], const double* par)

-- with NVRTC inside the library (rtgr_metric_compile) and return a tag usable wherever `minkowski`
or `kerr_schild` are (make_canvas, trace_rays, render).  The library differentiates through it with
forward-mode duals exactly as the reference's dmetric/christoffel do (src:298-331).  `par` (<= 16
numbers) reaches the function as its third argument.  Throws with the compiler's diagnostics if
the source does not compile.
"""
function user_metric(source::AbstractString; par=Float64[], ctx::Context=context())
    id = Ref{Int32}(-1)
    check(ccall((:rtgr_metric_compile, libpath), Cint, (Ptr{Cvoid}, Cstring, Ref{Int32}), ctx.handle, source, id))
    p = Float64.(collect(par))
    isempty(p) || check(ccall((:rtgr_metric_set_params, libpath), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Cint),
                              ctx.handle, id[], p, length(p)))
    MetricTag(id[], 1.0, 0.0, 0)
end

function cparams(m::MetricTag; tol=eps(Float64)^(3 / 4), λ0=0.0, λ1=100.0)
    CParams(m.kind, m.r_formula, m.M, m.a, λ0, λ1, tol, tol, 0.01, 10, 100000)
end

# ---- objects ----------------------------------------------------------------------------------
abstract type Object{T} end
struct Plane{T} <: Object{T}
    time::T
end
struct Sphere{T} <: Object{T}
    pos::NTuple{D,T}
    vel::NTuple{D,T}
    radius::T
    # accepts tuples, vectors or SVectors of any number type (an INNER constructor: an outer method with
    # this signature would replace the default one and then call itself)
    Sphere{T}(pos, vel, radius) where {T} = new{T}(NTuple{D,T}(Tuple(pos)), NTuple{D,T}(Tuple(vel)), T(radius))
end

const zero4 = (0.0, 0.0, 0.0, 0.0)
marshal(p::Plane) = CObject(0, 0, Float64(p.time), zero4, zero4, 0.0)
marshal(s::Sphere) = CObject(1, 0, 0.0, Float64.(s.pos), Float64.(s.vel), Float64(s.radius))
marshal(::Object) = error("Called distance on abstract object")
"Vector{Object{T}} holds boxed references; the C side wants a flat tagged array."
marshal(objs::AbstractVector{<:Object}) = CObject[marshal(o) for o in objs]

# ---- canvas -----------------------------------------------------------------------------------
struct Pixel{T}              # 11 numbers, isbits: identical to rtgr_pixel for T = Float64
    pos::NTuple{D,T}
    normal::NTuple{D,T}
    rgb::NTuple{3,T}
end
struct Canvas{T}
    pixels::Array{Pixel{T},2}   # ni x nj, column-major: linear index (i-1) + (j-1)*ni
end

"""
    screen_widths(view_angle_deg, ni, nj; x_dir=(0,1,0,0), y_dir=(0,0,0,1)) -> (widthx, widthy)

The two width vectors `make_canvas` (src:458-478) takes, for a screen with a VERTICAL view angle of `view_angle_deg`,
square pixels and a unit `normal` (the reference's README describes "a screen with a certain width and height, with a
view angle"): a pixel's ray direction is `normal + dx widthx + dy widthy` with dx, dy in (-1/2, 1/2), hence
|widthy| = 2 tan(angle/2) and |widthx| = |widthy| ni/nj.
"""
function screen_widths(view_angle_deg::Real, ni::Integer, nj::Integer; x_dir=(0, 1, 0, 0), y_dir=(0, 0, 0, 1))
    h = 2 * tand(view_angle_deg / 2)
    w = h * ni / nj
    (Tuple(w .* Float64.(collect(x_dir))), Tuple(h .* Float64.(collect(y_dir))))
end

function camera(pos, widthx, widthy, normal, ni::Integer, nj::Integer)
    t4(v) = NTuple{4,Float64}(Tuple(Float64.(collect(v))))
    CCamera(t4(pos), t4(widthx), t4(widthy), t4(normal), Int32(ni), Int32(nj))
end

"make_canvas(metric, pos, widthx, widthy, normal, ni, nj): the screen of src:458-478, built on the device."
function make_canvas(metric::MetricTag, pos, widthx, widthy, normal, ni::Int, nj::Int; ctx::Context=context())
    pixels = Array{Pixel{Float64}}(undef, ni, nj)
    check(ccall((:rtgr_make_canvas, libpath), Cint,
                (Ptr{Cvoid}, Ref{CParams}, Ref{CCamera}, Ptr{Pixel{Float64}}),
                ctx.handle, cparams(metric), camera(pos, widthx, widthy, normal, ni, nj), pixels))
    Canvas{Float64}(pixels)
end

"""
    trace_rays(metric, objs, c::Canvas)::Canvas

Drop-in for the reference's hot path (src:483-536): integrates one null geodesic per pixel until
it meets an object and colours the pixel.  Pure like the original: `c` is left untouched and a new
canvas is returned.  `stats` (optional `Ref{Stats}`) receives the work counters of the call.
"""
function trace_rays(metric::MetricTag, objs::AbstractVector{<:Object}, c::Canvas{Float64};
                    ctx::Context=context(), tol=eps(Float64)^(3 / 4), stats::Ref{Stats}=Ref{Stats}(),
                    pin::Bool=true)
    out = copy(c.pixels)
    trace_rays!(metric, objs, out; ctx=ctx, tol=tol, stats=stats, pin=pin)
    Canvas{Float64}(out)
end

"""
    trace_rays!(metric, objs, pixels::Array{Pixel{Float64},2}; pin=true)

In-place form: `pos`/`normal` of every pixel are read and `rgb` is written by the GPU directly in the
array's own memory (rtgr_trace_canvas).  With `pin = true` the array is page-locked for the duration of
the call (`rtgr_host_register`), which lets the kernel read and write it over PCIe while it computes
instead of staging copies; callers that trace the same array repeatedly should call `pin!`/`unpin!`
themselves once and pass `pin = false`.
"""
function trace_rays!(metric::MetricTag, objs::AbstractVector{<:Object}, pixels::Array{Pixel{Float64},2};
                     ctx::Context=context(), tol=eps(Float64)^(3 / 4), stats::Ref{Stats}=Ref{Stats}(),
                     pin::Bool=true, tile_offset::Integer=0, tile_stride::Integer=1)
    ni, nj = size(pixels)
    cobjs = marshal(objs)
    pinned = pin && pin!(pixels)
    try
        GC.@preserve pixels cobjs begin
            check(ccall((:rtgr_trace_canvas, libpath), Cint,
                        (Ptr{Cvoid}, Ref{CParams}, Ptr{CObject}, Cint, Ptr{Pixel{Float64}}, Cint, Cint, Cint, Cint,
                         Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ref{Stats}),
                        ctx.handle, cparams(metric; tol=tol), cobjs, length(cobjs), pixels, ni, nj,
                        tile_offset, tile_stride, C_NULL, C_NULL, C_NULL, C_NULL, stats))
        end
    finally
        pinned && unpin!(pixels)
    end
    pixels
end

"Page-lock the memory of `a` in place (rtgr_host_register); returns false if it already was."
function pin!(a::Array)
    ccall((:rtgr_host_is_pinned, libpath), Cint, (Ptr{Cvoid},), a) == 1 && return false
    check(ccall((:rtgr_host_register, libpath), Cint, (Ptr{Cvoid}, UInt64), a, sizeof(a)))
    true
end
unpin!(a::Array) = check(ccall((:rtgr_host_unregister, libpath), Cint, (Ptr{Cvoid},), a))

"""
    render(metric, objs, pos, widthx, widthy, normal, ni, nj) -> Array{UInt8,3} (3 x ni x nj)

Fused production path: make_canvas and the trace both run on the device and only the 8-bit image
comes back (memory order = row-major nj x ni x 3, i.e. the PNG of example1/2).
"""
function render(metric::MetricTag, objs::AbstractVector{<:Object}, pos, widthx, widthy, normal,
                ni::Int, nj::Int; ctx::Context=context(), tol=eps(Float64)^(3 / 4),
                stats::Ref{Stats}=Ref{Stats}())
    img = Array{UInt8}(undef, 3, ni, nj)
    cobjs = marshal(objs)
    check(ccall((:rtgr_render, libpath), Cint,
                (Ptr{Cvoid}, Ref{CParams}, Ptr{CObject}, Cint, Ref{CCamera}, Ptr{UInt8}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ref{Stats}),
                ctx.handle, cparams(metric; tol=tol), cobjs, length(cobjs),
                camera(pos, widthx, widthy, normal, ni, nj), img, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, stats))
    img
end

# ---- one frame shared by several GPUs: the cross-GPU dynamic tile queue ---------------------------
"""
    Frame(ni, nj; ctx)                 # owner: allocates queue + image in its GPU's memory
    Frame(ipc::Vector{UInt8}, ni, nj)  # another PROCESS (Distributed.jl / MPI.jl worker, one per GPU)

All participants draw 8x4-pixel patches of the frame from ONE queue head in the owner GPU's memory and
store their pixels straight into the owner's image over NVLink (rtgr_frame_*, see the header).  Every
participant calls `render!(frame, ...)` once per frame with the same scene; the caller puts a barrier
between consecutive frames and before `read(frame)`.
"""
mutable struct Frame
    handle::Ptr{Cvoid}
    ni::Int
    nj::Int
    ipc::Vector{UInt8}     # RTGR_IPC_HANDLE_BYTES = 64 bytes to hand to the other processes
    owner::Bool
end
function Frame(ni::Integer, nj::Integer; ctx::Context=context())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ipc = zeros(UInt8, 64)
    check(ccall((:rtgr_frame_create, libpath), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Ptr{Cvoid}}, Ptr{UInt8}),
                ctx.handle, ni, nj, h, ipc))
    Frame(h[], ni, nj, ipc, true)
end
function Frame(ipc::Vector{UInt8}, ni::Integer, nj::Integer; ctx::Context=context())
    length(ipc) == 64 || throw(ArgumentError("an IPC handle has 64 bytes"))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rtgr_frame_open, libpath), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Ref{Ptr{Cvoid}}),
                ctx.handle, ipc, ni, nj, h))
    Frame(h[], ni, nj, copy(ipc), false)
end
"This participant's share of the frame (rtgr_render_frame); `stats` receives its work counters."
function render!(f::Frame, metric::MetricTag, objs::AbstractVector{<:Object}, pos, widthx, widthy, normal;
                 tol=eps(Float64)^(3 / 4), stats::Ref{Stats}=Ref{Stats}())
    cobjs = marshal(objs)
    check(ccall((:rtgr_render_frame, libpath), Cint,
                (Ptr{Cvoid}, Ref{CParams}, Ptr{CObject}, Cint, Ref{CCamera}, Ref{Stats}),
                f.handle, cparams(metric; tol=tol), cobjs, length(cobjs),
                camera(pos, widthx, widthy, normal, f.ni, f.nj), stats))
    stats[]
end
"""
    trace_rays!(f::Frame, metric, objs, pixels::Array{Pixel{Float64},2})

`trace_rays` on ONE canvas shared by all participants of the frame (rtgr_trace_canvas_frame): the rays are drawn
from the frame's shared queue, `rgb` is written into `pixels` in place.  `pixels` must be the same physical
page-locked array in every participant -- the other devices of one context see it as is; separate processes map
the same shared memory (e.g. `Mmap.mmap` of a file in /dev/shm, or SharedArrays) and `pin!` their mapping.  When
all participants have returned (the caller's barrier) the canvas is complete: there is no gather step.
"""
function trace_rays!(f::Frame, metric::MetricTag, objs::AbstractVector{<:Object}, pixels::Array{Pixel{Float64},2};
                     tol=eps(Float64)^(3 / 4), stats::Ref{Stats}=Ref{Stats}())
    size(pixels) == (f.ni, f.nj) || throw(ArgumentError("the canvas is $(size(pixels)), the frame $(f.ni) x $(f.nj)"))
    cobjs = marshal(objs)
    GC.@preserve pixels cobjs begin
        check(ccall((:rtgr_trace_canvas_frame, libpath), Cint,
                    (Ptr{Cvoid}, Ref{CParams}, Ptr{CObject}, Cint, Ptr{Pixel{Float64}}, Cint, Cint, Ref{Stats}),
                    f.handle, cparams(metric; tol=tol), cobjs, length(cobjs), pixels, f.ni, f.nj, stats))
    end
    stats[]
end
"Tuning hint (rtgr_frame_set_participants): how many GPUs, all processes together, work on the frame."
participants!(f::Frame, n::Integer) = (check(ccall((:rtgr_frame_set_participants, libpath), Cint, (Ptr{Cvoid}, Cint), f.handle, n)); f)
"The image (3 x ni x nj UInt8, the memory order of the PNG) -- after the caller's barrier."
function Base.read(f::Frame)
    img = Array{UInt8}(undef, 3, f.ni, f.nj)
    check(ccall((:rtgr_frame_read, libpath), Cint, (Ptr{Cvoid}, Ptr{UInt8}), f.handle, img))
    img
end
function Base.close(f::Frame)
    if f.handle != C_NULL
        ccall((:rtgr_frame_close, libpath), Cvoid, (Ptr{Cvoid},), f.handle)
        f.handle = C_NULL
    end
    nothing
end

# ---- PNG output (dependency-free: stored deflate blocks) -----------------------------------------
const crc_table = let t = Vector{UInt32}(undef, 256)
    for n in 0:255
        c = UInt32(n)
        for _ in 1:8
            c = (c & 1) != 0 ? (0xedb88320 ⊻ (c >> 1)) : (c >> 1)
        end
        t[n+1] = c
    end
    t
end
function crc32(data::AbstractVector{UInt8}, crc::UInt32=0x00000000)
    c = ~crc
    for b in data
        c = crc_table[((c ⊻ b) & 0xff)+1] ⊻ (c >> 8)
    end
    ~c
end
function adler32(data::AbstractVector{UInt8})
    a, b = UInt32(1), UInt32(0)
    for x in data
        a = (a + x) % 65521
        b = (b + a) % 65521
    end
    (b << 16) | a
end
be32(x) = UInt8[(x>>24)&0xff, (x>>16)&0xff, (x>>8)&0xff, x&0xff]
function png_chunk(io::IO, tag::String, data::Vector{UInt8})
    body = vcat(Vector{UInt8}(tag), data)
    write(io, be32(UInt32(length(data))), body, be32(crc32(body)))
end
"Write a 3 x w x h UInt8 array (channel fastest, then column, then row) as an 8-bit RGB PNG."
function write_png(path::AbstractString, img::Array{UInt8,3})
    _, w, h = size(img)
    raw = UInt8[]
    sizehint!(raw, h * (3w + 1))
    for r in 1:h
        push!(raw, 0x00)                       # filter type 0
        append!(raw, vec(@view img[:, :, r]))
    end
    z = UInt8[0x78, 0x01]
    pos = 1
    while pos <= length(raw)
        n = min(65535, length(raw) - pos + 1)
        last = pos + n > length(raw)
        push!(z, last ? 0x01 : 0x00, n & 0xff, n >> 8, (~n) & 0xff, ((~n) >> 8) & 0xff)
        append!(z, @view raw[pos:pos+n-1])
        pos += n
    end
    append!(z, be32(adler32(raw)))
    open(path, "w") do io
        write(io, UInt8[0x89, 0x50, 0x4e, 0x47, 0x0d, 0x0a, 0x1a, 0x0a])
        png_chunk(io, "IHDR", vcat(be32(UInt32(w)), be32(UInt32(h)), UInt8[8, 2, 0, 0, 0]))
        png_chunk(io, "IDAT", z)
        png_chunk(io, "IEND", UInt8[])
    end
end

"8-bit image of a canvas as the reference saves it: row = j, column = i (the transposes at src:566-569), value = round(255 x)."
function image8(c::Canvas{Float64})
    ni, nj = size(c.pixels)
    img = Array{UInt8}(undef, 3, ni, nj)
    for j in 1:nj, i in 1:ni, k in 1:3
        img[k, i, j] = round(UInt8, 255 * clamp(c.pixels[i, j].rgb[k], 0.0, 1.0))
    end
    img
end

# ---- the reference's example entry points -------------------------------------------------------
const outdir = "scenes"

function run_example(metric::MetricTag, sphere_centre, campos, file::String)
    T = Float64
    caelum = Sphere{T}((0, 0, 0, 0), (1, 0, 0, 0), -10)       # sky: inside-out sphere of radius 10
    frustum = Plane{T}(-20)                                    # cut-off in coordinate time
    sphere = Sphere{T}(sphere_centre, (1, 0, 0, 0), T(1) / 2)
    objs = Object{T}[caelum, frustum, sphere]
    canvas = make_canvas(metric, campos, (0, 1, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0), 200, 200)
    canvas = trace_rays(metric, objs, canvas)
    mkpath(outdir)
    path = joinpath(outdir, file)
    rm(path, force=true)
    println("Output file is \"$path\"")
    write_png(path, image8(canvas))
    canvas
end

"Flat-space sphere scene (src:542-576) -> scenes/sphere.png"
example1() = run_example(minkowski, (0, 0, 0, 0), (0, 0, -2, 0), "sphere.png")
"Sphere next to the Kerr-Schild hole (src:578-612) -> scenes/sphere2.png"
example2() = run_example(kerr_schild, (0, 4, 0, 0), (0, 4, -2, 0), "sphere2.png")

end # module
