# Self-contained check of julia/RayTraceGRCUDA.jl for whoever has Julia (>= 1.4) and a B200 -- NEITHER exists in
# the image this repository was built in, so this file, like the module, has never been executed there.
#
#     RTGR_LIBRARY=/path/to/libraytracegr_cuda.so julia julia/runtests.jl
#
# It runs the reference's two entry points (src/RayTraceGR.jl:542-612) through the CUDA library and compares the
# 8-bit images with the decoded golden images the reference ships (scenes/sphere.png, scenes/sphere2.png), which
# this repository keeps as raw arrays in tests/golden/*.npy (decoded by tests/golden/make_golden.py) -- so no PNG
# or image package is needed.  Bars: the ones of tests/test_oracle.py and tests/test_gpu_parity.py
# (sphere2: >= 99.9 % of the pixels bit-exact, census of the three colours classes; sphere: >= 99.6 %).
using Test

include(joinpath(@__DIR__, "RayTraceGRCUDA.jl"))
using .RayTraceGRCUDA

"Minimal reader of a version-1/2 .npy file holding a C-ordered uint8 array."
function read_npy_u8(path)
    open(path) do io
        magic = read(io, 6)
        @assert magic == UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59] "not an .npy file"
        major = read(io, UInt8); read(io, UInt8)
        hlen = major == 1 ? Int(ltoh(read(io, UInt16))) : Int(ltoh(read(io, UInt32)))
        header = String(read(io, hlen))
        @assert occursin("'|u1'", header) || occursin("'u1'", header) "expected uint8 data"
        @assert occursin("'fortran_order': False", header)
        m = match(r"'shape': \(([^)]*)\)", header)
        dims = [parse(Int, strip(s)) for s in split(m.captures[1], ",") if !isempty(strip(s))]
        data = read(io, prod(dims))
        # C order (row-major) -> Julia array indexed [channel, column, row]
        reshape(data, reverse(dims)...)
    end
end

golden(name) = read_npy_u8(joinpath(@__DIR__, "..", "tests", "golden", name))

"Fraction of pixels whose three 8-bit channels all agree; `img` is image8(canvas): (3, ni, nj)."
function agreement(img::Array{UInt8,3}, gold::Array{UInt8,3})
    @assert size(img) == size(gold) "image $(size(img)) vs golden $(size(gold))"
    same = 0
    for j in axes(img, 3), i in axes(img, 2)
        same += all(img[c, i, j] == gold[c, i, j] for c in 1:3)
    end
    same / (size(img, 2) * size(img, 3))
end

@testset "RayTraceGRCUDA" begin
    @testset "example2 reproduces scenes/sphere2.png" begin
        c = example2()
        @test isfile(joinpath("scenes", "sphere2.png"))
        img = RayTraceGRCUDA.image8(c)
        gold = golden("sphere2.npy")
        @test agreement(img, gold) >= 0.999
        # census of the reference image (tests/test_oracle.py): 31 338 rays end on the caelum, 5 154 on the frustum,
        # 3 508 on the sphere, none miss.  The frustum's colour is (0, 1/2, 0) * 2/3 = 8-bit (0, 85, 0).
        frustum = count(((i, j),) -> img[1, i, j] == 0x00 && img[2, i, j] == 0x55 && img[3, i, j] == 0x00,
                        [(i, j) for i in axes(img, 2), j in axes(img, 3)])
        @test abs(frustum - 5154) <= 40
        missed = count(((i, j),) -> img[1, i, j] == 0xff && img[2, i, j] == 0x00 && img[3, i, j] == 0x00,
                       [(i, j) for i in axes(img, 2), j in axes(img, 3)])
        @test missed == 0
    end
    @testset "example1 reproduces scenes/sphere.png up to the silhouette ring" begin
        c = example1()
        img = RayTraceGRCUDA.image8(c)
        @test agreement(img, golden("sphere.npy")) >= 0.996
    end
    @testset "screen_widths" begin
        wx, wy = screen_widths(90, 3840, 2160)            # BASELINE configs[3]: 90 degrees vertical at 16:9
        @test isapprox(wx[2], 32 / 9; atol=1e-14) && isapprox(wy[4], 2.0; atol=1e-14)
    end
    @testset "trace_rays is pure and trace_rays! works in place" begin
        m = kerr_schild
        objs = Object{Float64}[Sphere{Float64}((0, 0, 0, 0), (1, 0, 0, 0), -10), Plane{Float64}(-20),
                               Sphere{Float64}((0, 4, 0, 0), (1, 0, 0, 0), 0.5)]
        c0 = make_canvas(m, (0, 4, -2, 0), (0, 1, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0), 64, 48)
        before = copy(c0.pixels)
        c1 = trace_rays(m, objs, c0)
        @test c0.pixels == before
        @test any(p -> p.rgb != (0.0, 0.0, 0.0), c1.pixels)
        trace_rays!(m, objs, c0.pixels)
        @test c0.pixels == c1.pixels
    end
end
