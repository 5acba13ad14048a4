# Times the UNMODIFIED reference (eschnett/RayTraceGR.jl) on the host CPU.  Not runnable in this
# repository's build image (no Julia); shipped so that anyone with Julia 1.4/1.5 and the reference's
# Manifest.toml can fill in the "Julia" row of BASELINE.md / DESIGN.md:
#
#   cd /path/to/RayTraceGR.jl && julia --project -t $(nproc) /path/to/run_julia_reference.jl
#
# Like the reference's own time.sh it runs example2() twice and reports the second timing (the first
# one includes compilation).  example2 = 200 x 200 rays, Kerr-Schild with M = 1, a = 0.
using RayTraceGR
println("threads = ", Threads.nthreads(), "  cpu threads = ", Sys.CPU_THREADS)
RayTraceGR.example2()
t = @elapsed RayTraceGR.example2()
println("example2: ", t, " s  -> ", 40000 / t, " rays/s")
RayTraceGR.example1()
t = @elapsed RayTraceGR.example1()
println("example1: ", t, " s  -> ", 40000 / t, " rays/s")
