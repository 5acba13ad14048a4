// Run-time compilation of a user-supplied metric: NVRTC turns the user's
// `rtgr_user_metric<T>` plus the library's own device headers (embedded as strings, see
// embed_sources.py) into an sm_100a cubin; the CUDA runtime loads it as a library
// (cudaLibraryLoadData) and the host side launches its three kernels in place of the built-in ones.
// libnvrtc is dlopen'ed on first use, so the library itself does not depend on it.
#pragma once
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>

#include <string>
#include <vector>

#include "rtgr_embedded.h"

namespace rtgr_jit {

struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

inline bool load_nvrtc(Nvrtc& n, std::string& err) {
    if (n.h) return true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (n.h) break; }
    if (!n.h) { err = std::string("cannot load libnvrtc (needed for user metrics): ") + dlerror(); return false; }
#define RTGR_SYM(field, sym)                                                              \
    *(void**)(&n.field) = dlsym(n.h, sym);                                                \
    if (!n.field) { err = std::string("libnvrtc lacks ") + sym; n.h = nullptr; return false; }
    RTGR_SYM(CreateProgram, "nvrtcCreateProgram")
    RTGR_SYM(CompileProgram, "nvrtcCompileProgram")
    RTGR_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    RTGR_SYM(GetCUBIN, "nvrtcGetCUBIN")
    RTGR_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    RTGR_SYM(GetProgramLog, "nvrtcGetProgramLog")
    RTGR_SYM(DestroyProgram, "nvrtcDestroyProgram")
    RTGR_SYM(GetErrorString, "nvrtcGetErrorString")
#undef RTGR_SYM
    return true;
}

// The translation unit handed to NVRTC: the library's kernels with METRIC = METRIC_USER around the
// user's function.  128 threads x 2 blocks/SM: the generic right-hand side carries 16 duals (80
// doubles) through the inverse and the contraction and wants the full 255 registers.
inline std::string program_source(const char* user_source) {
    std::string s;
    s += "#define RTGR_USER_METRIC 1\n#include \"rtgr_kernels.cuh\"\nnamespace rtgr_ad {\n#line 1 \"user_metric\"\n";
    s += user_source;
    s += "\n}  // namespace rtgr_ad\n"
         "extern \"C\" __global__ void __launch_bounds__(128, 2)\n"
         "rtgr_user_trace(rtgr::Job job, unsigned long long* next, unsigned long long* counters) {\n"
         "    rtgr_dev::trace_kernel_body<rtgr::METRIC_USER, 0>(job, next, counters);\n}\n"
         "extern \"C\" __global__ void __launch_bounds__(128, 2)\n"
         "rtgr_user_trace_stage(rtgr::Job job, unsigned long long* next, unsigned long long* counters) {\n"
         "    rtgr_dev::trace_kernel_body<rtgr::METRIC_USER, 0, false, true>(job, next, counters);\n}\n"
         "extern \"C\" __global__ void __launch_bounds__(128, 2)\n"
         "rtgr_user_trace_paths(rtgr::Job job, unsigned long long* next, unsigned long long* counters) {\n"
         "    rtgr_dev::trace_kernel_body<rtgr::METRIC_USER, 0, true>(job, next, counters);\n}\n"
         "extern \"C\" __global__ void rtgr_user_rhs(const double* states, long long n, double* derivs) {\n"
         "    rtgr_dev::rhs_kernel_body<rtgr::METRIC_USER, 0>(states, n, derivs);\n}\n"
         "extern \"C\" __global__ void rtgr_user_canvas(double* pixels) {\n"
         "    rtgr_dev::canvas_kernel_body<rtgr::METRIC_USER, 0>(pixels);\n}\n";
    return s;
}

// Compile to an sm_100a cubin.  Needs no GPU.  `log` receives the compiler's diagnostics.
inline bool compile(const char* user_source, std::vector<char>& cubin, std::string& log, std::string& err) {
    static Nvrtc nv;
    if (!user_source) { err = "metric source is NULL"; return false; }
    if (!load_nvrtc(nv, err)) return false;
    const std::string src = program_source(user_source);
    void* prog = nullptr;
    int rc = nv.CreateProgram(&prog, src.c_str(), "rtgr_user_metric.cu", rtgr_embedded_count, rtgr_embedded_sources,
                              rtgr_embedded_names);
    if (rc != 0) { err = std::string("nvrtcCreateProgram: ") + nv.GetErrorString(rc); return false; }
    // -default-device: unannotated functions (the ABI prototypes of raytracegr_cuda.h, and any helper the
    // user writes without __device__) are device functions
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
    rc = nv.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    if (nv.GetProgramLogSize(prog, &ls) == 0 && ls > 1) { log.resize(ls); nv.GetProgramLog(prog, &log[0]); }
    if (rc != 0) {
        err = std::string("metric source does not compile (") + nv.GetErrorString(rc) + "):\n" + log;
        nv.DestroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    if (nv.GetCUBINSize(prog, &cs) != 0 || cs == 0) { err = "nvrtcGetCUBINSize failed"; nv.DestroyProgram(&prog); return false; }
    cubin.resize(cs);
    nv.GetCUBIN(prog, cubin.data());
    nv.DestroyProgram(&prog);
    if (const char* dump = getenv("RTGR_JIT_DUMP")) {   // developer aid: keep the cubin for cuobjdump
        if (FILE* f = fopen(dump, "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return true;
}

}  // namespace rtgr_jit
