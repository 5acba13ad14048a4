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

// What the author of a metric may declare in its source (see include/raytracegr_cuda.h):
//   `#pragma rtgr stationary`            g does not depend on x[0]: duals carry d/dx^1..3 only
//   a definition of rtgr_user_kerr_schild instead of rtgr_user_metric: g = eta + f k (x) k, closed-form right-hand side
inline bool declares_stationary(const char* src) { return std::string(src).find("#pragma rtgr stationary") != std::string::npos; }
inline bool declares_ks_form(const char* src) { return std::string(src).find("rtgr_user_kerr_schild") != std::string::npos; }

// The translation unit handed to NVRTC: the library's kernels with METRIC = METRIC_USER around the
// user's function.  128 threads x 2 blocks/SM: the generic right-hand side carries 16 duals (80
// doubles) through the inverse and the contraction and wants the full 255 registers; the Kerr-Schild
// form carries five duals and runs with 3 blocks/SM.
inline std::string program_source(const char* user_source) {
    std::string s;
    const bool ks = declares_ks_form(user_source);
    if (declares_stationary(user_source)) s += "#define RTGR_AD_NPART 3\n";
    if (ks) s += "#define RTGR_USER_KS_FORM 1\n";
    s += "#define RTGR_USER_METRIC 1\n#include \"rtgr_kernels.cuh\"\nnamespace rtgr_ad {\n#line 1 \"user_metric\"\n";
    s += user_source;
    s += "\n";
    if (ks)     // the 4x4 matrix itself (make_canvas needs it) from the user's f and k
        s += "template <class T> __device__ void rtgr_user_metric(const T x[4], T g[4][4], const double* par) {\n"
             "    T f, k[4];\n    rtgr_user_kerr_schild<T>(x, f, k, par);\n"
             "    for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q)\n"
             "        g[p][q] = ((p == q) ? (p == 0 ? -1.0 : 1.0) : 0.0) + f * k[p] * k[q];\n}\n";
    s += "}  // namespace rtgr_ad\n";
    const std::string lb = ks ? "__launch_bounds__(128, 3)" : "__launch_bounds__(128, 2)";
    const char* kernels[3][2] = {{"rtgr_user_trace", ""}, {"rtgr_user_trace_stage", ", false, true"}, {"rtgr_user_trace_paths", ", true"}};
    for (auto& kn : kernels)
        s += "extern \"C\" __global__ void " + lb + "\n" + kn[0] +
             "(rtgr::Job job, unsigned long long* next, unsigned long long* counters) {\n"
             "    rtgr_dev::trace_kernel_body<rtgr::METRIC_USER, 0" + kn[1] + ">(job, next, counters);\n}\n";
    s += "extern \"C\" __global__ void rtgr_user_rhs(const double* states, long long n, double* derivs) {\n"
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
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device", "-diag-suppress=161"};   // (161: "unrecognized #pragma" -- the rtgr declarations)
    rc = nv.CompileProgram(prog, 5, opts);
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
