// imx_jit.cuh — runtime specialisation of the hot kernels (host side, included by imx_api.cu).
//
// The ahead-of-time kernels handle every configuration with uniform branches on the mode flags;
// profiling shows those branches and the generic size arithmetic are most of the executed
// instructions.  For large batches the library therefore recompiles THE SAME kernel sources
// (csrc/*.cuh, read from the directory this shared object lives in) with NVRTC for sm_100a, with
// every flag and size injected as a literal (-DIMX_JIT -DIMX_K_<name>=<value>), and launches the
// resulting cubin through the driver API.  libnvrtc / libcuda are dlopen()ed: if either is missing,
// or the compile fails, the ahead-of-time kernels keep serving (imx_kernel_variant() tells which).
#pragma once

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace imxjit {

struct Api {
    bool tried = false, ok = false;
    void *h_nvrtc = nullptr, *h_cuda = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction);
    CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**);
};

struct Kernels {
    CUmodule mod = nullptr;
    CUfunction step = nullptr, rollout = nullptr, step_many = nullptr;
    int step_regs = 0, rollout_regs = 0;
    std::string log;
};

static Api g_api;
static std::mutex g_mu;
static std::map<std::string, Kernels> g_cache;
static std::string g_last_log;

template <typename F>
static bool sym(void* h, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(h, name));
    return out != nullptr;
}

static bool load_api() {
    if (g_api.tried) return g_api.ok;
    g_api.tried = true;
    const char* nv[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", nullptr};
    for (int k = 0; nv[k] && !g_api.h_nvrtc; ++k) g_api.h_nvrtc = dlopen(nv[k], RTLD_NOW | RTLD_LOCAL);
    g_api.h_cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!g_api.h_nvrtc || !g_api.h_cuda) { g_last_log = "libnvrtc.so.12 or libcuda.so.1 not loadable"; return false; }
    Api& a = g_api;
    bool ok = sym(a.h_nvrtc, "nvrtcCreateProgram", a.CreateProgram) && sym(a.h_nvrtc, "nvrtcCompileProgram", a.CompileProgram) &&
              sym(a.h_nvrtc, "nvrtcGetCUBINSize", a.GetCUBINSize) && sym(a.h_nvrtc, "nvrtcGetCUBIN", a.GetCUBIN) &&
              sym(a.h_nvrtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize) && sym(a.h_nvrtc, "nvrtcGetProgramLog", a.GetProgramLog) &&
              sym(a.h_nvrtc, "nvrtcAddNameExpression", a.AddNameExpression) && sym(a.h_nvrtc, "nvrtcGetLoweredName", a.GetLoweredName) &&
              sym(a.h_nvrtc, "nvrtcDestroyProgram", a.DestroyProgram) && sym(a.h_cuda, "cuModuleLoadData", a.ModuleLoadData) &&
              sym(a.h_cuda, "cuModuleGetFunction", a.ModuleGetFunction) && sym(a.h_cuda, "cuFuncSetAttribute", a.FuncSetAttribute) &&
              sym(a.h_cuda, "cuLaunchKernel", a.LaunchKernel) && sym(a.h_cuda, "cuFuncGetAttribute", a.FuncGetAttribute) &&
              sym(a.h_cuda, "cuLaunchKernelEx", a.LaunchKernelEx);
    if (!ok) g_last_log = "a required NVRTC / driver symbol is missing";
    g_api.ok = ok;
    return ok;
}

static std::string lib_dir() {
    Dl_info info;
    if (!dladdr(reinterpret_cast<void*>(&load_api), &info) || !info.dli_fname) return ".";
    std::string p(info.dli_fname);
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
}

static bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[65536];
    size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}

// Compiles (or fetches from the in-process cache) the kernels specialised by `defines`.
// step_name / rollout_name are C++ name expressions of the template instantiations.
static const Kernels* get(const std::vector<std::string>& defines, const std::string& step_name, const std::string& many_name,
                          const std::string& rollout_name, int step_smem_bytes, int many_smem_bytes, int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!load_api()) return nullptr;
    // a CUmodule belongs to the context it was loaded in: one cache entry per (device, specialisation)
    std::string key = "dev" + std::to_string(device) + "|" + step_name + "|" + rollout_name;
    for (const auto& d : defines) key += "|" + d;
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second.mod ? &it->second : nullptr;
    Kernels& K = g_cache[key];       // inserted empty: a failed build is remembered and not retried

    const std::string dir = lib_dir();
    const char* inc_names[] = {"../../include/imx_b200.h", "imx_device.cuh", "imx_step.cuh", "imx_step_tma.cuh", "imx_rollout.cuh"};
    const std::string inc_paths[] = {dir + "/../include/imx_b200.h", dir + "/csrc/imx_device.cuh", dir + "/csrc/imx_step.cuh",
                                     dir + "/csrc/imx_step_tma.cuh", dir + "/csrc/imx_rollout.cuh"};
    std::string inc_src[5];
    const char* inc_ptr[5];
    for (int k = 0; k < 5; ++k) {
        if (!read_file(inc_paths[k], inc_src[k])) { g_last_log = "kernel source not found: " + inc_paths[k]; return nullptr; }
        inc_ptr[k] = inc_src[k].c_str();
    }
    const char* tu = "#include \"imx_step_tma.cuh\"\n#include \"imx_rollout.cuh\"\n";
    nvrtcProgram prog;
    if (g_api.CreateProgram(&prog, tu, "imx_jit.cu", 5, inc_ptr, inc_names) != NVRTC_SUCCESS) { g_last_log = "nvrtcCreateProgram failed"; return nullptr; }
    g_api.AddNameExpression(prog, step_name.c_str());
    g_api.AddNameExpression(prog, many_name.c_str());
    g_api.AddNameExpression(prog, rollout_name.c_str());
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DIMX_JIT=1"};
    for (const auto& d : defines) opts.push_back("-D" + d);
    std::vector<const char*> optv;
    for (const auto& o : opts) optv.push_back(o.c_str());
    const nvrtcResult rc = g_api.CompileProgram(prog, (int)optv.size(), optv.data());
    size_t logn = 0;
    g_api.GetProgramLogSize(prog, &logn);
    if (logn > 1) { K.log.resize(logn); g_api.GetProgramLog(prog, &K.log[0]); }
    if (rc != NVRTC_SUCCESS) { g_last_log = "NVRTC compile failed: " + K.log.substr(0, 1500); g_api.DestroyProgram(&prog); return nullptr; }
    const char *low_step = nullptr, *low_roll = nullptr, *low_many = nullptr;
    g_api.GetLoweredName(prog, many_name.c_str(), &low_many);
    g_api.GetLoweredName(prog, step_name.c_str(), &low_step);
    g_api.GetLoweredName(prog, rollout_name.c_str(), &low_roll);
    size_t n = 0;
    g_api.GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    g_api.GetCUBIN(prog, cubin.data());
    CUmodule mod = nullptr;
    CUresult cr = g_api.ModuleLoadData(&mod, cubin.data());
    if (cr != CUDA_SUCCESS || !low_step || !low_roll || !low_many) { g_last_log = "cuModuleLoadData failed for the specialised cubin"; g_api.DestroyProgram(&prog); return nullptr; }
    CUfunction fs = nullptr, fr = nullptr, fm = nullptr;
    if (g_api.ModuleGetFunction(&fs, mod, low_step) != CUDA_SUCCESS || g_api.ModuleGetFunction(&fr, mod, low_roll) != CUDA_SUCCESS ||
        g_api.ModuleGetFunction(&fm, mod, low_many) != CUDA_SUCCESS) {
        g_last_log = "specialised kernel symbol not found";
        g_api.DestroyProgram(&prog);
        return nullptr;
    }
    g_api.DestroyProgram(&prog);
    g_api.FuncSetAttribute(fs, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, step_smem_bytes);
    g_api.FuncSetAttribute(fm, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, many_smem_bytes);
    g_api.FuncGetAttribute(&K.step_regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fs);
    g_api.FuncGetAttribute(&K.rollout_regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fr);
    K.mod = mod; K.step = fs; K.rollout = fr; K.step_many = fm;
    return &K;
}

// Compile only (no driver needed): used by the CPU-side build check to prove that the specialised
// translation unit compiles for sm_100a.  Returns cubin size or 0.
static size_t compile_only(const std::vector<std::string>& defines, const std::string& step_name, const std::string& many_name,
                           const std::string& rollout_name,
                           std::string& log, std::vector<char>* cubin_out) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_api.h_nvrtc) {
        const char* nv[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", nullptr};
        for (int k = 0; nv[k] && !g_api.h_nvrtc; ++k) g_api.h_nvrtc = dlopen(nv[k], RTLD_NOW | RTLD_LOCAL);
    }
    if (!g_api.h_nvrtc) { log = "libnvrtc not loadable"; return 0; }
    Api& a = g_api;
    if (!(sym(a.h_nvrtc, "nvrtcCreateProgram", a.CreateProgram) && sym(a.h_nvrtc, "nvrtcCompileProgram", a.CompileProgram) &&
          sym(a.h_nvrtc, "nvrtcGetCUBINSize", a.GetCUBINSize) && sym(a.h_nvrtc, "nvrtcGetCUBIN", a.GetCUBIN) &&
          sym(a.h_nvrtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize) && sym(a.h_nvrtc, "nvrtcGetProgramLog", a.GetProgramLog) &&
          sym(a.h_nvrtc, "nvrtcAddNameExpression", a.AddNameExpression) && sym(a.h_nvrtc, "nvrtcGetLoweredName", a.GetLoweredName) &&
          sym(a.h_nvrtc, "nvrtcDestroyProgram", a.DestroyProgram))) { log = "NVRTC symbol missing"; return 0; }
    const std::string dir = lib_dir();
    const char* inc_names[] = {"../../include/imx_b200.h", "imx_device.cuh", "imx_step.cuh", "imx_step_tma.cuh", "imx_rollout.cuh"};
    const std::string inc_paths[] = {dir + "/../include/imx_b200.h", dir + "/csrc/imx_device.cuh", dir + "/csrc/imx_step.cuh",
                                     dir + "/csrc/imx_step_tma.cuh", dir + "/csrc/imx_rollout.cuh"};
    std::string inc_src[5];
    const char* inc_ptr[5];
    for (int k = 0; k < 5; ++k) {
        if (!read_file(inc_paths[k], inc_src[k])) { log = "kernel source not found: " + inc_paths[k]; return 0; }
        inc_ptr[k] = inc_src[k].c_str();
    }
    const char* tu = "#include \"imx_step_tma.cuh\"\n#include \"imx_rollout.cuh\"\n";
    nvrtcProgram prog;
    if (a.CreateProgram(&prog, tu, "imx_jit.cu", 5, inc_ptr, inc_names) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return 0; }
    a.AddNameExpression(prog, step_name.c_str());
    a.AddNameExpression(prog, many_name.c_str());
    a.AddNameExpression(prog, rollout_name.c_str());
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DIMX_JIT=1"};
    for (const auto& d : defines) opts.push_back("-D" + d);
    std::vector<const char*> optv;
    for (const auto& o : opts) optv.push_back(o.c_str());
    const nvrtcResult rc = a.CompileProgram(prog, (int)optv.size(), optv.data());
    size_t logn = 0;
    a.GetProgramLogSize(prog, &logn);
    if (logn > 1) { log.resize(logn); a.GetProgramLog(prog, &log[0]); }
    size_t n = 0;
    if (rc == NVRTC_SUCCESS) {
        a.GetCUBINSize(prog, &n);
        if (cubin_out) { cubin_out->resize(n); a.GetCUBIN(prog, cubin_out->data()); }
    }
    a.DestroyProgram(&prog);
    return n;
}

}  // namespace imxjit
