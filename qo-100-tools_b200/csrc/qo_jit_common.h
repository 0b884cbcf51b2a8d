/*
 * qo_jit_common.h -- NVRTC plumbing shared by the run-time compiled kernels (qo_nodal_jit.h: the nodal plan as straight-line
 * code; qo_chain_jit.h: a cascade's element list folded into the chain kernel).  libnvrtc is dlopen()ed on first use; the
 * product never links it, and a box without it keeps the ahead-of-time kernels.
 */
#pragma once
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <vector>

/* ---- NVRTC, loaded on first use ------------------------------------------------------------------------------ */
struct QnNvrtc {
    void *h;
    nvrtcResult (*create)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*compile)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*log_size)(nvrtcProgram, size_t *);
    nvrtcResult (*log)(nvrtcProgram, char *);
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t *);
    nvrtcResult (*cubin)(nvrtcProgram, char *);
    nvrtcResult (*destroy)(nvrtcProgram *);
    const char *(*errstr)(nvrtcResult);
    nvrtcResult (*version)(int *, int *);
    int major, minor;
    std::string why;
};
static QnNvrtc *qn_nvrtc(void)
{
    static QnNvrtc N;
    static std::once_flag once;
    std::call_once(once, [] {
        /* QN_NVRTC_FIRST: a TU may name the library it wants tried first (qo_chain_jit.h: the toolkit's own NVRTC, whose PTX level
         * matches the ahead-of-time build; a process that imported torch already has torch's older copy under the bare soname) */
#ifndef QN_NVRTC_FIRST
#define QN_NVRTC_FIRST NULL
#endif
        const char *cands[] = { getenv("QO100NET_NVRTC"), QN_NVRTC_FIRST, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "libnvrtc.so.13" };
        for (const char *c : cands) {
            if (!c || !*c) continue;
            N.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (N.h) break;
            const char *e = dlerror();
            N.why += std::string(c) + ": " + (e ? e : "?") + "; ";
        }
        if (!N.h) return;
#define QN_SYM(field, name) *(void **)&N.field = dlsym(N.h, name); if (!N.field) { N.why = std::string("libnvrtc lacks ") + name; dlclose(N.h); N.h = NULL; return; }
        QN_SYM(create, "nvrtcCreateProgram") QN_SYM(compile, "nvrtcCompileProgram") QN_SYM(log_size, "nvrtcGetProgramLogSize")
        QN_SYM(log, "nvrtcGetProgramLog") QN_SYM(cubin_size, "nvrtcGetCUBINSize") QN_SYM(cubin, "nvrtcGetCUBIN")
        QN_SYM(destroy, "nvrtcDestroyProgram") QN_SYM(errstr, "nvrtcGetErrorString") QN_SYM(version, "nvrtcVersion")
#undef QN_SYM
        if (N.version(&N.major, &N.minor) != NVRTC_SUCCESS) N.major = N.minor = 0;
    });
    return &N;
}

/* source -> cubin for sm_100a; log receives the compiler's output (ptxas -v included) */
static bool qn_jit_compile(const std::string &src, std::vector<char> &cubin, std::string &log, const char *file_name = "qo_nodal_jit.cu")
{
    QnNvrtc *N = qn_nvrtc();
    if (!N->h) { log = "libnvrtc not loadable (" + N->why + ")"; return false; }
    nvrtcProgram pr = NULL;
    nvrtcResult r = N->create(&pr, src.c_str(), file_name, 0, NULL, NULL);
    if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCreateProgram: ") + N->errstr(r); return false; }
    std::vector<std::string> opt = { "--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--ptxas-options=-v" };
    std::vector<const char *> optv;
    for (auto &o : opt) optv.push_back(o.c_str());
    r = N->compile(pr, (int)optv.size(), optv.data());
    size_t ls = 0;
    if (N->log_size(pr, &ls) == NVRTC_SUCCESS && ls > 1) { log.resize(ls); N->log(pr, &log[0]); }
    if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCompileProgram: ") + N->errstr(r) + "\n" + log; N->destroy(&pr); return false; }
    size_t cs = 0;
    if (N->cubin_size(pr, &cs) != NVRTC_SUCCESS || cs == 0) { log += "\nno cubin"; N->destroy(&pr); return false; }
    cubin.resize(cs);
    N->cubin(pr, cubin.data());
    N->destroy(&pr);
    return true;
}

