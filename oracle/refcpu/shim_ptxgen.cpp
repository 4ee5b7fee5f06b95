// ORACLE / reference build (test infrastructure only).
//
// cuj::gen::PTXGenerator without LLVM.  The reference turns every recorded cuj::Module into PTX through LLVM 11 + NVPTX
// (lib/cuj/cuj/src/gen/ptx.cpp), which this image does not have.  CUJ ships a second back end that needs no LLVM:
// CPPCodeGenerator (lib/cuj/cuj/src/gen/cpp.cpp) prints the same module as CUDA C++ text.  This file gives the reference's
// callers (wavefront.cpp:141-179 restated in refcpu.cpp, preview.cpp, env_sampler.cpp, tonemap.cpp) a PTXGenerator whose
// generate() = the reference's own CPPCodeGenerator + g++ for the host (prelude: cuda_on_host.h) + one launch trampoline per
// kernel, and whose "PTX" is a line naming the resulting shared object (loaded by shim_cuda.cpp's cuModuleLoadDataEx).
// Floating point: -O1 -ffp-contract=off, IEEE single precision, no fast-math: the reference's device build is fast-math
// (ptx.cpp:57-63), so this is the reference's ALGORITHM at full precision -- the same contract as the oracle.
#include <cuj/gen/cpp.h>
#include <cuj/gen/ptx.h>

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "refcpu_internal.h"

namespace {

std::string lib_dir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<void *>(&lib_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        auto k = p.rfind('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

std::string env_or(const char *name, const std::string &d) {
    const char *v = getenv(name);
    return v && *v ? v : d;
}

// object / module addresses inside generated symbol names (core/context.h:334-337) differ from run to run: rename them to
// ordinal ids so that the text -- and with it the cache key -- is a function of the scene only
std::string normalise_symbol_names(const std::string &text) {
    std::string out;
    out.reserve(text.size());
    std::map<std::string, int> ids;
    size_t i = 0;
    while (i < text.size()) {
        bool tagged = false;
        for (const char *tag : {"_of_object_0x", "_in_0x"}) {
            const size_t n = strlen(tag);
            if (text.compare(i, n, tag) == 0) {
                size_t j = i + n;
                while (j < text.size() && isxdigit((unsigned char)text[j])) ++j;
                const std::string addr = text.substr(i + n, j - i - n);
                auto it = ids.emplace(addr, (int)ids.size()).first;
                out.append(tag, n - 2);
                out += "n" + std::to_string(it->second);
                i = j;
                tagged = true;
                break;
            }
        }
        if (!tagged) out += text[i++];
    }
    return out;
}

uint64_t fnv1a(const std::string &s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

std::vector<std::string> split_args(const std::string &s) {
    std::vector<std::string> r;
    std::string cur;
    for (char c : s) {
        if (c == ',') { r.push_back(cur); cur.clear(); }
        else if (c != ' ' || !cur.empty()) cur += c;
    }
    while (!cur.empty() && cur.back() == ' ') cur.pop_back();
    if (!cur.empty()) r.push_back(cur);
    return r;
}

// one launcher per __global__ function: blocks on OpenMP threads, or (sequential) thread by thread with a hook
std::string make_trampolines(const std::string &text) {
    std::ostringstream o;
    const std::string key = "extern \"C\" __global__ CujTypeVoid ";
    size_t pos = 0;
    std::map<std::string, bool> done;
    while ((pos = text.find(key, pos)) != std::string::npos) {
        size_t b = pos + key.size();
        size_t lp = text.find('(', b), rp = text.find(')', lp);
        pos = rp;
        if (text.compare(rp, 2, ");") != 0) continue;  // definitions repeat the declaration without ';'
        const std::string name = text.substr(b, lp - b);
        if (done[name]) continue;
        done[name] = true;
        auto args = split_args(text.substr(lp + 1, rp - lp - 1));
        std::string call = name + "(";
        for (size_t i = 0; i < args.size(); ++i) {
            if (i) call += ", ";
            call += "*static_cast<" + args[i] + "*>(a[" + std::to_string(i) + "])";
        }
        call += ")";
        o << "extern \"C\" void refcpu_launch_" << name
          << "(void **a, const int *g, const int *b, int sequential, refcpu_thread_hook hook, void *user)\n{\n"
          << "    const int nblocks = g[0] * g[1] * g[2], nthreads = b[0] * b[1] * b[2];\n"
          << "    if(sequential)\n    {\n"
          << "        for(int bi = 0; bi < nblocks; ++bi) for(int ti = 0; ti < nthreads; ++ti)\n        {\n"
          << "            blockDim = { b[0], b[1], b[2] };\n"
          << "            blockIdx = { bi % g[0], (bi / g[0]) % g[1], bi / (g[0] * g[1]) };\n"
          << "            threadIdx = { ti % b[0], (ti / b[0]) % b[1], ti / (b[0] * b[1]) };\n"
          << "            if(hook) hook(bi * nthreads + ti, 0, user);\n"
          << "            " << call << ";\n"
          << "            if(hook) hook(bi * nthreads + ti, 1, user);\n"
          << "        }\n        return;\n    }\n"
          << "    #pragma omp parallel for schedule(dynamic, 1)\n"
          << "    for(int bi = 0; bi < nblocks; ++bi)\n    {\n"
          << "        blockDim = { b[0], b[1], b[2] };\n"
          << "        blockIdx = { bi % g[0], (bi / g[0]) % g[1], bi / (g[0] * g[1]) };\n"
          << "        for(int ti = 0; ti < nthreads; ++ti)\n        {\n"
          << "            threadIdx = { ti % b[0], (ti / b[0]) % b[1], ti / (b[0] * b[1]) };\n"
          << "            " << call << ";\n"
          << "        }\n    }\n}\n";
    }
    return o.str();
}

bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

}  // namespace

namespace cuj::gen {

void PTXGenerator::set_options(const Options &opts) { opts_ = opts; }

void PTXGenerator::generate(const dsl::Module &mod) {
    CPPCodeGenerator cpp;
    cpp.set_target(CPPCodeGenerator::Target::PTX);
    cpp.set_assert(false);
    cpp.generate(mod);
    std::string text = normalise_symbol_names(cpp.get_cpp_string());
    text = "#include \"cuda_on_host.h\"\n" + text + "\n" + make_trampolines(text);
    llvm_ir_ = text;  // get_llvm_ir(): the generated source, for inspection

    const std::string base = lib_dir();
    // generated text is reproducible only for scenes whose objects the reference does not order by address (its material / geometry /
    // medium ids follow std::map<RC<...>> pointer order, core/scene.cpp:112-158), so the cache mostly serves repeated opens of one
    // scene inside a process and the small fixed kernels (preview, tonemap); it lives outside the repo
    const std::string jit_dir = env_or("REFCPU_JIT_DIR", "/tmp/refcpu_jit_" + std::to_string(getuid()));
    const std::string inc_dir = env_or("REFCPU_PRELUDE_DIR", base + "/../refcpu");
    const std::string opt = env_or("REFCPU_JIT_OPT", "-O1");
    mkdir(jit_dir.c_str(), 0755);
    char name[64];
    snprintf(name, sizeof name, "k_%016llx", (unsigned long long)fnv1a(opt + text));
    const std::string so = jit_dir + "/" + name + ".so", src = jit_dir + "/" + name + ".cpp";
    if (!file_exists(so)) {
        { std::ofstream f(src); f << text; }
        const std::string tmp = so + ".tmp" + std::to_string(getpid());
        const std::string cmd = "/usr/bin/g++ -std=c++20 " + opt + " -ffp-contract=off -fno-strict-aliasing -fopenmp -fPIC -shared -w -I'" + inc_dir +
                                "' -o '" + tmp + "' '" + src + "' 2> '" + src + ".log'";
        if (getenv("REFCPU_VERBOSE")) fprintf(stderr, "refcpu: compiling %zu KB of generated C++: %s\n", text.size() / 1024, cmd.c_str());
        if (system(cmd.c_str()) != 0) throw std::runtime_error("refcpu: g++ failed on generated kernel source, see " + src + ".log");
        rename(tmp.c_str(), so.c_str());
        if (!getenv("REFCPU_KEEP_SRC")) { unlink(src.c_str()); unlink((src + ".log").c_str()); }
    }
    ptx_ = std::string(refcpu::JIT_TAG) + so + "\n";
}

const std::string &PTXGenerator::get_llvm_ir() const { return llvm_ir_; }
const std::string &PTXGenerator::get_ptx() const { return ptx_; }

}  // namespace cuj::gen
