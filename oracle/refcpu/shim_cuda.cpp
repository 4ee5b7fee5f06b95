// ORACLE / reference build (test infrastructure only; nothing under btrc_b200/ links or loads this).
//
// The subset of the CUDA runtime + driver API that the REFERENCE'S host code calls (utils/cuda/{buffer.h,array.cpp,texture.cpp,
// context.cpp,module.cpp}), implemented on HOST memory so that the reference's own scene set-up, film, SoA buffers and pipeline
// launch code run unmodified in a container without a GPU:
//   * cudaMalloc & co: bump allocation in one arena mapped at a FIXED address.  The reference bakes device pointers into its
//     generated kernels as immediates (cuj import_pointer); a deterministic arena makes the generated text reproducible, which
//     is what lets the compiled kernel library be cached across processes (shim_ptxgen.cpp).
//   * cudaArray / texture objects: texel storage + the sampling rule of oracle/orc_texture.h (the texture unit is vendor
//     arithmetic outside /root/reference; the same restatement of the CUDA programming guide's "Texture Fetching" appendix
//     serves the oracle -- parity at that boundary is unpinned for both, stated in DESIGN.md).
//   * cuLink* / cuModule* / cuLaunchKernel: "PTX" here is a line naming a host shared object built by shim_ptxgen.cpp; a launch
//     runs the kernel's generated trampoline (block loop on OpenMP threads, or thread by thread with a hook).
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <sys/mman.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../orc_texture.h"
#include "refcpu_internal.h"

namespace {

constexpr uintptr_t ARENA_BASE = 0x600000000000ull;
constexpr size_t ARENA_SIZE = 96ull << 30;  // virtual; pages are touched on demand
char *arena = nullptr;
size_t arena_used = 0;
std::mutex arena_mutex;

void arena_init() {
    if (arena) return;
    void *p = mmap(reinterpret_cast<void *>(ARENA_BASE), ARENA_SIZE, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_FIXED_NOREPLACE, -1, 0);
    if (p == MAP_FAILED)  // address taken (unlikely): any address still works, only the kernel cache loses its reproducibility
        p = mmap(nullptr, ARENA_SIZE, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("refcpu arena"); abort(); }
    arena = static_cast<char *>(p);
}

struct HostArray {
    cudaChannelFormatDesc desc;
    size_t w, h, d;
    std::vector<unsigned char> bytes;
    size_t texel_bytes() const { return (size_t)(desc.x + desc.y + desc.z + desc.w) / 8; }
    int channels() const { return (desc.x > 0) + (desc.y > 0) + (desc.z > 0) + (desc.w > 0); }
};

std::vector<orc::TexImage *> textures;  // texture object = index + 1
std::mutex tex_mutex;

int to_orc_mode(cudaTextureAddressMode m) {
    switch (m) {
    case cudaAddressModeWrap: return orc::ADDR_WRAP;
    case cudaAddressModeMirror: return orc::ADDR_MIRROR;
    case cudaAddressModeBorder: return orc::ADDR_BORDER;
    default: return orc::ADDR_CLAMP;
    }
}

struct LinkState { std::vector<std::string> inputs; std::string joined; };
struct HostModule { std::vector<void *> handles; };

refcpu::LaunchMode launch_mode;

}  // namespace

namespace refcpu {
void set_launch_mode(const LaunchMode &m) { launch_mode = m; }
void reset_arena_for_tests() { arena_used = 0; }
size_t arena_bytes_used() { return arena_used; }
void sample_tex2d(unsigned long long tex, float u, float v, float out[4]) { orc::sample_image(*textures[tex - 1], u, v, 0.0f, out); }
void sample_tex3d(unsigned long long tex, float u, float v, float w, float out[4]) { orc::sample_image(*textures[tex - 1], u, v, w, out); }
}  // namespace refcpu

extern "C" {

// ---- runtime API: memory ------------------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void **p, size_t bytes) {
    std::lock_guard<std::mutex> lk(arena_mutex);
    arena_init();
    size_t off = (arena_used + 255) & ~size_t(255);
    if (off + bytes > ARENA_SIZE) return cudaErrorMemoryAllocation;
    *p = arena + off;
    arena_used = off + (bytes ? bytes : 1);
    return cudaSuccess;
}
cudaError_t cudaFree(void *) { return cudaSuccess; }  // bump arena: released with the process
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "refcpu host shim: CUDA call failed"; }

// ---- runtime API: arrays and texture objects ------------------------------------------------------------------------------------
cudaChannelFormatDesc cudaCreateChannelDesc(int x, int y, int z, int w, cudaChannelFormatKind f) {
    cudaChannelFormatDesc d; d.x = x; d.y = y; d.z = z; d.w = w; d.f = f; return d;
}
cudaError_t cudaMallocArray(cudaArray_t *arr, const cudaChannelFormatDesc *desc, size_t w, size_t h, unsigned) {
    auto *a = new HostArray{*desc, w, h, 0, {}};
    a->bytes.resize(a->texel_bytes() * w * (h ? h : 1));
    *arr = reinterpret_cast<cudaArray_t>(a);
    return cudaSuccess;
}
cudaError_t cudaMalloc3DArray(cudaArray_t *arr, const cudaChannelFormatDesc *desc, cudaExtent e, unsigned) {
    auto *a = new HostArray{*desc, e.width, e.height, e.depth, {}};
    a->bytes.resize(a->texel_bytes() * e.width * e.height * e.depth);
    *arr = reinterpret_cast<cudaArray_t>(a);
    return cudaSuccess;
}
cudaError_t cudaFreeArray(cudaArray_t arr) { delete reinterpret_cast<HostArray *>(arr); return cudaSuccess; }
cudaError_t cudaMemcpy2DToArray(cudaArray_t dst, size_t wo, size_t ho, const void *src, size_t spitch, size_t width_bytes, size_t height, cudaMemcpyKind) {
    auto *a = reinterpret_cast<HostArray *>(dst);
    if (wo || ho) return cudaErrorInvalidValue;
    const size_t row = a->texel_bytes() * a->w;
    for (size_t y = 0; y < height; ++y) memcpy(a->bytes.data() + y * row, static_cast<const char *>(src) + y * spitch, width_bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms *p) {
    auto *a = reinterpret_cast<HostArray *>(p->dstArray);
    if (!a || !p->srcPtr.ptr) return cudaErrorInvalidValue;
    const size_t row = a->texel_bytes() * a->w;
    for (size_t z = 0; z < p->extent.depth; ++z)
        for (size_t y = 0; y < p->extent.height; ++y)
            memcpy(a->bytes.data() + (z * a->h + y) * row,
                   static_cast<const char *>(p->srcPtr.ptr) + (z * p->srcPtr.ysize + y) * p->srcPtr.pitch, row);
    return cudaSuccess;
}

cudaError_t cudaCreateTextureObject(cudaTextureObject_t *out, const cudaResourceDesc *res, const cudaTextureDesc *td, const cudaResourceViewDesc *) {
    if (res->resType != cudaResourceTypeArray) return cudaErrorInvalidValue;
    auto *a = reinterpret_cast<HostArray *>(res->res.array.array);
    auto *t = new orc::TexImage();
    t->w = (int)a->w; t->h = (int)(a->h ? a->h : 1); t->d = (int)a->d;
    const int nch = a->channels();
    if (nch != 1 && nch != 4) return cudaErrorInvalidValue;
    t->channels = nch;
    for (int i = 0; i < 3; ++i) t->addr[i] = to_orc_mode(td->addressMode[i]);
    t->linear = td->filterMode == cudaFilterModeLinear;
    for (int i = 0; i < 4; ++i) t->border[i] = td->borderColor[i];
    const size_t n = a->w * (a->h ? a->h : 1) * (a->d ? a->d : 1) * nch;
    t->data.resize(n);
    const int bits = a->desc.x;
    t->border_levels = a->desc.f == cudaChannelFormatKindFloat ? 0 : (a->desc.f == cudaChannelFormatKindSigned ? 127 : 255);
    if (a->desc.f == cudaChannelFormatKindFloat && bits == 32) {
        memcpy(t->data.data(), a->bytes.data(), n * 4);
    } else if (bits == 8 && td->readMode == cudaReadModeNormalizedFloat && a->desc.f == cudaChannelFormatKindUnsigned) {
        for (size_t i = 0; i < n; ++i) {
            float v = a->bytes[i] / 255.0f;
            if (td->sRGB && (nch == 1 || (i & 3) != 3)) v = orc::srgb_to_linear(v);  // alpha stays linear
            t->data[i] = v;
        }
    } else if (bits == 8 && td->readMode == cudaReadModeNormalizedFloat && a->desc.f == cudaChannelFormatKindSigned) {
        for (size_t i = 0; i < n; ++i) {  // SNorm8: max(c / 127, -1)
            float v = (float)(signed char)a->bytes[i] / 127.0f;
            t->data[i] = v < -1.0f ? -1.0f : v;
        }
    } else {
        delete t;
        return cudaErrorInvalidValue;
    }
    std::lock_guard<std::mutex> lk(tex_mutex);
    textures.push_back(t);
    *out = textures.size();
    return cudaSuccess;
}
cudaError_t cudaDestroyTextureObject(cudaTextureObject_t) { return cudaSuccess; }

// ---- driver API: context, "PTX" modules, launches ---------------------------------------------------------------------------------
CUresult cuInit(unsigned) { return CUDA_SUCCESS; }
CUresult cuDeviceGet(CUdevice *d, int ordinal) { *d = ordinal; return CUDA_SUCCESS; }
CUresult cuCtxCreate(CUcontext *c, unsigned, CUdevice) { *c = reinterpret_cast<CUcontext>(uintptr_t(1)); return CUDA_SUCCESS; }
CUresult cuCtxDestroy(CUcontext) { return CUDA_SUCCESS; }
CUresult cuCtxGetCurrent(CUcontext *c) { *c = reinterpret_cast<CUcontext>(uintptr_t(1)); return CUDA_SUCCESS; }
CUresult cuCtxSetCurrent(CUcontext) { return CUDA_SUCCESS; }
CUresult cuGetErrorName(CUresult, const char **s) { *s = "REFCPU_ERROR"; return CUDA_SUCCESS; }
CUresult cuGetErrorString(CUresult, const char **s) { *s = "refcpu host shim: driver call failed"; return CUDA_SUCCESS; }

CUresult cuLinkCreate(unsigned, CUjit_option *, void **, CUlinkState *out) { *out = reinterpret_cast<CUlinkState>(new LinkState()); return CUDA_SUCCESS; }
CUresult cuLinkAddData(CUlinkState st, CUjitInputType, void *data, size_t size, const char *, unsigned, CUjit_option *, void **) {
    reinterpret_cast<LinkState *>(st)->inputs.emplace_back(static_cast<const char *>(data), size);
    return CUDA_SUCCESS;
}
CUresult cuLinkComplete(CUlinkState st, void **cubin, size_t *size) {
    auto *ls = reinterpret_cast<LinkState *>(st);
    ls->joined.clear();
    for (auto &s : ls->inputs) { ls->joined += s; ls->joined += '\n'; }
    *cubin = ls->joined.data(); *size = ls->joined.size();
    return CUDA_SUCCESS;
}
CUresult cuLinkDestroy(CUlinkState st) { delete reinterpret_cast<LinkState *>(st); return CUDA_SUCCESS; }

CUresult cuModuleLoadDataEx(CUmodule *out, const void *image, unsigned, CUjit_option *, void **) {
    auto *m = new HostModule();
    std::string text(static_cast<const char *>(image));
    size_t pos = 0;
    const std::string tag = refcpu::JIT_TAG;
    while ((pos = text.find(tag, pos)) != std::string::npos) {
        size_t b = pos + tag.size(), e = text.find('\n', b);
        std::string path = text.substr(b, e == std::string::npos ? std::string::npos : e - b);
        pos = b;
        void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) { fprintf(stderr, "refcpu: dlopen(%s): %s\n", path.c_str(), dlerror()); delete m; return CUDA_ERROR_INVALID_IMAGE; }
        using SetCb = void (*)(const refcpu::Callbacks *);
        if (auto set = reinterpret_cast<SetCb>(dlsym(h, "refcpu_set_callbacks"))) {
            refcpu::Callbacks cb{&refcpu::sample_tex2d, &refcpu::sample_tex3d};
            set(&cb);
        }
        m->handles.push_back(h);
    }
    if (m->handles.empty()) { delete m; return CUDA_ERROR_INVALID_IMAGE; }
    *out = reinterpret_cast<CUmodule>(m);
    return CUDA_SUCCESS;
}
CUresult cuModuleUnload(CUmodule) { return CUDA_SUCCESS; }  // handles stay open: kernels may be cached in several modules
CUresult cuModuleGetFunction(CUfunction *f, CUmodule mod, const char *name) {
    auto *m = reinterpret_cast<HostModule *>(mod);
    const std::string sym = std::string("refcpu_launch_") + name;
    for (void *h : m->handles)
        if (void *p = dlsym(h, sym.c_str())) { *f = reinterpret_cast<CUfunction>(p); return CUDA_SUCCESS; }
    return CUDA_ERROR_NOT_FOUND;
}
CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned, CUstream, void **args, void **) {
    auto fn = reinterpret_cast<refcpu::LaunchFn>(f);
    const int grid[3] = {(int)gx, (int)gy, (int)gz}, block[3] = {(int)bx, (int)by, (int)bz};
    fn(args, grid, block, launch_mode.sequential ? 1 : 0, launch_mode.hook, launch_mode.user);
    return CUDA_SUCCESS;
}

}  // extern "C"
