// dynamic.cu -- user-supplied operators, the CUDA analogue of the reference's only plugin interface:
// `trait ShaderCommand { fn source() -> ShaderSource::SpirV; fn data(ShaderData) -> Descriptor }` with
// CommandBuffer::{construct,unary,binary}_dynamic (command/dynamic.rs:7-60, command.rs:2933-3060;
// tests/custom.rs).  The reference hands a SPIR-V fragment shader to wgpu; here the plugin is CUDA C
// source defining ONE device function,
//
//     __device__ float4 zos_shade(float2 uv, const unsigned char* params, zos_tex in0, zos_tex in1);
//
// evaluated once per destination pixel at uv = pixel centre / size, exactly like a full-screen
// fragment shader.  `params` is the invocation's binary data (ShaderData::set_data), `in0` / `in1` are
// the operands as sampled textures of working values (`in0.fetch(uv)`: nearest texel).  The source is
// compiled at program creation with NVRTC for sm_100a (cached per context by source text) -- the
// counterpart of the reference compiling its pipelines on first use (run.rs:2915-2939).
//
// Around the plugin the operator is built from the library's own kernels: operands are unpacked to
// RGBA32F working images, the plugin writes an RGBA32F image, and that is packed to the declared
// texel (for staged texels through the f16 attachment rounding, like every other draw).
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>

#include <cuda.h>

#include "zos_internal.h"

namespace zos {

namespace {

const char* kPrelude = R"ZOS(
struct zos_tex {
  const float4* p; int w, h; unsigned long long pitch;  // pitch in texels
  __device__ float4 at(int x, int y) const { x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1); return p[(unsigned long long)y * pitch + x]; }
  __device__ float4 fetch(float2 uv) const { return at((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
};
__device__ float4 zos_shade(float2 uv, const unsigned char* params, zos_tex in0, zos_tex in1);
extern "C" __global__ void zos_dynamic_entry(float4* out, int w, int h, unsigned long long pitch, zos_tex in0, zos_tex in1, const unsigned char* params) {
  const unsigned total = (unsigned)w * (unsigned)h;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int j = idx / (unsigned)w, i = idx - j * (unsigned)w;
    const float2 uv = make_float2(((float)i + 0.5f) / (float)w, ((float)j + 0.5f) / (float)h);
    out[(unsigned long long)j * pitch + i] = zos_shade(uv, params, in0, in1);
  }
}
#line 1 "zos_shade.cu"
)ZOS";

// ---- NVRTC through dlopen: the library has no link-time dependency on it
struct Nvrtc {
  void* lib = nullptr;
  int (*create)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*compile)(void*, int, const char* const*) = nullptr;
  int (*log_size)(void*, size_t*) = nullptr;
  int (*get_log)(void*, char*) = nullptr;
  int (*cubin_size)(void*, size_t*) = nullptr;
  int (*get_cubin)(void*, char*) = nullptr;
  int (*destroy)(void**) = nullptr;
  bool ok = false;
};
Nvrtc& nvrtc() {
  static Nvrtc n;
  static bool tried = false;
  if (tried) return n;
  tried = true;
  for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
    n.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (n.lib) break;
  }
  if (!n.lib) return n;
#define ZOS_SYM(field, sym) *(void**)(&n.field) = dlsym(n.lib, sym)
  ZOS_SYM(create, "nvrtcCreateProgram"); ZOS_SYM(compile, "nvrtcCompileProgram"); ZOS_SYM(log_size, "nvrtcGetProgramLogSize");
  ZOS_SYM(get_log, "nvrtcGetProgramLog"); ZOS_SYM(cubin_size, "nvrtcGetCUBINSize"); ZOS_SYM(get_cubin, "nvrtcGetCUBIN");
  ZOS_SYM(destroy, "nvrtcDestroyProgram");
#undef ZOS_SYM
  n.ok = n.create && n.compile && n.log_size && n.get_log && n.cubin_size && n.get_cubin && n.destroy;
  return n;
}

struct Driver {
  CUresult (*module_load)(CUmodule*, const void*) = nullptr;
  CUresult (*module_unload)(CUmodule) = nullptr;
  CUresult (*get_function)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  bool ok = false;
};
Driver& driver() {
  static Driver d;
  static bool tried = false;
  if (tried) return d;
  tried = true;
  auto get = [](const char* name) -> void* {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return fn;
  };
  *(void**)(&d.module_load) = get("cuModuleLoadData");
  *(void**)(&d.module_unload) = get("cuModuleUnload");
  *(void**)(&d.get_function) = get("cuModuleGetFunction");
  *(void**)(&d.launch) = get("cuLaunchKernel");
  d.ok = d.module_load && d.module_unload && d.get_function && d.launch;
  return d;
}

struct Tex { const float4* p; int w, h; unsigned long long pitch; };  // == zos_tex of the prelude

DevImage f32_image(void* ptr, int w, int h) {
  DevImage im;
  memset(&im, 0, sizeof im);
  im.p0 = (uint8_t*)ptr;
  im.w = w; im.h = h; im.bpp = 16;
  im.pitch = zos_aligned_row_stride((uint32_t)w, 16);
  im.block = ZOS_BLOCK_PIXEL;
  im.fmt = zos_texfmt{ZOS_TRANSFER_LINEAR, ZOS_PARTS_RGBA, ZOS_BITS_FLOAT32X4, ZOS_STORAGE_FLOAT};
  return im;
}

}  // namespace
}  // namespace zos

struct zos_dynamic {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
};

using namespace zos;

extern "C" {

// Compiles (or fetches from the context's cache) the plugin.  The compiler log is left in zos_last_error on failure.
zos_status zos_dynamic_create(zos_ctx* ctx, const char* cuda_source, zos_dynamic** out) {
  if (!ctx || !cuda_source || !out) return ZOS_ERR_INVALID;
  *out = nullptr;
  auto it = ctx->dynamic_cache.find(cuda_source);
  if (it != ctx->dynamic_cache.end()) { *out = it->second; return ZOS_OK; }
  Nvrtc& n = nvrtc();
  if (!n.ok) return fail(ctx, ZOS_ERR_UNSUPPORTED, "dynamic operators need libnvrtc.so.12 (not found)");
  cudaSetDevice(ctx->device);
  cudaFree(nullptr);  // make sure the primary context exists for the driver calls below
  Driver& d = driver();
  if (!d.ok) return fail(ctx, ZOS_ERR_UNSUPPORTED, "dynamic operators need the CUDA driver module API");
  std::string src = std::string(kPrelude) + cuda_source;
  void* prog = nullptr;
  if (n.create(&prog, src.c_str(), "zos_dynamic.cu", 0, nullptr, nullptr) != 0) return fail(ctx, ZOS_ERR_CUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-default-device"};
  const int rc = n.compile(prog, 4, opts);
  if (rc != 0) {
    size_t ls = 0;
    n.log_size(prog, &ls);
    std::string log(ls + 1, '\0');
    if (ls) n.get_log(prog, &log[0]);
    n.destroy(&prog);
    return fail(ctx, ZOS_ERR_INVALID, "dynamic operator does not compile (LaunchError): %.400s", log.c_str());
  }
  size_t cs = 0;
  n.cubin_size(prog, &cs);
  std::string cubin(cs, '\0');
  n.get_cubin(prog, &cubin[0]);
  n.destroy(&prog);
  zos_dynamic* dyn = new zos_dynamic();
  if (d.module_load(&dyn->mod, cubin.data()) != CUDA_SUCCESS || d.get_function(&dyn->fn, dyn->mod, "zos_dynamic_entry") != CUDA_SUCCESS) {
    if (dyn->mod) d.module_unload(dyn->mod);
    delete dyn;
    return fail(ctx, ZOS_ERR_CUDA, "cuModuleLoadData failed for the compiled dynamic operator");
  }
  ctx->dynamic_cache[cuda_source] = dyn;
  *out = dyn;
  return ZOS_OK;
}

// One frame: dst = shade(in0?, in1?, params).  Operands are unpacked to RGBA32F working images first, the
// result is packed into dst's texel afterwards (both with the library's own kernels).
zos_status zos_dynamic_launch(zos_ctx* ctx, zos_dynamic* dyn, const zos_image* dst, const zos_image* in0, const zos_image* in1,
                              const void* params, uint64_t params_len) {
  if (!ctx || !dyn || !dst) return ZOS_ERR_INVALID;
  if (params_len > 4096) return fail(ctx, ZOS_ERR_INVALID, "dynamic operator: at most 4096 bytes of parameters");
  Driver& d = driver();
  DevImage D, I[2];
  zos_status st;
  if ((st = make_dev_image(ctx, dst, &D, "dst")) != ZOS_OK) return st;
  const zos_image* ins[2] = {in0, in1};
  for (int k = 0; k < 2; k++)
    if (ins[k] && (st = make_dev_image(ctx, ins[k], &I[k], k ? "in1" : "in0")) != ZOS_OK) return st;
  if (D.block != ZOS_BLOCK_PIXEL || (in0 && I[0].block != ZOS_BLOCK_PIXEL) || (in1 && I[1].block != ZOS_BLOCK_PIXEL))
    return fail(ctx, ZOS_ERR_UNSUPPORTED, "dynamic operators take pixel images");
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  void* scratch[3] = {nullptr, nullptr, nullptr};
  void* dparams = nullptr;
  auto cleanup = [&]() {
    for (void* p : scratch) if (p) cudaFreeAsync(p, s);
    if (dparams) cudaFreeAsync(dparams, s);
  };
  DevImage out32 = f32_image(nullptr, D.w, D.h);
  if ((st = check_cuda(ctx, cudaMallocAsync(&scratch[0], out32.pitch * D.h, s), "cudaMallocAsync")) != ZOS_OK) { cleanup(); return st; }
  out32.p0 = (uint8_t*)scratch[0];
  Tex tex[2] = {{nullptr, 1, 1, 1}, {nullptr, 1, 1, 1}};
  for (int k = 0; k < 2; k++) {
    if (!ins[k]) continue;
    DevImage w32 = f32_image(nullptr, I[k].w, I[k].h);
    if ((st = check_cuda(ctx, cudaMallocAsync(&scratch[1 + k], w32.pitch * I[k].h, s), "cudaMallocAsync")) != ZOS_OK) { cleanup(); return st; }
    w32.p0 = (uint8_t*)scratch[1 + k];
    if ((st = launch_rowwise(ctx, &I[k], nullptr, w32, nullptr, nullptr, 0, 1)) != ZOS_OK) { cleanup(); return st; }
    tex[k] = Tex{(const float4*)w32.p0, I[k].w, I[k].h, (unsigned long long)(w32.pitch / 16)};
  }
  if ((st = check_cuda(ctx, cudaMallocAsync(&dparams, params_len ? params_len : 16, s), "cudaMallocAsync")) != ZOS_OK) { cleanup(); return st; }
  if (params_len && (st = check_cuda(ctx, cudaMemcpyAsync(dparams, params, params_len, cudaMemcpyHostToDevice, s), "dynamic params")) != ZOS_OK) { cleanup(); return st; }
  float4* outp = (float4*)out32.p0;
  int w = D.w, h = D.h;
  unsigned long long pitch = out32.pitch / 16;
  const unsigned char* pp = (const unsigned char*)dparams;
  void* args[] = {&outp, &w, &h, &pitch, &tex[0], &tex[1], &pp};
  const uint64_t total = (uint64_t)w * h;
  const unsigned grid = (unsigned)grid_for(ctx, total, 256, 32);
  if (d.launch(dyn->fn, grid, 1, 1, 256, 1, 1, 0, (CUstream)s, args, nullptr) != CUDA_SUCCESS) { cleanup(); return fail(ctx, ZOS_ERR_CUDA, "cuLaunchKernel failed for the dynamic operator"); }
  ctx->launches++;
  st = launch_rowwise(ctx, &out32, nullptr, D, nullptr, nullptr, 0, 1);
  cleanup();
  return st;
}

}  // extern "C"

namespace zos {
void destroy_dynamic_cache(zos_ctx* ctx) {
  Driver& d = driver();
  for (auto& kv : ctx->dynamic_cache) {
    if (kv.second->mod && d.ok) d.module_unload(kv.second->mod);
    delete kv.second;
  }
  ctx->dynamic_cache.clear();
}
}  // namespace zos
