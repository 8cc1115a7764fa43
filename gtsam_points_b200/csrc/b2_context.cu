// b2_context.cu -- library state, error reporting, context (device + stream + staging).
#include <cstring>

#include "b2_internal.hpp"

namespace b2 {

static thread_local char g_last_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

b2_status fail(b2_status st, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return st;
}

}  // namespace b2

b2_status b2_ctx::ensure_stage(size_t host_bytes, size_t dev_bytes) {
  // the new buffer is allocated BEFORE the old one is released: a failed allocation leaves the context usable
  if (host_bytes > h_stage_bytes) {
    const size_t want = b2::round_up(host_bytes * 2, 4096);
    void* fresh = nullptr;
    B2_CUDA(cudaHostAlloc(&fresh, want, cudaHostAllocMapped | cudaHostAllocPortable));
    if (h_stage) {
      cudaStreamSynchronize(stream);  // nothing in flight may still read the old staging area
      cudaFreeHost(h_stage);
    }
    h_stage = fresh;
    h_stage_bytes = want;
  }
  if (dev_bytes > d_stage_bytes) {
    const size_t want = b2::round_up(dev_bytes * 2, 4096);
    void* fresh = nullptr;
    B2_CUDA(cudaMalloc(&fresh, want));
    if (d_stage) cudaFree(d_stage);
    d_stage = fresh;
    d_stage_bytes = want;
  }
  return B2_OK;
}

extern "C" {

const char* b2_last_error(void) { return b2::g_last_error; }
const char* b2_version(void) { return "b2points 0.1 (sm_100a)"; }

b2_status b2_ctx_create(int device, void* stream, b2_ctx** out) {
  B2_REQUIRE(out != nullptr, "b2_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    // No CPU fallback exists anywhere in this library: without a CUDA device nothing can run.
    return b2::fail(B2_ERR_NO_DEVICE, "b2_ctx_create: no CUDA device available (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
  }
  B2_REQUIRE(device >= 0 && device < count, "b2_ctx_create: device %d out of range [0, %d)", device, count);
  B2_CUDA(cudaSetDevice(device));
  b2_ctx* ctx = new b2_ctx;
  ctx->device = device;
  if (stream) {
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->owns_stream = false;
  } else {
    cudaError_t es = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (es != cudaSuccess) {
      delete ctx;
      return b2::fail(B2_ERR_CUDA, "cudaStreamCreateWithFlags: %s", cudaGetErrorString(es));
    }
    ctx->owns_stream = true;
  }
  cudaDeviceProp prop;
  cudaError_t ep = cudaGetDeviceProperties(&prop, device);
  if (ep != cudaSuccess) {
    delete ctx;
    return b2::fail(B2_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ep));
  }
  ctx->sm_count = prop.multiProcessorCount;
  {
    // completion word for zero-copy host calls: pinned + mapped, plus the device counter that decides who writes it
    void* h = nullptr;
    void* d = nullptr;
    cudaError_t ed = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (ed == cudaSuccess) ed = cudaHostGetDevicePointer(&d, h, 0);
    if (ed == cudaSuccess) ed = cudaMalloc(reinterpret_cast<void**>(&ctx->d_done_counter), sizeof(unsigned int));
    if (ed == cudaSuccess) ed = cudaMemset(ctx->d_done_counter, 0, sizeof(unsigned int));
    if (ed != cudaSuccess) {
      if (h) cudaFreeHost(h);
      if (ctx->d_done_counter) cudaFree(ctx->d_done_counter);
      if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
      delete ctx;
      return b2::fail(B2_ERR_CUDA, "b2_ctx_create: completion word: %s", cudaGetErrorString(ed));
    }
    *static_cast<volatile unsigned int*>(h) = 0u;
    ctx->h_done = static_cast<volatile unsigned int*>(h);
    ctx->d_done_flag = static_cast<volatile unsigned int*>(d);
  }
  *out = ctx;
  return B2_OK;
}

b2_status b2_ctx_destroy(b2_ctx* ctx) {
  if (!ctx) return B2_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->h_done) cudaFreeHost(const_cast<unsigned int*>(ctx->h_done));
  if (ctx->d_done_counter) cudaFree(ctx->d_done_counter);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B2_OK;
}

b2_status b2_ctx_synchronize(b2_ctx* ctx) {
  B2_REQUIRE(ctx != nullptr, "b2_ctx_synchronize: ctx is NULL");
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

void* b2_ctx_stream(b2_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

b2_status b2_device_malloc(b2_ctx* ctx, size_t bytes, void** out) {
  B2_REQUIRE(ctx && out, "b2_device_malloc: NULL argument");
  *out = nullptr;
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return B2_OK;
}

b2_status b2_device_free(b2_ctx* ctx, void* ptr) {
  B2_REQUIRE(ctx != nullptr, "b2_device_free: ctx is NULL");
  if (!ptr) return B2_OK;
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  B2_CUDA(cudaFree(ptr));
  return B2_OK;
}

b2_status b2_memcpy_d2h(b2_ctx* ctx, void* dst, const void* src, size_t bytes) {
  B2_REQUIRE(ctx && (bytes == 0 || (dst && src)), "b2_memcpy_d2h: NULL argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

b2_status b2_memcpy_h2d(b2_ctx* ctx, void* dst, const void* src, size_t bytes) {
  B2_REQUIRE(ctx && (bytes == 0 || (dst && src)), "b2_memcpy_h2d: NULL argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

}  // extern "C"
