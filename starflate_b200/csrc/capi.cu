// C ABI of the B200-native DEFLATE decompressor (include/starflate_b200.h).
// Host-side plumbing only: context, launch geometry, staging copies.  No CPU decode path
// exists anywhere in this library: without a CUDA device every call fails.
#include "../../include/starflate_b200.h"
#include "huff_lanes.cuh"
#include "lz_warp.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

namespace {

#ifndef SFB_ROOT_LIT
#define SFB_ROOT_LIT 8
#define SFB_ROOT_DIST 6
#define SFB_POOL 96
#define SFB_WARPS 4
#endif
#ifndef SFB_LZ_CTAS_PER_SM
#define SFB_LZ_CTAS_PER_SM 0  /* 0 = as many as fit */
#endif
using LaneCfg = sfb::Cfg<SFB_ROOT_LIT, SFB_ROOT_DIST, SFB_POOL, SFB_WARPS>;

}  // namespace

struct sfb200_ctx {
  int device = -1;
  int sm_count = 0;
  int ctas_per_sm = 0;
  int regs_per_thread = 0;
  int lz_ctas_per_sm = 0;
  int lz_regs_per_thread = 0;
  unsigned long long* d_counter = nullptr;  // [0] pass-1 group counter, [1] pass-2 stream counter
  uint32_t* d_lens = nullptr;  // per-resident-lane code-length scratch
  uint32_t* d_bits = nullptr;  // match-head bitmap: 1 bit per dst byte (grown on demand)
  uint64_t d_bits_words = 0;
  uint64_t* d_written = nullptr;  // pass-1 -> pass-2 sizes when the caller passes written == NULL
  uint64_t d_written_n = 0;
  uint64_t launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // call start | pass 1 start | pass 1 end | pass 2 end
  bool ev_valid = false;
  std::string err;
  // staging for the host-buffer entry points (grown on demand)
  uint8_t* d_src = nullptr;
  uint64_t d_src_cap = 0;
  uint8_t* d_dst = nullptr;
  uint64_t d_dst_cap = 0;
  uint64_t* d_meta = nullptr;  // src_off, src_len, dst_off, dst_cap, written : 5*n u64, then status n u8
  uint64_t d_meta_n = 0;
};

namespace {

int fail(sfb200_ctx* ctx, cudaError_t e, const char* what)
{
  if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? SFB200_RC_OUT_OF_MEMORY : SFB200_RC_CUDA_ERROR;
}

#define SFB_TRY(ctx, call)                                   \
  do {                                                       \
    const cudaError_t e_ = (call);                           \
    if (e_ != cudaSuccess) return fail((ctx), e_, #call);    \
  } while (0)

int grow(sfb200_ctx* ctx, uint8_t** p, uint64_t* cap, uint64_t need)
{
  if (need <= *cap) return SFB200_RC_OK;
  if (*p) SFB_TRY(ctx, cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  const uint64_t want = need + need / 8 + 256;
  SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(p), want));
  *cap = want;
  return SFB200_RC_OK;
}

}  // namespace

extern "C" {

void sfb200_destroy(sfb200_ctx* ctx);

int sfb200_abi_version(void) { return SFB200_ABI_VERSION; }

int sfb200_create(int device, sfb200_ctx** out)
{
  if (!out) return SFB200_RC_BAD_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return SFB200_RC_NO_DEVICE;
  }
  if (device < 0 || device >= count) return SFB200_RC_BAD_ARGUMENT;
  sfb200_ctx* ctx = new (std::nothrow) sfb200_ctx;
  if (!ctx) return SFB200_RC_OUT_OF_MEMORY;
  ctx->device = device;
  auto bail = [&](int rc) {
    delete ctx;
    return rc;
  };
  if (cudaSetDevice(device) != cudaSuccess) return bail(SFB200_RC_NO_DEVICE);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(SFB200_RC_CUDA_ERROR);
  ctx->sm_count = prop.multiProcessorCount;
  auto kern = sfb::huff_lanes_kernel<LaneCfg>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           LaneCfg::SMEM_BYTES) != cudaSuccess)
    return bail(SFB200_RC_CUDA_ERROR);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LaneCfg::WARPS * 32,
                                                    LaneCfg::SMEM_BYTES) != cudaSuccess ||
      per_sm < 1)
    return bail(SFB200_RC_CUDA_ERROR);
  ctx->ctas_per_sm = per_sm;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) ctx->regs_per_thread = fa.numRegs;
  {
    int lz_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lz_per_sm, sfb::lz_resolve_kernel,
                                                      sfb::LZ_THREADS, 0) != cudaSuccess ||
        lz_per_sm < 1)
      return bail(SFB200_RC_CUDA_ERROR);
    // fewer resident warps keep the streams' 32 KiB windows inside the L2
    int cap = SFB_LZ_CTAS_PER_SM;
    if (const char* e = std::getenv("SFB200_LZ_CTAS_PER_SM")) cap = std::atoi(e);
    ctx->lz_ctas_per_sm = (cap > 0 && cap < lz_per_sm) ? cap : lz_per_sm;
    cudaFuncAttributes lfa;
    if (cudaFuncGetAttributes(&lfa, sfb::lz_resolve_kernel) == cudaSuccess)
      ctx->lz_regs_per_thread = lfa.numRegs;
  }
  if (cudaMalloc(reinterpret_cast<void**>(&ctx->d_counter), 2 * sizeof(unsigned long long)) !=
      cudaSuccess)
    return bail(SFB200_RC_OUT_OF_MEMORY);
  const size_t lens_bytes = static_cast<size_t>(ctx->sm_count) * static_cast<size_t>(per_sm) *
                            LaneCfg::WARPS * sfb::SCRATCH_WORDS * 32 * sizeof(uint32_t);
  if (cudaMalloc(reinterpret_cast<void**>(&ctx->d_lens), lens_bytes) != cudaSuccess) {
    cudaFree(ctx->d_counter);
    return bail(SFB200_RC_OUT_OF_MEMORY);
  }
  for (auto& e : ctx->ev)
    if (cudaEventCreate(&e) != cudaSuccess) {
      sfb200_destroy(ctx);
      return SFB200_RC_CUDA_ERROR;
    }
  *out = ctx;
  return SFB200_RC_OK;
}

void sfb200_destroy(sfb200_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_counter);
  cudaFree(ctx->d_lens);
  cudaFree(ctx->d_bits);
  cudaFree(ctx->d_written);
  cudaFree(ctx->d_src);
  cudaFree(ctx->d_dst);
  cudaFree(ctx->d_meta);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  delete ctx;
}

const char* sfb200_last_error(const sfb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sfb200_get_launch_info(sfb200_ctx* ctx, sfb200_launch_info* out)
{
  if (!ctx || !out) return SFB200_RC_BAD_ARGUMENT;
  out->sm_count = ctx->sm_count;
  out->warps_per_cta = LaneCfg::WARPS;
  out->ctas_per_sm = ctx->ctas_per_sm;
  out->smem_bytes_per_cta = LaneCfg::SMEM_BYTES;
  out->regs_per_thread = ctx->regs_per_thread;
  out->lz_threads_per_cta = sfb::LZ_THREADS;
  out->lz_ctas_per_sm = ctx->lz_ctas_per_sm;
  out->lz_regs_per_thread = ctx->lz_regs_per_thread;
  out->kernel_launches = ctx->launches;
  return SFB200_RC_OK;
}

int sfb200_decompress_batch_device(sfb200_ctx* ctx, const uint8_t* src_base,
                                   const uint64_t* src_off, const uint64_t* src_len,
                                   uint8_t* dst_base, uint64_t dst_bytes, const uint64_t* dst_off,
                                   const uint64_t* dst_cap, uint8_t* status,
                                   uint64_t* written, uint64_t n, void* cuda_stream)
{
  if (!ctx) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !status) return SFB200_RC_BAD_ARGUMENT;
  SFB_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // scratch: one bit per dst byte (+ a guard word), grown on demand and kept
  const uint64_t bits_words = dst_bytes / 32 + 2;
  if (bits_words > ctx->d_bits_words) {
    if (ctx->d_bits) SFB_TRY(ctx, cudaFree(ctx->d_bits));
    ctx->d_bits = nullptr;
    ctx->d_bits_words = 0;
    const uint64_t want = bits_words + bits_words / 8 + 64;
    SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_bits), want * sizeof(uint32_t)));
    ctx->d_bits_words = want;
  }
  if (!written) {
    if (n > ctx->d_written_n) {
      if (ctx->d_written) SFB_TRY(ctx, cudaFree(ctx->d_written));
      ctx->d_written = nullptr;
      ctx->d_written_n = 0;
      const uint64_t want = n + n / 8 + 64;
      SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_written), want * sizeof(uint64_t)));
      ctx->d_written_n = want;
    }
    written = ctx->d_written;
  }
  ctx->ev_valid = false;
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_counter, 0, 2 * sizeof(unsigned long long), st));
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_bits, 0, bits_words * sizeof(uint32_t), st));
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
  // pass 1: Huffman layer, one lane per stream
  sfb::BatchArgs a;
  a.src_base = src_base;
  a.src_off = src_off;
  a.src_len = src_len;
  a.dst_base = dst_base;
  a.dst_off = dst_off;
  a.dst_cap = dst_cap;
  a.status = status;
  a.written = written;
  a.n = n;
  a.group_counter = ctx->d_counter;
  a.lens_scratch = ctx->d_lens;
  a.match_bits = ctx->d_bits;
  {
    const uint64_t groups = (n + 31) / 32;
    const uint64_t want = (groups + LaneCfg::WARPS - 1) / LaneCfg::WARPS;
    const uint64_t resident =
        static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->ctas_per_sm);
    const unsigned grid = static_cast<unsigned>(want < resident ? want : resident);
    sfb::huff_lanes_kernel<LaneCfg><<<grid, LaneCfg::WARPS * 32, LaneCfg::SMEM_BYTES, st>>>(a);
    SFB_TRY(ctx, cudaGetLastError());
  }
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
  // pass 2: LZ77 back-references, one warp per stream
  sfb::ResolveArgs r;
  r.dst_base = dst_base;
  r.dst_off = dst_off;
  r.written = written;
  r.match_bits = ctx->d_bits;
  r.n = n;
  r.stream_counter = ctx->d_counter + 1;
  {
    constexpr uint64_t wpc = sfb::LZ_THREADS / 32;
    const uint64_t want = (n + wpc - 1) / wpc;
    const uint64_t resident =
        static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->lz_ctas_per_sm);
    const unsigned grid = static_cast<unsigned>(want < resident ? want : resident);
    sfb::lz_resolve_kernel<<<grid, sfb::LZ_THREADS, 0, st>>>(r);
    SFB_TRY(ctx, cudaGetLastError());
  }
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
  ctx->ev_valid = true;
  ctx->launches += 2;
  return SFB200_RC_OK;
}

int sfb200_last_pass_ms(sfb200_ctx* ctx, float* out3)
{
  if (!ctx || !out3 || !ctx->ev_valid) return SFB200_RC_BAD_ARGUMENT;
  SFB_TRY(ctx, cudaSetDevice(ctx->device));
  SFB_TRY(ctx, cudaEventSynchronize(ctx->ev[3]));
  for (int i = 0; i < 3; ++i) SFB_TRY(ctx, cudaEventElapsedTime(&out3[i], ctx->ev[i], ctx->ev[i + 1]));
  return SFB200_RC_OK;
}

int sfb200_checksum_batch_device(sfb200_ctx* ctx, const uint8_t* base, const uint64_t* off,
                                 const uint64_t* len, uint64_t* out, uint64_t n,
                                 void* cuda_stream)
{
  if (!ctx) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  SFB_TRY(ctx, cudaSetDevice(ctx->device));
  const unsigned threads = 128;
  const uint64_t blocks = (n * 32 + threads - 1) / threads;
  sfb::checksum_kernel<<<static_cast<unsigned>(blocks), threads, 0,
                         static_cast<cudaStream_t>(cuda_stream)>>>(base, off, len, out, n);
  SFB_TRY(ctx, cudaGetLastError());
  ctx->launches += 1;
  return SFB200_RC_OK;
}

int sfb200_decompress_batch_host(sfb200_ctx* ctx, const uint8_t* src, uint64_t src_bytes,
                                 const uint64_t* src_off, const uint64_t* src_len,
                                 uint8_t* dst, uint64_t dst_bytes, const uint64_t* dst_off,
                                 const uint64_t* dst_cap, uint8_t* status, uint64_t* written,
                                 uint64_t n)
{
  if (!ctx) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !status) return SFB200_RC_BAD_ARGUMENT;
  if ((src_bytes && !src) || (dst_bytes && !dst)) return SFB200_RC_BAD_ARGUMENT;
  for (uint64_t i = 0; i < n; ++i) {
    if (src_off[i] > src_bytes || src_len[i] > src_bytes - src_off[i]) return SFB200_RC_BAD_ARGUMENT;
    if (dst_off[i] > dst_bytes || dst_cap[i] > dst_bytes - dst_off[i]) return SFB200_RC_BAD_ARGUMENT;
  }
  SFB_TRY(ctx, cudaSetDevice(ctx->device));
  int rc = grow(ctx, &ctx->d_src, &ctx->d_src_cap, src_bytes + 16);
  if (rc) return rc;
  rc = grow(ctx, &ctx->d_dst, &ctx->d_dst_cap, dst_bytes + 16);
  if (rc) return rc;
  if (n > ctx->d_meta_n) {
    if (ctx->d_meta) SFB_TRY(ctx, cudaFree(ctx->d_meta));
    ctx->d_meta = nullptr;
    ctx->d_meta_n = 0;
    const uint64_t want = n + n / 8 + 64;
    SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_meta), want * (5 * 8 + 1)));
    ctx->d_meta_n = want;
  }
  const uint64_t cap_n = ctx->d_meta_n;
  uint64_t* m_src_off = ctx->d_meta;
  uint64_t* m_src_len = m_src_off + cap_n;
  uint64_t* m_dst_off = m_src_len + cap_n;
  uint64_t* m_dst_cap = m_dst_off + cap_n;
  uint64_t* m_written = m_dst_cap + cap_n;
  uint8_t* m_status = reinterpret_cast<uint8_t*>(m_written + cap_n);
  cudaStream_t st = nullptr;
  if (src_bytes) SFB_TRY(ctx, cudaMemcpyAsync(ctx->d_src, src, src_bytes, cudaMemcpyHostToDevice, st));
  // dst travels both ways so that bytes the decoder never touches keep the caller's values
  if (dst_bytes) SFB_TRY(ctx, cudaMemcpyAsync(ctx->d_dst, dst, dst_bytes, cudaMemcpyHostToDevice, st));
  SFB_TRY(ctx, cudaMemcpyAsync(m_src_off, src_off, n * 8, cudaMemcpyHostToDevice, st));
  SFB_TRY(ctx, cudaMemcpyAsync(m_src_len, src_len, n * 8, cudaMemcpyHostToDevice, st));
  SFB_TRY(ctx, cudaMemcpyAsync(m_dst_off, dst_off, n * 8, cudaMemcpyHostToDevice, st));
  SFB_TRY(ctx, cudaMemcpyAsync(m_dst_cap, dst_cap, n * 8, cudaMemcpyHostToDevice, st));
  rc = sfb200_decompress_batch_device(ctx, ctx->d_src, m_src_off, m_src_len, ctx->d_dst, dst_bytes,
                                      m_dst_off, m_dst_cap, m_status, m_written, n, st);
  if (rc) return rc;
  if (dst_bytes) SFB_TRY(ctx, cudaMemcpyAsync(dst, ctx->d_dst, dst_bytes, cudaMemcpyDeviceToHost, st));
  SFB_TRY(ctx, cudaMemcpyAsync(status, m_status, n, cudaMemcpyDeviceToHost, st));
  if (written) SFB_TRY(ctx, cudaMemcpyAsync(written, m_written, n * 8, cudaMemcpyDeviceToHost, st));
  SFB_TRY(ctx, cudaStreamSynchronize(st));
  return SFB200_RC_OK;
}

int sfb200_decompress(sfb200_ctx* ctx, const uint8_t* src, size_t src_len, uint8_t* dst,
                      size_t dst_cap, uint8_t* status, uint64_t* written)
{
  if (!status) return SFB200_RC_BAD_ARGUMENT;
  const uint64_t zero = 0, sl = src_len, dc = dst_cap;
  return sfb200_decompress_batch_host(ctx, src, sl, &zero, &sl, dst, dc, &zero, &dc, status,
                                      written, 1);
}

}  // extern "C"
