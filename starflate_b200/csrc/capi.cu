// C ABI of the B200-native DEFLATE decompressor (include/starflate_b200.h).
// Host-side plumbing only: context, launch geometry, staging copies.  No CPU decode path
// exists anywhere in this library: without a CUDA device every call fails.
#include "../../include/starflate_b200.h"
#include "huff_lanes.cuh"
#include "huff_stream.cuh"
#include "lz_warp.cuh"
#include "stored_copy.cuh"
#include "lz_window.cuh"
#include "lz_jump.cuh"
#include "container.cuh"
#include "deflate_compress.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

#ifndef SFB_ROOT_LIT
#define SFB_ROOT_LIT 8
#define SFB_ROOT_DIST 6
#define SFB_POOL 96
#define SFB_WARPS 4
#endif
#ifndef SFB_LANE_CTAS
#define SFB_LANE_CTAS 2
#endif
#ifndef SFB_LZ_CTAS_PER_SM
#define SFB_LZ_CTAS_PER_SM 6  /* pass-2 CTAs per SM (0 = as many as fit); see lz_warp.cuh */
#endif
// Two LUT geometries for pass 1.  LaneCfg holds any block's codes (8/6-bit roots,
// 896 B of shared memory per lane, 8 warps per SM).  SmallCfg (6/5-bit roots, 448 B per lane,
// 16 warps per SM) is tried first: low-entropy alphabets — short codes — fit it, and twice the
// lanes in flight is twice the throughput of this latency-bound pass; a stream whose first
// block does not fit is handed to a LaneCfg launch.
using LaneCfg = sfb::Cfg<SFB_ROOT_LIT, SFB_ROOT_DIST, SFB_POOL, SFB_WARPS, SFB_LANE_CTAS>;
using SmallCfg = sfb::Cfg<6, 5, 96, 8, 2>;
// WideCfg (9/6-bit roots, a 256-entry sub-table pool, 1 728 B per lane, ONE CTA of 4 warps per SM) is
// for batches with no more streams than that has lanes (BASELINE config 4: 16 384 x 1 MiB) — such a
// batch cannot fill the SMs anyway, and its blocks are the ones whose alphabets overflow the tables
// of LaneCfg (C4: one 1-bit code for length 258 and 256 literals of 9-10 bits, 15 % of the tokens,
// which took the canonical long-code path through global scratch one lane at a time).
using WideCfg = sfb::Cfg<9, 6, 256, 4, 1>;
// single-stream mode (huff_stream.cuh): one warp per CTA, the large geometry
// (7/5-bit roots: 18.5 KiB of shared memory per warp, 8 warps per SM instead of 7 — this kernel has one
//  warp per CTA and is latency-bound too; measured 12 % faster on a 256 MiB stream and 24 % on 1 184 x
//  64 KiB streams than with the 8/6-bit roots of the lane kernel, which in turn loses with them)
#ifndef SFB_STREAM_ROOT_LIT
#define SFB_STREAM_ROOT_LIT 7
#define SFB_STREAM_ROOT_DIST 5
#endif
using StreamCfg = sfb::Cfg<SFB_STREAM_ROOT_LIT, SFB_STREAM_ROOT_DIST, SFB_POOL, 1>;
// work counters per call, four per wave k: [4k] small pass 1, [4k+1] large pass 1, [4k+2] pass 2,
// [4k+3] number of streams handed from the small to the large geometry
constexpr uint64_t kMaxWaves = 256;
constexpr uint64_t kCountersPerWave = 4;

}  // namespace

struct sfb200_ctx {
  int device = -1;
  int sm_count = 0;
  int ctas_per_sm = 0;
  int regs_per_thread = 0;
  int lz_ctas_per_sm = 0;
  int lz_regs_per_thread = 0;
  int lzw_minb = 5;            // lz_window_kernel instantiation (register budget for 4 / 5 / 6 CTAs per SM;
                               // measured with the long periodic fill in: 5 (48 registers) beats 6 (40, spills
                               // in the chunk loop) on C2 8.5 / 8.9 ms and C4 5.0 / 5.5 ms — SFB200_LZW_CTAS)
  bool lzw_pf = false;         // SFB200_LZW_PF=1: lz_window_kernel with the source prefetch (measured slower; 5 CTAs per SM only)
  uint8_t* d_queue = nullptr;  // hand-over queue of the overlapped mode (QueueArgs): counters, turns, states, items
  uint64_t d_queue_cap = 0;
  int queue_mode = 0;          // SFB200_QUEUE=1: pass 2 runs beside pass 1 (lz_window_queue_kernel)
  int queue_ctas = 4;          // ... with this many of its CTAs per SM on the second stream (SFB200_QUEUE_CTAS)
  bool compress_configured = false;
  int compress_ctas_per_sm = 0;
  bool no_stored = false;      // SFB200_NO_STORED=1: stored streams go through the lane kernel like the others (A/B runs)
  uint32_t no_pair = 0;        // SFB200_NO_PAIR=1: pass 1 takes one token per iteration (A/B runs)
  bool lz_v1 = false;          // SFB200_LZ_V1=1: the first-generation pass 2 (lz_warp.cuh), kept for A/B runs
  int small_ctas_per_sm = 0;   // SmallCfg pass 1 (0: not usable)
  int wide_ctas_per_sm = 0;    // WideCfg pass 1 (0: not usable; SFB200_NO_WIDE=1 turns it off)
  int small_regs_per_thread = 0;
  int stream_ctas_per_sm = 0;  // huff_stream_kernel (0: not usable)
  int stream_regs_per_thread = 0;
  uint32_t* d_defer = nullptr;  // streams handed from the small to the large geometry
  uint64_t d_defer_n = 0;
  uint32_t* d_order = nullptr;  // processing order (streams sorted by first block type); behind its d_order_n
                                // entries: one byte per stream, "stored_streams_kernel finished it"
  uint64_t d_order_n = 0;
  unsigned long long* d_counter = nullptr;  // 2 * kMaxWaves work counters
  uint32_t* d_lens = nullptr;  // per-resident-lane code-length scratch
  uint32_t* d_bits = nullptr;  // match-head bitmap: 1 bit per dst byte (grown on demand)
  uint64_t d_bits_words = 0;
  bool count_configured = false, check_configured = false;  // per-device function attributes set
  uint8_t* d_cont = nullptr;  // container route: payload offsets / sizes, kinds, trailer values, written
  uint64_t d_cont_cap = 0;
  uint8_t* d_find = nullptr;  // single-stream pass 1 (block_finder.cuh): counters, hash table, jobs, candidates
  uint64_t d_find_cap = 0;
  uint8_t* d_jump = nullptr;  // single-stream pass 2 (lz_jump.cuh): pointers, tile flags, round flags
  uint64_t d_jump_cap = 0;
  uint64_t* d_written = nullptr;  // pass-1 -> pass-2 sizes when the caller passes written == NULL
  uint64_t d_written_n = 0;
  uint64_t launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // call start | pass 1 start | pass 1 end | pass 2 end
  bool ev_valid = false;
  std::string err;
  // staging for the host-buffer entry points: a ring of kStageSlots sub-batch slots in device memory
  // (bounded whatever the batch: SFB200_HOST_STAGING_MB), plus the single-stream helpers' buffers
  uint8_t* d_src = nullptr;   // (single-stream helpers: size discovery, containers)
  uint64_t d_src_cap = 0;
  uint8_t* d_dst = nullptr;
  uint64_t d_dst_cap = 0;
  uint8_t* st_src[3] = {nullptr, nullptr, nullptr};
  uint8_t* st_dst[3] = {nullptr, nullptr, nullptr};
  uint64_t st_src_cap = 0, st_dst_cap = 0;      // per slot
  int st_slots = 0;                             // slots allocated so far (a small call needs one)
  uint64_t* st_meta[3] = {nullptr, nullptr, nullptr};  // per slot: src_off, src_len, dst_off, dst_cap, written (u64 x n), status (u8 x n)
  uint64_t* st_hmeta[3] = {nullptr, nullptr, nullptr};  // pinned mirror: the four input arrays (rebased) + written + status
  uint64_t st_meta_n = 0;
  cudaStream_t s_meta = nullptr;                // status / written read-back
  cudaEvent_t st_ev[3][4] = {};                 // per slot: in_done | run_done | meta_done | out_done
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};  // H2D | kernels | D2H
  cudaStream_t ps[2] = {nullptr, nullptr};           // pass-1 chain | pass-2 chain (wave overlap)
  std::vector<cudaEvent_t> wave_ev;
};

namespace {

int fail(sfb200_ctx* ctx, cudaError_t e, const char* what)
{
  if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? SFB200_RC_OUT_OF_MEMORY : SFB200_RC_CUDA_ERROR;
}

// Every entry point works on its context's device and leaves the calling thread's current
// device as it found it (a process with one context per GPU — or PyTorch beside us — must not
// see its current device change under it).
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = prev == device || cudaSetDevice(device) == cudaSuccess;
    if (prev == device) prev = -1;  // nothing to restore
  }
  ~DeviceGuard()
  {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define SFB_ENTER(ctx)                                                   \
  DeviceGuard guard_((ctx)->device);                                     \
  if (!guard_.ok) return fail((ctx), cudaGetLastError(), "cudaSetDevice"); \
  (ctx)->err.clear()

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
  DeviceGuard guard_(device);
  if (!guard_.ok) return bail(SFB200_RC_NO_DEVICE);
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
    if (const char* e = std::getenv("SFB200_LZ_V1")) ctx->lz_v1 = e[0] == '1';
    if (const char* e = std::getenv("SFB200_NO_PAIR")) ctx->no_pair = e[0] == '1';
    if (const char* e = std::getenv("SFB200_NO_STORED")) ctx->no_stored = e[0] == '1';
    if (const char* e = std::getenv("SFB200_QUEUE")) ctx->queue_mode = std::atoi(e);
    if (const char* e = std::getenv("SFB200_LZW_PF")) ctx->lzw_pf = e[0] == '1';
    if (const char* e = std::getenv("SFB200_QUEUE_CTAS")) ctx->queue_ctas = std::max(1, std::atoi(e));
    static_assert(sfb::LZ_THREADS == sfb::LZW_THREADS, "one launch geometry for both pass-2 kernels");
    static_assert(LaneCfg::WARPS == WideCfg::WARPS, "the wide launch takes the CTA count computed for the large one");
    int lz_per_sm = 0;
    if (const char* e = std::getenv("SFB200_LZW_CTAS")) {
      const int v = std::atoi(e);
      if (v == 4 || v == 5 || v == 6) ctx->lzw_minb = v;
    }
    const void* lzw = ctx->lzw_minb == 4   ? reinterpret_cast<const void*>(sfb::lz_window_kernel<4>)
                      : ctx->lzw_minb == 5 ? reinterpret_cast<const void*>(sfb::lz_window_kernel<5>)
                                           : reinterpret_cast<const void*>(sfb::lz_window_kernel<6>);
    if ((ctx->lz_v1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lz_per_sm, sfb::lz_resolve_kernel, sfb::LZ_THREADS, 0)
                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lz_per_sm, lzw, sfb::LZW_THREADS, 0)) !=
            cudaSuccess ||
        lz_per_sm < 1)
      return bail(SFB200_RC_CUDA_ERROR);
    // The queue-mode pass 2 must be able to share an SM with pass 1: an SM changes its shared-memory
    // carve-out only when it is idle, so a kernel that asks for none of it, once resident, keeps
    // pass 1 (112 KiB per CTA) off that SM for as long as it runs — ask for pass 1's carve-out.
    if (ctx->queue_mode) {
      auto qk = sfb::huff_lanes_kernel<LaneCfg, false, true>;
      if (cudaFuncSetAttribute(qk, cudaFuncAttributeMaxDynamicSharedMemorySize, LaneCfg::SMEM_BYTES) != cudaSuccess ||
          cudaFuncSetAttribute(qk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) !=
              cudaSuccess)
        return bail(SFB200_RC_CUDA_ERROR);
      const void* lzq = ctx->lzw_minb == 4   ? reinterpret_cast<const void*>(sfb::lz_window_queue_kernel<4>)
                        : ctx->lzw_minb == 5 ? reinterpret_cast<const void*>(sfb::lz_window_queue_kernel<5>)
                                             : reinterpret_cast<const void*>(sfb::lz_window_queue_kernel<6>);
      if (cudaFuncSetAttribute(lzq, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) !=
          cudaSuccess)
        return bail(SFB200_RC_CUDA_ERROR);
    }
    // fewer resident warps keep the streams' 32 KiB windows inside the L2
    int cap = SFB_LZ_CTAS_PER_SM;
    if (const char* e = std::getenv("SFB200_LZ_CTAS_PER_SM")) cap = std::atoi(e);
    ctx->lz_ctas_per_sm = (cap > 0 && cap < lz_per_sm) ? cap : lz_per_sm;
    cudaFuncAttributes lfa;
    if ((ctx->lz_v1 ? cudaFuncGetAttributes(&lfa, sfb::lz_resolve_kernel) : cudaFuncGetAttributes(&lfa, lzw)) == cudaSuccess)
      ctx->lz_regs_per_thread = lfa.numRegs;
  }
  if (cudaMalloc(reinterpret_cast<void**>(&ctx->d_counter), (kCountersPerWave * kMaxWaves + 16) * sizeof(unsigned long long)) !=
      cudaSuccess)
    return bail(SFB200_RC_OUT_OF_MEMORY);
  {
    auto sk = sfb::huff_lanes_kernel<SmallCfg>;
    int sp = 0;
    if (cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallCfg::SMEM_BYTES) ==
            cudaSuccess &&
        cudaFuncSetAttribute(sk, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sp, sk, SmallCfg::WARPS * 32,
                                                      SmallCfg::SMEM_BYTES) == cudaSuccess)
      ctx->small_ctas_per_sm = sp;
    cudaFuncAttributes sfa;
    if (cudaFuncGetAttributes(&sfa, sk) == cudaSuccess) ctx->small_regs_per_thread = sfa.numRegs;
    cudaGetLastError();
  }
  {
    auto sk = sfb::huff_lanes_kernel<WideCfg>;
    int sp = 0;
    bool off = false;
    if (const char* e = std::getenv("SFB200_NO_WIDE")) off = e[0] == '1';
    if (!off &&
        cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, WideCfg::SMEM_BYTES) ==
            cudaSuccess &&
        cudaFuncSetAttribute(sk, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sp, sk, WideCfg::WARPS * 32,
                                                      WideCfg::SMEM_BYTES) == cudaSuccess)
      ctx->wide_ctas_per_sm = sp;
    cudaGetLastError();
  }
  {
    auto sk = sfb::huff_stream_kernel<StreamCfg>;
    int sp = 0;
    if (cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg::SMEM_BYTES) ==
            cudaSuccess &&
        cudaFuncSetAttribute(sk, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sp, sk, 32, StreamCfg::SMEM_BYTES) == cudaSuccess)
      ctx->stream_ctas_per_sm = sp;
    cudaFuncAttributes sfa;
    if (cudaFuncGetAttributes(&sfa, sk) == cudaSuccess) ctx->stream_regs_per_thread = sfa.numRegs;
    cudaGetLastError();
  }
  const size_t big_warps = std::max(static_cast<size_t>(per_sm) * LaneCfg::WARPS,
                                    static_cast<size_t>(ctx->stream_ctas_per_sm));
  const size_t small_warps = static_cast<size_t>(ctx->small_ctas_per_sm) * SmallCfg::WARPS;
  const size_t lens_bytes = static_cast<size_t>(ctx->sm_count) * std::max(big_warps, small_warps) *
                            sfb::SCRATCH_WORDS * 32 * sizeof(uint32_t);
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
  DeviceGuard guard_(ctx->device);
  cudaFree(ctx->d_counter);
  cudaFree(ctx->d_lens);
  cudaFree(ctx->d_bits);
  cudaFree(ctx->d_written);
  cudaFree(ctx->d_jump);
  cudaFree(ctx->d_find);
  cudaFree(ctx->d_cont);
  cudaFree(ctx->d_defer);
  cudaFree(ctx->d_order);
  cudaFree(ctx->d_queue);
  cudaFree(ctx->d_src);
  cudaFree(ctx->d_dst);
  for (int k = 0; k < 3; ++k) {
    cudaFree(ctx->st_src[k]);
    cudaFree(ctx->st_dst[k]);
    cudaFree(ctx->st_meta[k]);
    if (ctx->st_hmeta[k]) cudaFreeHost(ctx->st_hmeta[k]);
    for (auto& e : ctx->st_ev[k])
      if (e) cudaEventDestroy(e);
  }
  if (ctx->s_meta) cudaStreamDestroy(ctx->s_meta);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  for (auto& st : ctx->hs)
    if (st) cudaStreamDestroy(st);
  for (auto& st : ctx->ps)
    if (st) cudaStreamDestroy(st);
  for (auto& e : ctx->wave_ev) cudaEventDestroy(e);
  delete ctx;
}

const char* sfb200_last_error(const sfb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// Diagnostic of the experimental queue mode (SFB200_QUEUE=1; not part of the declared ABI): waits for
// the device and copies the queue's eight counters — tail, head, final_tail, done, dbg[0..3].
int sfb200_debug_queue(sfb200_ctx* ctx, unsigned long long* out8)
{
  if (!ctx || !out8 || !ctx->d_queue) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  SFB_TRY(ctx, cudaDeviceSynchronize());
  SFB_TRY(ctx, cudaMemcpy(out8, ctx->d_queue, 64, cudaMemcpyDeviceToHost));
  return SFB200_RC_OK;
}

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
  out->small_warps_per_cta = SmallCfg::WARPS;
  out->small_ctas_per_sm = ctx->small_ctas_per_sm;
  out->small_regs_per_thread = ctx->small_regs_per_thread;
  out->stream_ctas_per_sm = ctx->stream_ctas_per_sm;
  out->stream_regs_per_thread = ctx->stream_regs_per_thread;
  out->kernel_launches = ctx->launches;
  return SFB200_RC_OK;
}

}  // extern "C"

namespace {
// chunked input (sfb200_inflate_stream_feed): where each stream starts, and where its decode got to
struct Resume {
  const uint64_t* start_bit;
  const uint64_t* start_out;
  uint64_t* blk_end;
};

int batch_device_impl(sfb200_ctx* ctx, const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len,
                      uint8_t* dst_base, uint64_t dst_bytes, const uint64_t* dst_off, const uint64_t* dst_cap,
                      uint8_t* status, uint64_t* written, uint64_t n, void* cuda_stream, const Resume* resume);
}  // namespace

extern "C" {

int sfb200_decompress_batch_device(sfb200_ctx* ctx, const uint8_t* src_base,
                                   const uint64_t* src_off, const uint64_t* src_len,
                                   uint8_t* dst_base, uint64_t dst_bytes, const uint64_t* dst_off,
                                   const uint64_t* dst_cap, uint8_t* status,
                                   uint64_t* written, uint64_t n, void* cuda_stream)
{
  return batch_device_impl(ctx, src_base, src_off, src_len, dst_base, dst_bytes, dst_off, dst_cap, status, written,
                           n, cuda_stream, nullptr);
}

}  // extern "C"

namespace {
int batch_device_impl(sfb200_ctx* ctx, const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len,
                      uint8_t* dst_base, uint64_t dst_bytes, const uint64_t* dst_off, const uint64_t* dst_cap,
                      uint8_t* status, uint64_t* written, uint64_t n, void* cuda_stream, const Resume* resume)
{
  if (!ctx) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !status) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // the kernels index dst (and the bitmap) from a 128-byte aligned base
  const uint64_t delta = reinterpret_cast<uintptr_t>(dst_base) & 127u;
  dst_base -= delta;
  // scratch: one bit per dst byte (+ guard words), grown on demand and kept
  const uint64_t bits_words = (dst_bytes + delta) / 32 + 8;
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
  // Few streams: a lane per stream would leave the GPU empty, so each stream gets a warp that
  // decodes 32 spans of it speculatively (huff_stream.cuh).  Same results, by construction.
  // Measured on B200 with 64 KiB streams: 1 184 streams 1.8 ms against 4.7 ms lane per stream, 3 552
  // 4.5 against 4.8 ms, 4 736 5.8 against 5.0 ms: up to three streams per resident warp of the lane
  // kernel.  Small streams do not pay for a warp each (4 KiB pages are 7x slower that way): stream
  // sizes are only known on the device, so the room the caller gave them stands in (at least
  // 32 KiB of dst per stream on average).
  bool stream_mode = ctx->stream_ctas_per_sm > 0 &&
                     n * 32 <= 3 * static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->ctas_per_sm) *
                                       LaneCfg::WARPS * 32 &&
                     dst_bytes / n >= 32768;
  if (const char* e = std::getenv("SFB200_STREAM_MODE")) stream_mode = e[0] == '1' && ctx->stream_ctas_per_sm > 0;
  if (resume) stream_mode = false;  // (a decode that continues at a block boundary: the lane kernel only)
  // One stream: pass 2 works on the whole output at once (lz_jump.cuh) instead of one warp walking it
  // (also a FEW large streams, one after the other: sizes are only known on the device, so "large" is
  //  judged by the room the caller gave them)
  bool jump = stream_mode && (n == 1 || (n <= 16 && dst_bytes / n >= (1ull << 20)));
  if (const char* e = std::getenv("SFB200_JUMP")) jump = jump && e[0] != '0';
  uint64_t jump_tiles = 0, stripe_tiles = 1ull << 23;  // (one stripe unless SFB200_JUMP_STRIPE_MB says otherwise)
  if (jump) {
    jump_tiles = (dst_bytes + delta + 128 + sfb::JUMP_TILE - 1) / sfb::JUMP_TILE + 1;
    if (const char* e = std::getenv("SFB200_JUMP_STRIPE_MB")) {
      const long mb = std::atol(e);
      if (mb > 0) stripe_tiles = static_cast<uint64_t>(mb) * ((1u << 20) / sfb::JUMP_TILE);
    }
    const uint64_t stripes = (jump_tiles + stripe_tiles - 1) / stripe_tiles;
    const uint64_t need = jump_tiles * sfb::JUMP_TILE * 4 + jump_tiles * 4 * (1 + sfb::JUMP_TILE / 128) +
                          stripes * (sfb::JUMP_MAX_ROUNDS + 1) * 4;
    const int rc = grow(ctx, &ctx->d_jump, &ctx->d_jump_cap, need);
    if (rc == SFB200_RC_OUT_OF_MEMORY) {
      // no room for the pointer array (4 bytes per dst byte): a warp walks the stream instead —
      // slower, same results
      cudaGetLastError();
      ctx->err.clear();
      jump = false;
    } else if (rc != SFB200_RC_OK) {
      return rc;
    }
  }
  // ... and pass 1 decodes its blocks side by side (block_finder.cuh).  The input's size is only known
  // on the device; the scratch is sized for an input a little larger than the output, and block
  // starts that do not fit are simply not found (their blocks are then decoded by the tail job).
  bool blocks = jump;
  if (const char* e = std::getenv("SFB200_BLOCKS")) blocks = blocks && e[0] != '0';
  uint32_t cand_cap = 0, job_cap = 0, tab_size = 0, rec_cap = 0;
  if (blocks) {
    const uint64_t src_bits_max = 8 * (dst_bytes + dst_bytes / 32 + 65536);
    cand_cap = static_cast<uint32_t>(std::min<uint64_t>(src_bits_max / 128 + 4096, 1ull << 30));
    job_cap = static_cast<uint32_t>(std::min<uint64_t>(src_bits_max / 2048 + 1024, 1ull << 27));
    tab_size = 1024;
    while (tab_size < 2 * job_cap) tab_size *= 2;
    rec_cap = static_cast<uint32_t>(std::min<uint64_t>(src_bits_max / (16 * sfb::SPAN_BITS) + 65536, 1ull << 26));
    const uint64_t need = 64 + 4ull * tab_size + sizeof(sfb::BlockJob) * static_cast<uint64_t>(job_cap) +
                          8ull * cand_cap + sizeof(sfb::WinRec) * static_cast<uint64_t>(rec_cap) +
                          28ull * job_cap + 64;
    const int rc = grow(ctx, &ctx->d_find, &ctx->d_find_cap, need);
    if (rc == SFB200_RC_OUT_OF_MEMORY) {  // blocks one after the other instead of side by side
      cudaGetLastError();
      ctx->err.clear();
      blocks = false;
      jump = false;
    } else if (rc != SFB200_RC_OK) {
      return rc;
    }
  }
  // Geometry.  Measured on C2 (profiles/r01_small_geometry_c2.md): with 16 warps per SM the small
  // geometry makes pass 1 issue-bound (63 % of issue slots, 11.1 ms against 12.1 ms for the
  // large one), and the few percent of streams it hands on cost a whole extra stream latency
  // (4.5 ms) — so the large geometry is the default and the small one is opt-in
  // (SFB200_GEOMETRY=small) until the hand-over can overlap.
  bool use_small = false;
  if (const char* e = std::getenv("SFB200_GEOMETRY")) use_small = e[0] == 's' && ctx->small_ctas_per_sm > 0;
  if (use_small && n > ctx->d_defer_n) {
    if (ctx->d_defer) SFB_TRY(ctx, cudaFree(ctx->d_defer));
    ctx->d_defer = nullptr;
    ctx->d_defer_n = 0;
    const uint64_t want = n + n / 8 + 64;
    SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_defer), want * sizeof(uint32_t)));
    ctx->d_defer_n = want;
  }
  // Waves.  The two passes run BACK TO BACK over the whole batch: one launch each, inside which
  // warps pull groups of streams dynamically.  Cutting the batch into waves and running pass 2 of
  // wave k beside pass 1 of wave k+1 on two internal streams is kept for experiments
  // (SFB200_OVERLAP=1; wave size SFB200_WAVE_STREAMS) — measured at the end of round 2 it loses on
  // every shape: 65 536 x 64 KiB 19.4 against 18.6 ms, 131 072 x 64 KiB 39.7 against 37.0 ms,
  // 1 048 576 x 4 KiB 18.5 against 18.1 ms: pass 1 is latency-bound and slows down
  // by more than pass 2 gains when it shares its SMs.
  const uint64_t big_lanes = static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->ctas_per_sm) *
                             LaneCfg::WARPS * 32;
  const uint64_t small_lanes = static_cast<uint64_t>(ctx->sm_count) *
                               static_cast<uint64_t>(ctx->small_ctas_per_sm) * SmallCfg::WARPS * 32;
  // (a launch takes one wave of streams, or several when the batch has more than eight waves:
  //  inside a launch warps pull groups dynamically, which evens out unequal streams and lets a
  //  lane keep its fixed-Huffman tables from one stream to the next)
  const uint64_t lanes = use_small ? small_lanes : big_lanes;
  uint64_t wave = lanes * std::max<uint64_t>(1, (n + 8 * lanes - 1) / (8 * lanes));
  if (const char* e = std::getenv("SFB200_WAVE_STREAMS")) {  // (experiments: the size of a wave)
    const uint64_t v = std::strtoull(e, nullptr, 10);
    if (v >= 1024) wave = v;
  }
  uint64_t n_waves = (n + wave - 1) / wave;
  bool overlap_waves = false;
  if (const char* e = std::getenv("SFB200_OVERLAP")) overlap_waves = e[0] == '1';
  if (!overlap_waves) n_waves = 1;
  if (n_waves > kMaxWaves) n_waves = 1;  // (counters are a fixed array: enormous batches run unsplit)
  if (stream_mode) n_waves = 1;
  // Queue mode: ONE launch of pass 1, pass 2 beside it on the second stream, fed segment by segment
  // (deflate_lane.cuh: QueueArgs; lz_window.cuh: lz_window_queue_kernel).
  const bool qmode = ctx->queue_mode != 0 && !stream_mode && !use_small && !resume && !jump && !ctx->lz_v1;
  uint64_t qcap = 0;
  if (qmode) {
    n_waves = 1;
    qcap = 2 * n + ((dst_bytes + delta) >> sfb::Q_SEG_SHIFT) + 64;
    const int rc = grow(ctx, &ctx->d_queue, &ctx->d_queue_cap, 64 + 24 * n + 8 * qcap + 64);
    if (rc != SFB200_RC_OK) return rc;
  }
  const uint64_t per_wave = n_waves == 1 ? n : wave;
  const bool overlap = n_waves > 1 || qmode;
  if (overlap) {
    if (!ctx->ps[0]) {  // pass 1's stream: the higher priority (its CTAs are placed first)
      int lo = 0, hi = 0;
      SFB_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      SFB_TRY(ctx, cudaStreamCreateWithPriority(&ctx->ps[0], cudaStreamNonBlocking, hi));
    }
    for (auto& ps : ctx->ps)
      if (!ps) SFB_TRY(ctx, cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    for (uint64_t k = ctx->wave_ev.size(); k < n_waves + 3; ++k) {
      cudaEvent_t e = nullptr;
      SFB_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->wave_ev.push_back(e);
    }
  }
  cudaStream_t s1 = overlap ? ctx->ps[0] : st;  // pass-1 chain
  cudaStream_t s2 = overlap ? ctx->ps[1] : st;  // pass-2 chain
  ctx->ev_valid = false;
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_counter, 0,
                               kCountersPerWave * std::max<uint64_t>(n_waves, 1) * sizeof(unsigned long long), st));
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_counter + kCountersPerWave * kMaxWaves, 0, 16 * sizeof(unsigned long long), st));
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_bits, 0, bits_words * sizeof(uint32_t), st));
  if (qmode) SFB_TRY(ctx, cudaMemsetAsync(ctx->d_queue, 0, 64 + 24 * n + 8 * qcap, st));
  // Order the streams by the type of their first block (see lz_warp.cuh: PrepArgs)
  uint8_t* handled = nullptr;
  bool sorted = n >= 64 && n < 0xffffffffull && !stream_mode && !resume;
  if (const char* e = std::getenv("SFB200_NO_SORT"))
    if (e[0] == '1') sorted = false;
  if (sorted) {
    if (n > ctx->d_order_n) {
      if (ctx->d_order) SFB_TRY(ctx, cudaFree(ctx->d_order));
      ctx->d_order = nullptr;
      ctx->d_order_n = 0;
      const uint64_t want = n + n / 8 + 64;
      SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_order), want * (sizeof(uint32_t) + 1)));
      ctx->d_order_n = want;
    }
    sfb::PrepArgs pa;
    pa.src_base = src_base;
    pa.src_off = src_off;
    pa.src_len = src_len;
    pa.n = n;
    pa.type_count = ctx->d_counter + kCountersPerWave * kMaxWaves;
    pa.type_cursor = pa.type_count + 4;
    pa.order = ctx->d_order;
    const unsigned grid = static_cast<unsigned>((n + sfb::PREP_THREADS - 1) / sfb::PREP_THREADS);
    sfb::prep_count_kernel<<<grid, sfb::PREP_THREADS, 0, st>>>(pa);
    sfb::prep_scatter_kernel<<<grid, sfb::PREP_THREADS, 0, st>>>(pa);
    SFB_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    // Streams of stored blocks: a warp each, wide copies (stored_copy.cuh); the lane kernel skips
    // what this kernel finished.
    if (!ctx->no_stored) {
      handled = reinterpret_cast<uint8_t*>(ctx->d_order + ctx->d_order_n);
      SFB_TRY(ctx, cudaMemsetAsync(handled, 0, n, st));
      sfb::StoredArgs sa;
      sa.src_base = src_base;
      sa.src_off = src_off;
      sa.src_len = src_len;
      sa.dst_base = dst_base + delta;
      sa.dst_off = dst_off;
      sa.dst_cap = dst_cap;
      sa.status = status;
      sa.written = written;
      sa.handled = handled;
      sa.order = ctx->d_order;
      sa.type_count = pa.type_count;
      sa.counter = pa.type_count + 8;
      const uint64_t warps = static_cast<uint64_t>(ctx->sm_count) * 8 * (sfb::STORED_THREADS / 32);
      const uint64_t want = (std::min<uint64_t>(n, warps) + sfb::STORED_THREADS / 32 - 1) / (sfb::STORED_THREADS / 32);
      sfb::stored_streams_kernel<<<static_cast<unsigned>(want), sfb::STORED_THREADS, 0, st>>>(sa);
      SFB_TRY(ctx, cudaGetLastError());
      ctx->launches += 1;
    }
  }
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
  if (overlap) {
    SFB_TRY(ctx, cudaEventRecord(ctx->wave_ev[n_waves], st));
    SFB_TRY(ctx, cudaStreamWaitEvent(s1, ctx->wave_ev[n_waves], 0));
    SFB_TRY(ctx, cudaStreamWaitEvent(s2, ctx->wave_ev[n_waves], 0));
  }
  for (uint64_t k = 0; k < n_waves; ++k) {
    const uint64_t first = k * per_wave;
    const uint64_t cnt = std::min(per_wave, n - first);
    unsigned long long* const ctr = ctx->d_counter + kCountersPerWave * k;
    const uint32_t* const order = sorted ? ctx->d_order + first : nullptr;
    // pass 1: Huffman layer, one lane per stream
    sfb::BatchArgs a{};
    a.src_base = src_base;
    a.src_off = src_off;
    a.src_len = src_len;
    a.dst_base = dst_base;
    a.dst_delta = delta;
    a.dst_off = dst_off;
    a.dst_cap = dst_cap;
    a.status = status;
    a.written = written;
    a.n = cnt;
    a.idx_base = first;
    a.lens_scratch = ctx->d_lens;
    a.match_bits = ctx->d_bits;
    a.defer_list = nullptr;
    a.defer_count = nullptr;
    a.todo_list = order;
    a.todo_count = nullptr;
    a.no_pair = ctx->no_pair;
    a.handled = handled;
    a.start_bit = resume ? resume->start_bit : nullptr;
    a.start_out = resume ? resume->start_out : nullptr;
    a.blk_end = resume ? resume->blk_end : nullptr;
    if (qmode) {
      unsigned long long* const c = reinterpret_cast<unsigned long long*>(ctx->d_queue);
      a.q.tail = c + 0;
      a.q.head = c + 1;
      a.q.final_tail = c + 2;
      a.q.done = reinterpret_cast<unsigned int*>(c + 3);
      a.q.dbg = c + 4;
      a.q.turn = reinterpret_cast<unsigned int*>(ctx->d_queue + 64);
      a.q.state = reinterpret_cast<uint32_t*>(ctx->d_queue + 64 + 4 * n);
      a.q.items = reinterpret_cast<unsigned long long*>(ctx->d_queue + 64 + 24 * n);
      a.q.cap = qcap;
      a.q.seg_shift = sfb::Q_SEG_SHIFT;
      a.q.lag = sfb::Q_LAG;
    }
    const uint64_t groups = (cnt + 31) / 32;
    if (stream_mode) {
      sfb::StreamArgs sa;
      sa.src_base = src_base;
      sa.src_off = src_off;
      sa.src_len = src_len;
      sa.dst_base = dst_base;
      sa.dst_delta = delta;
      sa.dst_off = dst_off;
      sa.dst_cap = dst_cap;
      sa.status = status;
      sa.written = written;
      sa.list = nullptr;
      sa.idx_base = first;
      sa.n = cnt;
      sa.stream_counter = ctr + 0;
      sa.lens_scratch = ctx->d_lens;
      sa.match_bits = ctx->d_bits;
      sa.mode = 0;
      sa.jobs = nullptr;
      sa.job_count = nullptr;
      sa.job_cap = 0;
      sa.job_tab = nullptr;
      sa.tab_mask = 0;
      sa.tail_job = nullptr;
      sa.recs = nullptr;
      sa.rec_count = nullptr;
      sa.rec_cap = 0;
      const uint64_t resident =
          static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->stream_ctas_per_sm);
      if (blocks) {
        // find block starts | count each candidate block | chain | write the chained blocks
        uint32_t* const counters = reinterpret_cast<uint32_t*>(ctx->d_find);  // cand_count, job_count, tail_job
        sfb::FindArgs f;
        f.src_base = src_base;
        f.src_off = src_off;
        f.src_len = src_len;
        f.idx = 0;
        f.cand_count = counters + 0;
        f.job_count = counters + 1;
        f.tail_job = counters + 2;
        f.job_tab = counters + 16;
        f.tab_mask = tab_size - 1;
        f.jobs = reinterpret_cast<sfb::BlockJob*>(ctx->d_find + 64 + 4ull * tab_size);
        f.job_cap = job_cap;
        f.cand = reinterpret_cast<uint64_t*>(ctx->d_find + 64 + 4ull * tab_size +
                                             sizeof(sfb::BlockJob) * static_cast<uint64_t>(job_cap));
        f.cand_cap = cand_cap;
        f.dst_cap = dst_cap;
        sa.recs = reinterpret_cast<sfb::WinRec*>(reinterpret_cast<uint8_t*>(f.cand) + 8ull * cand_cap);
        sa.rec_count = counters + 3;
        sa.rec_cap = rec_cap;
        sfb::ChainScratch cs;  // behind the window records (8-byte aligned: WinRec is 272 bytes)
        cs.wa = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sa.recs) + sizeof(sfb::WinRec) * static_cast<uint64_t>(rec_cap));
        cs.wb = cs.wa + job_cap;
        cs.ja = reinterpret_cast<uint32_t*>(cs.wb + job_cap);
        cs.jb = cs.ja + job_cap;
        cs.mark = cs.jb + job_cap;
        bool seq_chain = false;
        if (const char* e = std::getenv("SFB200_CHAIN")) seq_chain = e[0] == 's';
        sa.jobs = f.jobs;
        sa.job_count = f.job_count;
        sa.job_cap = job_cap;
        sa.job_tab = f.job_tab;
        sa.tab_mask = f.tab_mask;
        sa.tail_job = f.tail_job;
        const unsigned grid = static_cast<unsigned>(resident);
        for (uint64_t si = 0; si < cnt; ++si) {
          f.idx = first + si;
          sa.idx_base = first + si;
          SFB_TRY(ctx, cudaMemsetAsync(ctx->d_find, 0, 64 + 4ull * tab_size, s1));
          if (si) SFB_TRY(ctx, cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), s1));
          sfb::find_candidates_kernel<<<ctx->sm_count * 8, sfb::FIND_THREADS, 0, s1>>>(f);
          sfb::verify_candidates_kernel<<<ctx->sm_count * 4, sfb::FIND_THREADS, 0, s1>>>(f);
          sa.mode = 1;
          sa.stream_counter = ctr + 0;
          sfb::huff_stream_kernel<StreamCfg><<<grid, 32, StreamCfg::SMEM_BYTES, s1>>>(sa);
          if (seq_chain) sfb::chain_kernel<<<1, 32, 0, s1>>>(f);
          else sfb::chain_parallel_kernel<<<1, sfb::CHAIN_THREADS, 0, s1>>>(f, cs);
          sa.mode = 2;
          sa.stream_counter = ctr + 1;
          sfb::huff_stream_kernel<StreamCfg><<<grid, 32, StreamCfg::SMEM_BYTES, s1>>>(sa);
          SFB_TRY(ctx, cudaGetLastError());
        }
        ctx->launches += 5 * cnt;
      } else {
        const unsigned grid = static_cast<unsigned>(cnt < resident ? cnt : resident);
        sfb::huff_stream_kernel<StreamCfg><<<grid, 32, StreamCfg::SMEM_BYTES, s1>>>(sa);
        SFB_TRY(ctx, cudaGetLastError());
        ctx->launches += 1;
      }
    } else if (use_small) {
      a.group_counter = ctr + 0;
      a.defer_list = ctx->d_defer + first;
      a.defer_count = ctr + 3;
      const uint64_t want = (groups + SmallCfg::WARPS - 1) / SmallCfg::WARPS;
      const uint64_t resident =
          static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->small_ctas_per_sm);
      const unsigned grid = static_cast<unsigned>(want < resident ? want : resident);
      sfb::huff_lanes_kernel<SmallCfg><<<grid, SmallCfg::WARPS * 32, SmallCfg::SMEM_BYTES, s1>>>(a);
      SFB_TRY(ctx, cudaGetLastError());
      ctx->launches += 1;
      // the streams it handed on (none for low-entropy data: this launch then ends at once)
      a.defer_list = nullptr;
      a.defer_count = nullptr;
      a.todo_list = ctx->d_defer + first;
      a.todo_count = ctr + 3;
    }
    if (!stream_mode) {
      a.group_counter = ctr + 1;
      const uint64_t want = (groups + LaneCfg::WARPS - 1) / LaneCfg::WARPS;
      const uint64_t resident =
          static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->ctas_per_sm);
      const unsigned grid = static_cast<unsigned>(want < resident ? want : resident);
      // (a batch that fits the lanes of the wide geometry takes it: see WideCfg)
      const uint64_t wide_resident =
          static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->wide_ctas_per_sm);
      if (qmode)
        sfb::huff_lanes_kernel<LaneCfg, false, true><<<grid, LaneCfg::WARPS * 32, LaneCfg::SMEM_BYTES, s1>>>(a);
      else if (!use_small && n_waves == 1 && wide_resident != 0 && want <= wide_resident)
        sfb::huff_lanes_kernel<WideCfg><<<static_cast<unsigned>(want), WideCfg::WARPS * 32, WideCfg::SMEM_BYTES, s1>>>(a);
      else
        sfb::huff_lanes_kernel<LaneCfg><<<grid, LaneCfg::WARPS * 32, LaneCfg::SMEM_BYTES, s1>>>(a);
      SFB_TRY(ctx, cudaGetLastError());
      ctx->launches += 1;
    }
    if (overlap && !qmode) {
      SFB_TRY(ctx, cudaEventRecord(ctx->wave_ev[k], s1));
      SFB_TRY(ctx, cudaStreamWaitEvent(s2, ctx->wave_ev[k], 0));
    }
    if (k + 1 == n_waves) SFB_TRY(ctx, cudaEventRecord(ctx->ev[2], s1));  // end of all of pass 1
    // pass 2: LZ77 back-references, one warp per stream
    sfb::ResolveArgs r{};
    r.dst_base = dst_base;
    r.dst_delta = delta;
    r.dst_off = dst_off;
    r.written = written;
    r.match_bits = ctx->d_bits;
    r.n = cnt;
    r.idx_base = first;
    r.todo_list = order;
    r.stream_counter = ctr + 2;
    if (qmode) {
      // beside pass 1 (second stream), and the rest of the GPU's worth behind it (its own stream)
      // (beside pass 1: CTAs of 4 warps — 6 144 registers, what is left next to two CTAs of pass 1 —
      //  and few enough of them that pass 1 finds room on some SM whatever the order the two kernels
      //  start in: its stream has the higher priority, but nothing in CUDA promises the order)
      const unsigned per_sm = static_cast<unsigned>(ctx->lz_ctas_per_sm);
      const unsigned qa = std::min<unsigned>(static_cast<unsigned>(ctx->queue_ctas), 4u);
      const unsigned grid_a = static_cast<unsigned>(ctx->sm_count) * qa;
      const unsigned grid_b = static_cast<unsigned>(ctx->sm_count) * std::max(1u, per_sm - qa / 2u);
      if (ctx->lzw_minb == 4) {
        sfb::lz_window_queue_kernel<4><<<grid_a, 128, 0, s2>>>(r, a.q);
        sfb::lz_window_queue_kernel<4><<<grid_b, sfb::LZW_THREADS, 0, s1>>>(r, a.q);
      } else if (ctx->lzw_minb == 5) {
        sfb::lz_window_queue_kernel<5><<<grid_a, 128, 0, s2>>>(r, a.q);
        sfb::lz_window_queue_kernel<5><<<grid_b, sfb::LZW_THREADS, 0, s1>>>(r, a.q);
      } else {
        sfb::lz_window_queue_kernel<6><<<grid_a, 128, 0, s2>>>(r, a.q);
        sfb::lz_window_queue_kernel<6><<<grid_b, sfb::LZW_THREADS, 0, s1>>>(r, a.q);
      }
      SFB_TRY(ctx, cudaGetLastError());
      ctx->launches += 1;
    } else if (jump) {
      sfb::JumpArgs j;
      j.dst_base = dst_base;
      j.dst_delta = delta;
      j.dst_off = dst_off;
      j.written = written;
      j.match_bits = ctx->d_bits;
      j.ptr = reinterpret_cast<uint32_t*>(ctx->d_jump);
      j.tile_done = j.ptr + jump_tiles * sfb::JUMP_TILE;
      j.todo = j.tile_done + jump_tiles;
      const uint64_t stripes = (jump_tiles + stripe_tiles - 1) / stripe_tiles;
      uint32_t* const todo0 = j.todo;
      j.open = j.todo + stripes * (sfb::JUMP_MAX_ROUNDS + 1);
      constexpr uint64_t wpc = sfb::JUMP_THREADS / 32;
      const uint64_t resident = static_cast<uint64_t>(ctx->sm_count) * 8;
      for (uint64_t si = 0; si < cnt; ++si) {
        j.idx = first + si;
        SFB_TRY(ctx, cudaMemsetAsync(j.tile_done, 0, (jump_tiles + stripes * (sfb::JUMP_MAX_ROUNDS + 1)) * 4, s2));
        // pointers of stripe sp + 1 are set up BEFORE the rounds of stripe sp: the match that reaches
        // into a stripe from the left is read from its descriptor, which the rounds of the stripe
        // before replace by the final bytes
        auto set_stripe = [&](uint64_t sp) {
          j.tile_lo = static_cast<uint32_t>(sp * stripe_tiles);
          j.tile_hi = static_cast<uint32_t>(std::min<uint64_t>(jump_tiles, (sp + 1) * stripe_tiles));
          j.todo = todo0 + sp * (sfb::JUMP_MAX_ROUNDS + 1);
          const uint64_t want = (j.tile_hi - j.tile_lo + wpc - 1) / wpc;
          return static_cast<unsigned>(want < resident ? want : resident);
        };
        j.round = 0;
        sfb::lz_jump_init_kernel<<<set_stripe(0), sfb::JUMP_THREADS, 0, s2>>>(j);
        ctx->launches += 1;
        for (uint64_t sp = 0; sp < stripes; ++sp) {
          if (sp + 1 < stripes) {
            j.round = 0;
            sfb::lz_jump_init_kernel<<<set_stripe(sp + 1), sfb::JUMP_THREADS, 0, s2>>>(j);
            ctx->launches += 1;
          }
          const unsigned grid = set_stripe(sp);
          // a chain inside a stripe has fewer hops than the stripe has bytes, and a round divides
          // them by three: no more launches than that can need (+ 1 to see that nothing moved)
          int max_rounds = 2;
          for (uint64_t reach = 3; reach < static_cast<uint64_t>(j.tile_hi - j.tile_lo) * sfb::JUMP_TILE; reach *= 3)
            ++max_rounds;
          if (max_rounds > sfb::JUMP_MAX_ROUNDS) max_rounds = sfb::JUMP_MAX_ROUNDS;
          for (int r = 1; r <= max_rounds; ++r) {
            j.round = static_cast<uint32_t>(r);
            sfb::lz_jump_round_kernel<<<grid, sfb::JUMP_THREADS, 0, s2>>>(j);
            ctx->launches += 1;
          }
          SFB_TRY(ctx, cudaGetLastError());
        }
      }
    } else {
      constexpr uint64_t wpc = sfb::LZ_THREADS / 32;
      const uint64_t resident =
          static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->lz_ctas_per_sm);
      const uint64_t want = (cnt + wpc - 1) / wpc;
      const unsigned grid = static_cast<unsigned>(want < resident ? want : resident);
      if (ctx->lz_v1) sfb::lz_resolve_kernel<<<grid, sfb::LZ_THREADS, 0, s2>>>(r);
      else if (ctx->lzw_minb == 4) sfb::lz_window_kernel<4><<<grid, sfb::LZW_THREADS, 0, s2>>>(r);
      else if (ctx->lzw_minb == 5 && ctx->lzw_pf) sfb::lz_window_kernel<5, true><<<grid, sfb::LZW_THREADS, 0, s2>>>(r);
      else if (ctx->lzw_minb == 5) sfb::lz_window_kernel<5><<<grid, sfb::LZW_THREADS, 0, s2>>>(r);
      else sfb::lz_window_kernel<6><<<grid, sfb::LZW_THREADS, 0, s2>>>(r);
      SFB_TRY(ctx, cudaGetLastError());
    }
    if (qmode || !jump) ctx->launches += 1;   // (the jump route counted its launches one by one)
  }
  if (overlap) {
    SFB_TRY(ctx, cudaEventRecord(ctx->wave_ev[n_waves + 1], s1));
    SFB_TRY(ctx, cudaEventRecord(ctx->wave_ev[n_waves + 2], s2));
    SFB_TRY(ctx, cudaStreamWaitEvent(st, ctx->wave_ev[n_waves + 1], 0));
    SFB_TRY(ctx, cudaStreamWaitEvent(st, ctx->wave_ev[n_waves + 2], 0));
  }
  SFB_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
  ctx->ev_valid = true;
  if (qmode) {
    // (experimental mode: synchronous, so that a consumer that gave up waiting — lz_window.cuh:
    //  lzq_give_up — fails the call instead of leaving matches unresolved behind a success)
    unsigned long long q8[8] = {0};
    SFB_TRY(ctx, cudaStreamSynchronize(st));
    SFB_TRY(ctx, cudaMemcpy(q8, ctx->d_queue, sizeof(q8), cudaMemcpyDeviceToHost));
    if (q8[4] != 0ull) {
      ctx->err = "queue mode: a pass-2 consumer gave up waiting (" + std::string(q8[4] == 1ull ? "for an item" : "for its turn") +
                 ", " + std::to_string(q8[5]) + " / " + std::to_string(q8[6]) + " / " + std::to_string(q8[7]) + ")";
      return SFB200_RC_CUDA_ERROR;
    }
  }
  return SFB200_RC_OK;
}

}  // namespace

extern "C" {

int sfb200_decompressed_size_batch_device(sfb200_ctx* ctx, const uint8_t* src_base,
                                          const uint64_t* src_off, const uint64_t* src_len,
                                          uint8_t* status, uint64_t* size, uint64_t n,
                                          void* cuda_stream)
{
  if (!ctx) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !status || !size) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ctx->ev_valid = false;
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_counter, 0, kCountersPerWave * sizeof(unsigned long long), st));
  sfb::BatchArgs a{};
  a.src_base = src_base;
  a.src_off = src_off;
  a.src_len = src_len;
  a.dst_base = nullptr;  // the counting writer never touches dst, its capacities or the bitmap
  a.dst_delta = 0;
  a.dst_off = nullptr;
  a.dst_cap = nullptr;
  a.status = status;
  a.written = size;
  a.n = n;
  a.idx_base = 0;
  a.group_counter = ctx->d_counter + 1;
  a.lens_scratch = ctx->d_lens;
  a.match_bits = nullptr;
  a.defer_list = nullptr;
  a.defer_count = nullptr;
  a.todo_list = nullptr;
  a.todo_count = nullptr;
  a.no_pair = ctx->no_pair;
  a.handled = nullptr;
  a.start_bit = nullptr;
  a.start_out = nullptr;
  a.blk_end = nullptr;
  auto kern = sfb::huff_lanes_kernel<LaneCfg, true>;
  if (!ctx->count_configured) {  // (function attributes are per device)
    SFB_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LaneCfg::SMEM_BYTES));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    ctx->count_configured = true;
  }
  const uint64_t groups = (n + 31) / 32;
  const uint64_t want = (groups + LaneCfg::WARPS - 1) / LaneCfg::WARPS;
  const uint64_t resident = static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->ctas_per_sm);
  const unsigned grid = static_cast<unsigned>(want < resident ? want : resident);
  kern<<<grid, LaneCfg::WARPS * 32, LaneCfg::SMEM_BYTES, st>>>(a);
  SFB_TRY(ctx, cudaGetLastError());
  ctx->launches += 1;
  return SFB200_RC_OK;
}

int sfb200_last_pass_ms(sfb200_ctx* ctx, float* out3)
{
  if (!ctx || !out3 || !ctx->ev_valid) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
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
  SFB_ENTER(ctx);
  const unsigned threads = 128;
  const uint64_t blocks = (n * 32 + threads - 1) / threads;
  sfb::checksum_kernel<<<static_cast<unsigned>(blocks), threads, 0,
                         static_cast<cudaStream_t>(cuda_stream)>>>(base, off, len, out, n);
  SFB_TRY(ctx, cudaGetLastError());
  ctx->launches += 1;
  return SFB200_RC_OK;
}

// Host buffers, in BOUNDED device memory.  The batch is cut into sub-batches of consecutive
// streams that flow through a ring of three staging slots (SFB200_HOST_STAGING_MB, default 3456:
// per slot 768 MiB of dst and 384 MiB of src) on four CUDA streams:
//   s_in   H2D of the sub-batch's compressed bytes and (rebased) offsets
//   s_run  the kernels (sfb200_decompress_batch_device on the slot's buffers)
//   s_meta status / written of the sub-batch -> pinned host memory
//   s_out  D2H of the bytes the decoder produced
// H2D of sub-batch k+2, the kernels of k+1 and the D2H of k overlap, so the call is bounded by the
// slower PCIe direction rather than by the sum of the stages, and a batch far larger than the GPU's
// memory decodes in the same ~3.4 GiB.  dst is NOT uploaded: only bytes the decoder produced are
// copied back (whole runs of adjacent, completely filled regions in one copy; for a stream that
// stopped early just its `written` prefix), so everything else in the caller's buffer keeps its
// value, as the reference guarantees.  A single stream larger than a slot makes the slots grow to
// hold it (one stream is one unit of work: SURVEY.md §8 f3 has the chunked single stream as next).
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
  uint64_t max_src = 0, max_dst = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (src_off[i] > src_bytes || src_len[i] > src_bytes - src_off[i]) return SFB200_RC_BAD_ARGUMENT;
    if (dst_off[i] > dst_bytes || dst_cap[i] > dst_bytes - dst_off[i]) return SFB200_RC_BAD_ARGUMENT;
    max_src = std::max(max_src, src_len[i]);
    max_dst = std::max(max_dst, dst_cap[i]);
  }
  SFB_ENTER(ctx);
  constexpr int S = 3;
  // ---- slot sizes ------------------------------------------------------------------------------
  uint64_t budget = 3456ull << 20;
  if (const char* e = std::getenv("SFB200_HOST_STAGING_MB")) {
    const long v = std::atol(e);
    if (v > 0) budget = static_cast<uint64_t>(v) << 20;
  }
  if (const char* e = std::getenv("SFB200_HOST_STAGING_KB")) {  // (tests: force many sub-batches)
    const long v = std::atol(e);
    if (v > 0) budget = static_cast<uint64_t>(v) << 10;
  }
  uint64_t want_dst = std::max<uint64_t>(budget / S * 2 / 3, 4096), want_src = std::max<uint64_t>(budget / S / 3, 4096);
  // (a stream is staged whole; 256 bytes of slack for the 128-byte rebasing of the windows)
  want_dst = std::max(want_dst, max_dst + 256);
  want_src = std::max(want_src, max_src + 256);
  // no more than the batch needs
  want_dst = std::min(want_dst, dst_bytes + 256);
  want_src = std::min(want_src, src_bytes + 256);
  const uint64_t want_n = std::min<uint64_t>(n, 1ull << 22);  // streams per sub-batch
  auto sync_all = [&]() {
    for (auto& st : ctx->hs)
      if (st) cudaStreamSynchronize(st);
    if (ctx->s_meta) cudaStreamSynchronize(ctx->s_meta);
  };
  const uint64_t slot_src = want_src, slot_dst = want_dst;  // (what a sub-batch may span; the slots may be larger already)

  // ---- sub-batches of consecutive streams whose src and dst windows fit a slot -------------------
  struct Sub {
    uint64_t first, count, src_lo, src_hi, dst_lo, dst_hi;
  };
  std::vector<Sub> subs;
  // (the first sub-batches are an eighth, a quarter and a half of a slot: the download — the slower
  //  PCIe direction, which bounds the call — starts after the upload and the kernels of 1/8 of a
  //  slot instead of a whole one, and from then on never waits: a sub-batch's kernels take a
  //  quarter of the time its predecessor's download does even when it is twice as large)
  for (uint64_t i = 0; i < n;) {
    Sub sb{i, 0, ~0ull, 0, ~0ull, 0};
    const uint64_t ramp = subs.size() == 0 ? 8 : subs.size() == 1 ? 4 : subs.size() == 2 ? 2 : 1;
    const uint64_t lim_src = slot_src / ramp, lim_dst = slot_dst / ramp;
    while (i < n && sb.count < want_n) {
      const uint64_t slo = std::min(sb.src_lo, src_off[i]), shi = std::max(sb.src_hi, src_off[i] + src_len[i]);
      const uint64_t dlo = std::min(sb.dst_lo, dst_off[i]) & ~127ull, dhi = std::max(sb.dst_hi, dst_off[i] + dst_cap[i]);
      if (sb.count && (shi - slo > lim_src || dhi - dlo > lim_dst)) break;
      sb.src_lo = slo;
      sb.src_hi = shi;
      sb.dst_lo = dlo;  // (128-byte aligned: keeps the windows' alignment and the bitmap phase of the caller's layout)
      sb.dst_hi = dhi;
      ++sb.count;
      ++i;
    }
    subs.push_back(sb);
  }
  // ---- the slots this call needs (grown, never shrunk; a call with one sub-batch takes one) -----
  const int need_slots = static_cast<int>(std::min<size_t>(S, subs.size()));
  if (slot_dst > ctx->st_dst_cap || slot_src > ctx->st_src_cap) {
    sync_all();
    for (int k = 0; k < S; ++k) {
      if (ctx->st_src[k]) SFB_TRY(ctx, cudaFree(ctx->st_src[k]));
      if (ctx->st_dst[k]) SFB_TRY(ctx, cudaFree(ctx->st_dst[k]));
      ctx->st_src[k] = ctx->st_dst[k] = nullptr;
    }
    ctx->st_slots = 0;
    ctx->st_src_cap = std::max(slot_src, ctx->st_src_cap);
    ctx->st_dst_cap = std::max(slot_dst, ctx->st_dst_cap);
  }
  for (; ctx->st_slots < need_slots; ++ctx->st_slots) {
    const int k = ctx->st_slots;
    SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->st_src[k]), ctx->st_src_cap + 64));
    SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->st_dst[k]), ctx->st_dst_cap + 64));
  }
  if (want_n > ctx->st_meta_n) {
    sync_all();
    for (int k = 0; k < S; ++k) {
      if (ctx->st_meta[k]) SFB_TRY(ctx, cudaFree(ctx->st_meta[k]));
      if (ctx->st_hmeta[k]) SFB_TRY(ctx, cudaFreeHost(ctx->st_hmeta[k]));
      ctx->st_meta[k] = ctx->st_hmeta[k] = nullptr;
    }
    ctx->st_meta_n = 0;
    const uint64_t cap = want_n + want_n / 8 + 64;
    for (int k = 0; k < S; ++k) {
      SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->st_meta[k]), cap * (5 * 8 + 1) + 64));
      SFB_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->st_hmeta[k]), cap * (5 * 8 + 1) + 64));
    }
    ctx->st_meta_n = cap;
  }
  for (auto& st : ctx->hs)
    if (!st) SFB_TRY(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  if (!ctx->s_meta) SFB_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_meta, cudaStreamNonBlocking));
  for (int k = 0; k < S; ++k)
    for (auto& e : ctx->st_ev[k])
      if (!e) SFB_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaStream_t s_in = ctx->hs[0], s_run = ctx->hs[1], s_out = ctx->hs[2], s_meta = ctx->s_meta;
  const uint64_t cap_n = ctx->st_meta_n;
  // the scratch of the device call is sized once, for the largest sub-batch (no reallocation —
  // cudaFree synchronises the device — while the pipeline runs)
  {
    uint64_t big = 0, big_n = 0;
    for (const Sub& sb : subs) {
      big = std::max(big, sb.dst_hi - sb.dst_lo);
      big_n = std::max(big_n, sb.count);
    }
    const uint64_t bits_words = (big + 128) / 32 + 8;
    if (bits_words > ctx->d_bits_words) {
      sync_all();
      if (ctx->d_bits) SFB_TRY(ctx, cudaFree(ctx->d_bits));
      ctx->d_bits = nullptr;
      ctx->d_bits_words = 0;
      const uint64_t want = bits_words + bits_words / 8 + 64;
      SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_bits), want * sizeof(uint32_t)));
      ctx->d_bits_words = want;
    }
    if (big_n > ctx->d_order_n) {
      sync_all();
      if (ctx->d_order) SFB_TRY(ctx, cudaFree(ctx->d_order));
      ctx->d_order = nullptr;
      ctx->d_order_n = 0;
      const uint64_t want = big_n + big_n / 8 + 64;
      SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_order), want * (sizeof(uint32_t) + 1)));
      ctx->d_order_n = want;
    }
  }
  auto cleanup = [&](int code) {
    sync_all();
    return code;
  };
#define SFB_TRYC(call)                                           \
  do {                                                           \
    const cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return cleanup(fail(ctx, e_, #call)); \
  } while (0)
  // slot k % S: device arrays and their pinned mirror
  auto d_arr = [&](int slot, int which) { return ctx->st_meta[slot] + static_cast<uint64_t>(which) * cap_n; };
  auto h_arr = [&](int slot, int which) { return ctx->st_hmeta[slot] + static_cast<uint64_t>(which) * cap_n; };
  auto enqueue = [&](size_t k) -> int {
    const Sub& sb = subs[k];
    const int slot = static_cast<int>(k % S);
    const uint64_t f = sb.first, c = sb.count;
    // the slot is free once the bytes of its previous occupant have left for the host
    // (and the host has read that sub-batch's status / written from the pinned mirror: it has —
    //  sub-batch k - S was finished before this one is enqueued)
    if (k >= static_cast<size_t>(S)) SFB_TRYC(cudaStreamWaitEvent(s_in, ctx->st_ev[slot][3], 0));
    uint64_t* h_so = h_arr(slot, 0), * h_sl = h_arr(slot, 1), * h_do = h_arr(slot, 2), * h_dc = h_arr(slot, 3);
    for (uint64_t j = 0; j < c; ++j) {
      h_so[j] = src_off[f + j] - sb.src_lo;
      h_sl[j] = src_len[f + j];
      h_do[j] = dst_off[f + j] - sb.dst_lo;
      h_dc[j] = dst_cap[f + j];
    }
    if (sb.src_hi > sb.src_lo)
      SFB_TRYC(cudaMemcpyAsync(ctx->st_src[slot], src + sb.src_lo, sb.src_hi - sb.src_lo, cudaMemcpyHostToDevice, s_in));
    // (the four arrays are contiguous in both mirrors when c == cap_n; copy them one by one)
    for (int w = 0; w < 4; ++w)
      SFB_TRYC(cudaMemcpyAsync(d_arr(slot, w), h_arr(slot, w), c * 8, cudaMemcpyHostToDevice, s_in));
    SFB_TRYC(cudaEventRecord(ctx->st_ev[slot][0], s_in));
    SFB_TRYC(cudaStreamWaitEvent(s_run, ctx->st_ev[slot][0], 0));
    uint8_t* const d_status = reinterpret_cast<uint8_t*>(d_arr(slot, 5));
    const int r = sfb200_decompress_batch_device(ctx, ctx->st_src[slot], d_arr(slot, 0), d_arr(slot, 1), ctx->st_dst[slot],
                                                 sb.dst_hi - sb.dst_lo, d_arr(slot, 2), d_arr(slot, 3), d_status,
                                                 d_arr(slot, 4), c, s_run);
    if (r) return cleanup(r);
    SFB_TRYC(cudaEventRecord(ctx->st_ev[slot][1], s_run));
    SFB_TRYC(cudaStreamWaitEvent(s_meta, ctx->st_ev[slot][1], 0));
    SFB_TRYC(cudaMemcpyAsync(h_arr(slot, 4), d_arr(slot, 4), c * 8, cudaMemcpyDeviceToHost, s_meta));
    SFB_TRYC(cudaMemcpyAsync(h_arr(slot, 5), d_status, c, cudaMemcpyDeviceToHost, s_meta));
    SFB_TRYC(cudaEventRecord(ctx->st_ev[slot][2], s_meta));
    return SFB200_RC_OK;
  };
  int rc = SFB200_RC_OK;
  size_t enq = 0;
  for (; enq < subs.size() && enq < static_cast<size_t>(S) - 1; ++enq) {
    rc = enqueue(enq);
    if (rc) return rc;
  }
  for (size_t k = 0; k < subs.size(); ++k) {
    // keep S - 1 sub-batches in flight behind this one (slot (k + S - 1) % S is the one sub-batch
    // k - 1 has just left)
    if (enq < subs.size()) {
      rc = enqueue(enq++);
      if (rc) return rc;
    }
    const Sub& sb = subs[k];
    const int slot = static_cast<int>(k % S);
    SFB_TRYC(cudaEventSynchronize(ctx->st_ev[slot][2]));
    const uint64_t* const wr_host = h_arr(slot, 4);
    const uint8_t* const st_host = reinterpret_cast<const uint8_t*>(h_arr(slot, 5));
    // bytes produced -> the caller's buffer, in as few copies as the layout allows
    SFB_TRYC(cudaStreamWaitEvent(s_out, ctx->st_ev[slot][1], 0));
    uint64_t run_lo = 0, run_hi = 0;  // pending run [run_lo, run_hi) of produced bytes (caller's offsets)
    auto flush_run = [&]() -> int {
      if (run_hi > run_lo)
        SFB_TRYC(cudaMemcpyAsync(dst + run_lo, ctx->st_dst[slot] + (run_lo - sb.dst_lo), run_hi - run_lo,
                                 cudaMemcpyDeviceToHost, s_out));
      run_lo = run_hi = 0;
      return SFB200_RC_OK;
    };
    for (uint64_t j = 0; j < sb.count; ++j) {
      const uint64_t i = sb.first + j;
      const uint64_t w = wr_host[j];
      status[i] = st_host[j];
      if (written) written[i] = w;
      if (w == 0) continue;
      if (run_hi > run_lo && dst_off[i] == run_hi) {
        run_hi += w;
      } else {
        rc = flush_run();
        if (rc) return rc;
        run_lo = dst_off[i];
        run_hi = run_lo + w;
      }
      if (w != dst_cap[i]) {  // the rest of this region is not ours: the run ends here
        rc = flush_run();
        if (rc) return rc;
      }
    }
    rc = flush_run();
    if (rc) return rc;
    SFB_TRYC(cudaEventRecord(ctx->st_ev[slot][3], s_out));
  }
#undef SFB_TRYC
  return cleanup(SFB200_RC_OK);
}

uint64_t sfb200_staging_bytes(const sfb200_ctx* ctx)
{
  return ctx ? static_cast<uint64_t>(ctx->st_slots) * (ctx->st_src_cap + ctx->st_dst_cap + 128) : 0;
}

// One batch, several GPUs: the streams are cut into contiguous shards balanced by the bytes they
// move (sum of src_len + dst_cap) and every context decodes its shard through
// sfb200_decompress_batch_host on a thread of its own.  No communication between the devices:
// streams are independent.  Returns the first non-zero infrastructure code, if any.
int sfb200_decompress_batch_host_multi(sfb200_ctx* const* ctxs, int n_ctx, const uint8_t* src, uint64_t src_bytes,
                                       const uint64_t* src_off, const uint64_t* src_len, uint8_t* dst,
                                       uint64_t dst_bytes, const uint64_t* dst_off, const uint64_t* dst_cap,
                                       uint8_t* status, uint64_t* written, uint64_t n)
{
  if (!ctxs || n_ctx <= 0) return SFB200_RC_BAD_ARGUMENT;
  for (int k = 0; k < n_ctx; ++k)
    if (!ctxs[k]) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !status) return SFB200_RC_BAD_ARGUMENT;
  if (n_ctx == 1)
    return sfb200_decompress_batch_host(ctxs[0], src, src_bytes, src_off, src_len, dst, dst_bytes, dst_off, dst_cap,
                                        status, written, n);
  std::vector<uint64_t> cut(static_cast<size_t>(n_ctx) + 1, n);
  sfb200_partition_streams(src_len, dst_cap, n, n_ctx, cut.data());
  std::vector<int> rcs(static_cast<size_t>(n_ctx), SFB200_RC_OK);
  std::vector<std::thread> th;
  for (int k = 0; k < n_ctx; ++k) {
    const uint64_t lo = cut[k], cnt = cut[k + 1] - cut[k];
    if (cnt == 0) continue;
    th.emplace_back([=, &rcs] {
      rcs[k] = sfb200_decompress_batch_host(ctxs[k], src, src_bytes, src_off + lo, src_len + lo, dst, dst_bytes,
                                            dst_off + lo, dst_cap + lo, status + lo, written ? written + lo : nullptr,
                                            cnt);
    });
  }
  for (auto& t : th) t.join();
  for (int r : rcs)
    if (r) return r;
  return SFB200_RC_OK;
}

// cut[0] = 0 <= cut[1] <= ... <= cut[parts] = n: contiguous shards with (nearly) equal sums of
// src_len + dst_cap.
void sfb200_partition_streams(const uint64_t* src_len, const uint64_t* dst_cap, uint64_t n, int parts, uint64_t* cut)
{
  long double total = 0;
  for (uint64_t i = 0; i < n; ++i) total += static_cast<long double>(src_len[i]) + static_cast<long double>(dst_cap[i]);
  cut[0] = 0;
  long double acc = 0;
  uint64_t i = 0;
  for (int k = 1; k < parts; ++k) {
    const long double goal = total * k / parts;
    while (i < n && acc + (static_cast<long double>(src_len[i]) + static_cast<long double>(dst_cap[i])) / 2 <= goal) {
      acc += static_cast<long double>(src_len[i]) + static_cast<long double>(dst_cap[i]);
      ++i;
    }
    cut[k] = i;
  }
  cut[parts] = n;
}

int sfb200_decompress(sfb200_ctx* ctx, const uint8_t* src, size_t src_len, uint8_t* dst,
                      size_t dst_cap, uint8_t* status, uint64_t* written)
{
  if (!status) return SFB200_RC_BAD_ARGUMENT;
  const uint64_t zero = 0, sl = src_len, dc = dst_cap;
  return sfb200_decompress_batch_host(ctx, src, sl, &zero, &sl, dst, dc, &zero, &dc, status,
                                      written, 1);
}

// ---------------------------------------------------------------------------------------------
// Chunked input for ONE stream (the reference anticipates it at
// src/decompress.cpp:214).  The unit of progress is the BLOCK: a call decodes, with the lane kernel,
// every block that is complete in the input fed so far, keeps what those blocks produced and
// carries three things to the next call — the bytes from the next block header on (device), the bit
// offset of that header, and the last 32 KiB of output (the LZ77 window, in front of the output
// buffer, so that distances reach into it as they would in one call).
struct sfb200_inflate_stream {
  sfb200_ctx* ctx = nullptr;
  uint8_t* d_in[2] = {nullptr, nullptr};   // buffered input (the other one is the target of the next compaction)
  uint64_t d_in_cap[2] = {0, 0};
  int cur = 0;
  uint64_t in_len = 0;     // bytes buffered
  uint64_t bit = 0;        // bit offset of the next block header inside the buffer (0..7 between calls)
  uint8_t* d_out = nullptr;   // [32 KiB window | output of one call]
  uint64_t d_out_cap = 0;
  uint8_t* d_tmp = nullptr;   // 32 KiB (window moves that overlap)
  uint64_t hist = 0;       // valid window bytes, right-aligned in the first 32 KiB of d_out
  uint64_t* d_meta = nullptr;  // src_off, src_len, dst_off, dst_cap, written, start_bit, start_out, blk_end[2], status
  bool finished = false;
  uint8_t final_status = 0;
};

namespace {
constexpr uint64_t kWindow = 32768;
int keep_bytes(sfb200_ctx* ctx, uint8_t** p, uint64_t* cap, uint64_t need, uint64_t keep)
{
  if (need <= *cap) return SFB200_RC_OK;
  uint8_t* q = nullptr;
  const uint64_t want = need + need / 2 + 4096;
  SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&q), want));
  if (keep) SFB_TRY(ctx, cudaMemcpy(q, *p, keep, cudaMemcpyDeviceToDevice));
  if (*p) SFB_TRY(ctx, cudaFree(*p));
  *p = q;
  *cap = want;
  return SFB200_RC_OK;
}
}  // namespace

int sfb200_inflate_stream_create(sfb200_ctx* ctx, sfb200_inflate_stream** out)
{
  if (!ctx || !out) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  sfb200_inflate_stream* s = new (std::nothrow) sfb200_inflate_stream;
  if (!s) return SFB200_RC_OUT_OF_MEMORY;
  s->ctx = ctx;
  if (cudaMalloc(reinterpret_cast<void**>(&s->d_tmp), kWindow) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&s->d_meta), 16 * sizeof(uint64_t)) != cudaSuccess) {
    cudaFree(s->d_tmp);
    delete s;
    cudaGetLastError();
    return SFB200_RC_OUT_OF_MEMORY;
  }
  *out = s;
  return SFB200_RC_OK;
}

void sfb200_inflate_stream_destroy(sfb200_inflate_stream* s)
{
  if (!s) return;
  DeviceGuard guard_(s->ctx->device);
  cudaFree(s->d_in[0]);
  cudaFree(s->d_in[1]);
  cudaFree(s->d_out);
  cudaFree(s->d_tmp);
  cudaFree(s->d_meta);
  delete s;
}

int sfb200_inflate_stream_feed(sfb200_inflate_stream* s, const uint8_t* src, size_t src_len, int last, uint8_t* dst,
                               size_t dst_cap, uint64_t* written, uint8_t* status, int* finished)
{
  if (!s || !status || !written || !finished || (src_len && !src) || (dst_cap && !dst)) return SFB200_RC_BAD_ARGUMENT;
  sfb200_ctx* const ctx = s->ctx;
  SFB_ENTER(ctx);
  *written = 0;
  if (s->finished) {
    *status = s->final_status;
    *finished = 1;
    return SFB200_RC_OK;
  }
  *finished = 0;
  *status = SFB200_SUCCESS;
  if (s->in_len + src_len >= 0xffffff00ull || kWindow + dst_cap >= 0xffffff00ull) return SFB200_RC_BAD_ARGUMENT;
  // the new input goes behind what is buffered
  int rc = keep_bytes(ctx, &s->d_in[s->cur], &s->d_in_cap[s->cur], s->in_len + src_len + 64, s->in_len);
  if (rc != SFB200_RC_OK) return rc;
  if (src_len) SFB_TRY(ctx, cudaMemcpy(s->d_in[s->cur] + s->in_len, src, src_len, cudaMemcpyHostToDevice));
  s->in_len += src_len;
  if (dst_cap == 0 && !last) return SFB200_RC_OK;  // (only buffering)
  rc = keep_bytes(ctx, &s->d_out, &s->d_out_cap, kWindow + dst_cap + 256, s->d_out ? kWindow : 0);
  if (rc != SFB200_RC_OK) return rc;
  // one lane-kernel decode from the pending block header, behind the window
  uint64_t h[16] = {0};
  h[0] = 0;                        // src_off
  h[1] = s->in_len;                // src_len
  h[2] = kWindow - s->hist;        // dst_off
  h[3] = s->hist + dst_cap;        // dst_cap
  h[4] = s->hist;                  // written (overwritten)
  h[5] = s->bit;                   // start_bit
  h[6] = s->hist;                  // start_out
  h[7] = s->bit;                   // blk_end: bit, output position (stay like this when no header is reached)
  h[8] = s->hist;
  h[9] = SFB200_ERROR;             // status (first byte)
  SFB_TRY(ctx, cudaMemcpy(s->d_meta, h, sizeof h, cudaMemcpyHostToDevice));
  Resume rs;
  rs.start_bit = s->d_meta + 5;
  rs.start_out = s->d_meta + 6;
  rs.blk_end = s->d_meta + 7;
  rc = batch_device_impl(ctx, s->d_in[s->cur], s->d_meta + 0, s->d_meta + 1, s->d_out, kWindow + dst_cap, s->d_meta + 2,
                         s->d_meta + 3, reinterpret_cast<uint8_t*>(s->d_meta + 9), s->d_meta + 4, 1, nullptr, &rs);
  if (rc != SFB200_RC_OK) return rc;
  SFB_TRY(ctx, cudaMemcpy(h, s->d_meta, sizeof h, cudaMemcpyDeviceToHost));
  const uint8_t st = static_cast<uint8_t>(h[9] & 0xffu);
  const uint64_t total = h[4], be_bit = h[7], be_out = h[8];
  uint64_t produced = 0;
  if (st == SFB200_SUCCESS || (last && st != SFB200_DST_TOO_SMALL)) {
    // the final block was decoded, or no more input will come: the stream ends as one call would end it
    produced = total - s->hist;
    s->finished = true;
    s->final_status = st;
    *status = st;
    *finished = 1;
  } else if (st == SFB200_DST_TOO_SMALL && be_out == s->hist) {
    *status = SFB200_DST_TOO_SMALL;  // not even the next block fits dst: call again with more room (nothing is lost)
    return SFB200_RC_OK;
  } else {
    // ran out of input (or of room) inside a block: keep the complete blocks, redo that one next time
    produced = be_out - s->hist;
    s->bit = be_bit;
  }
  if (produced) SFB_TRY(ctx, cudaMemcpy(dst, s->d_out + kWindow, produced, cudaMemcpyDeviceToHost));
  *written = produced;
  if (!s->finished) {
    // the window: the last 32 KiB of everything produced so far
    const uint64_t have = s->hist + produced;
    const uint64_t keep = have < kWindow ? have : kWindow;
    if (produced) {
      const uint8_t* from = s->d_out + kWindow + produced - keep;
      uint8_t* to = s->d_out + kWindow - keep;
      if (produced >= kWindow) {
        SFB_TRY(ctx, cudaMemcpy(to, from, keep, cudaMemcpyDeviceToDevice));
      } else {
        SFB_TRY(ctx, cudaMemcpy(s->d_tmp, from, keep, cudaMemcpyDeviceToDevice));
        SFB_TRY(ctx, cudaMemcpy(to, s->d_tmp, keep, cudaMemcpyDeviceToDevice));
      }
    }
    s->hist = keep;
    // the input: drop the bytes in front of the pending block header
    const uint64_t drop = s->bit >> 3;
    if (drop) {
      const int other = s->cur ^ 1;
      const uint64_t rest = s->in_len - drop;
      rc = keep_bytes(ctx, &s->d_in[other], &s->d_in_cap[other], rest + 64, 0);
      if (rc != SFB200_RC_OK) return rc;
      if (rest) SFB_TRY(ctx, cudaMemcpy(s->d_in[other], s->d_in[s->cur] + drop, rest, cudaMemcpyDeviceToDevice));
      s->cur = other;
      s->in_len = rest;
      s->bit &= 7u;
    }
  }
  return SFB200_RC_OK;
}


// ---------------------------------------------------------------------------------------------
// Batched compression (deflate_compress.cuh)
uint64_t sfb200_compress_bound(uint64_t src_len)
{
  // stored blocks are the fallback: 5 header bytes per 65 535 payload bytes
  return src_len + 5ull * (src_len ? (src_len + 65534ull) / 65535ull : 1ull);
}

int sfb200_compress_batch_device(sfb200_ctx* ctx, const uint8_t* src_base, const uint64_t* src_off,
                                 const uint64_t* src_len, uint8_t* dst_base, const uint64_t* dst_off,
                                 const uint64_t* dst_cap, uint8_t* status, uint64_t* written, uint64_t n,
                                 void* cuda_stream)
{
  if (!ctx) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !status || !written) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (!ctx->compress_configured) {  // (function attributes are per device)
    SFB_TRY(ctx, cudaFuncSetAttribute(sfb::deflate_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      sfb::CMP_SMEM_BYTES));
    int per_sm = 0;
    SFB_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sfb::deflate_compress_kernel,
                                                               sfb::CMP_WARPS * 32, sfb::CMP_SMEM_BYTES));
    ctx->compress_ctas_per_sm = per_sm < 1 ? 1 : per_sm;
    ctx->compress_configured = true;
  }
  SFB_TRY(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), st));
  sfb::CompressArgs a;
  a.src_base = src_base;
  a.src_off = src_off;
  a.src_len = src_len;
  a.dst_base = dst_base;
  a.dst_off = dst_off;
  a.dst_cap = dst_cap;
  a.status = status;
  a.written = written;
  a.n = n;
  a.counter = ctx->d_counter;
  a.fixed_only = 0;   // SFB200_COMPRESS_FIXED=1: the first generation (one pass, fixed-Huffman blocks only) for A/B runs
  if (const char* e = std::getenv("SFB200_COMPRESS_FIXED")) a.fixed_only = e[0] == '1';
  const uint64_t want = (n + sfb::CMP_WARPS - 1) / sfb::CMP_WARPS;
  const uint64_t resident = static_cast<uint64_t>(ctx->sm_count) * static_cast<uint64_t>(ctx->compress_ctas_per_sm);
  sfb::deflate_compress_kernel<<<static_cast<unsigned>(want < resident ? want : resident), sfb::CMP_WARPS * 32,
                                 sfb::CMP_SMEM_BYTES, st>>>(a);
  SFB_TRY(ctx, cudaGetLastError());
  ctx->launches += 1;
  return SFB200_RC_OK;
}

int sfb200_compress(sfb200_ctx* ctx, const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap,
                    uint8_t* status, uint64_t* written)
{
  if (!ctx || !status || !written || (src_len && !src) || (dst_cap && !dst)) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  uint8_t* buf = nullptr;
  const uint64_t in_pad = (static_cast<uint64_t>(src_len) + 255) & ~255ull;
  const uint64_t total = in_pad + dst_cap + 256 + 64;
  SFB_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&buf), total));
  auto done = [&](int rc) {
    cudaFree(buf);
    return rc;
  };
  uint64_t h[6] = {0, src_len, 0, dst_cap, 0, 0};
  uint64_t* meta = reinterpret_cast<uint64_t*>(buf + in_pad + ((dst_cap + 255) & ~255ull));
  if (src_len && cudaMemcpy(buf, src, src_len, cudaMemcpyHostToDevice) != cudaSuccess)
    return done(fail(ctx, cudaGetLastError(), "cudaMemcpy(src)"));
  if (cudaMemcpy(meta, h, sizeof h, cudaMemcpyHostToDevice) != cudaSuccess)
    return done(fail(ctx, cudaGetLastError(), "cudaMemcpy(meta)"));
  int rc = sfb200_compress_batch_device(ctx, buf, meta + 0, meta + 1, buf + in_pad, meta + 2, meta + 3,
                                        reinterpret_cast<uint8_t*>(meta + 5), meta + 4, 1, nullptr);
  if (rc != SFB200_RC_OK) return done(rc);
  if (cudaMemcpy(h, meta, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess)
    return done(fail(ctx, cudaGetLastError(), "cudaMemcpy(meta back)"));
  *status = static_cast<uint8_t>(h[5] & 0xffu);
  *written = h[4];
  if (h[4] && cudaMemcpy(dst, buf + in_pad, h[4], cudaMemcpyDeviceToHost) != cudaSuccess)
    return done(fail(ctx, cudaGetLastError(), "cudaMemcpy(dst)"));
  return done(SFB200_RC_OK);
}


int sfb200_decompress_container_batch_device(sfb200_ctx* ctx, int container, const uint8_t* src_base,
                                             const uint64_t* src_off, const uint64_t* src_len,
                                             uint8_t* dst_base, uint64_t dst_bytes,
                                             const uint64_t* dst_off, const uint64_t* dst_cap,
                                             uint8_t* status, uint64_t* written, uint64_t n,
                                             void* cuda_stream)
{
  if (!ctx || container < 0 || container > 3) return SFB200_RC_BAD_ARGUMENT;
  if (n == 0) return SFB200_RC_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !status) return SFB200_RC_BAD_ARGUMENT;
  if (container == SFB200_CONTAINER_RAW)
    return sfb200_decompress_batch_device(ctx, src_base, src_off, src_len, dst_base, dst_bytes, dst_off, dst_cap,
                                          status, written, n, cuda_stream);
  SFB_ENTER(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  int rc = grow(ctx, &ctx->d_cont, &ctx->d_cont_cap, n * 36 + 64);
  if (rc != SFB200_RC_OK) return rc;
  sfb::ContainerArgs c;
  c.src_base = src_base;
  c.src_off = src_off;
  c.src_len = src_len;
  c.container = static_cast<uint32_t>(container);
  c.n = n;
  c.pay_off = reinterpret_cast<uint64_t*>(ctx->d_cont);
  c.pay_len = c.pay_off + n;
  uint64_t* const own_written = c.pay_len + n;
  c.kind = reinterpret_cast<uint32_t*>(own_written + n);
  c.expect = c.kind + n;
  c.isize = c.expect + n;
  c.dst_base = dst_base;
  c.dst_off = dst_off;
  c.status = status;
  c.written = written ? written : own_written;
  const unsigned pgrid = static_cast<unsigned>((n + 127) / 128);
  sfb::container_parse_kernel<<<pgrid, 128, 0, st>>>(c);
  SFB_TRY(ctx, cudaGetLastError());
  rc = sfb200_decompress_batch_device(ctx, src_base, c.pay_off, c.pay_len, dst_base, dst_bytes, dst_off, dst_cap,
                                      status, c.written, n, cuda_stream);
  if (rc != SFB200_RC_OK) return rc;
  if (!ctx->check_configured) {  // (function attributes are per device)
    SFB_TRY(ctx, cudaFuncSetAttribute(sfb::container_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      sfb::CHECK_SMEM_BYTES));
    ctx->check_configured = true;
  }
  const uint64_t want = (n * 32 + sfb::CHECK_THREADS - 1) / sfb::CHECK_THREADS;
  const uint64_t resident = static_cast<uint64_t>(ctx->sm_count);  // (one block per SM: 128 KiB of tables)
  sfb::container_check_kernel<<<static_cast<unsigned>(want < resident ? want : resident), sfb::CHECK_THREADS,
                                sfb::CHECK_SMEM_BYTES, st>>>(c);
  SFB_TRY(ctx, cudaGetLastError());
  ctx->launches += 2;
  return SFB200_RC_OK;
}

int sfb200_decompress_container(sfb200_ctx* ctx, int container, const uint8_t* src, size_t src_len,
                                uint8_t* dst, size_t dst_cap, uint8_t* status, uint64_t* written)
{
  if (!ctx || !status || (!src && src_len) || (!dst && dst_cap)) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  const uint64_t meta_at = (static_cast<uint64_t>(src_len) + 15) & ~15ull;
  int rc = grow(ctx, &ctx->d_src, &ctx->d_src_cap, meta_at + 64);
  if (rc != SFB200_RC_OK) return rc;
  rc = grow(ctx, &ctx->d_dst, &ctx->d_dst_cap, static_cast<uint64_t>(dst_cap) + 64);
  if (rc != SFB200_RC_OK) return rc;
  // metadata behind the stream in the staging buffer: src_off, src_len, dst_off, dst_cap, written (u64), status (u8)
  uint64_t meta[5] = {0, static_cast<uint64_t>(src_len), 0, static_cast<uint64_t>(dst_cap), 0};
  uint64_t* const d_m = reinterpret_cast<uint64_t*>(ctx->d_src + meta_at);
  uint8_t* const d_st = reinterpret_cast<uint8_t*>(d_m + 5);
  if (src_len) SFB_TRY(ctx, cudaMemcpyAsync(ctx->d_src, src, src_len, cudaMemcpyHostToDevice, nullptr));
  SFB_TRY(ctx, cudaMemcpyAsync(d_m, meta, sizeof(meta), cudaMemcpyHostToDevice, nullptr));
  rc = sfb200_decompress_container_batch_device(ctx, container, ctx->d_src, d_m, d_m + 1, ctx->d_dst, dst_cap, d_m + 2,
                                                d_m + 3, d_st, d_m + 4, 1, nullptr);
  if (rc != SFB200_RC_OK) return rc;
  uint64_t wr = 0;
  SFB_TRY(ctx, cudaMemcpyAsync(&wr, d_m + 4, 8, cudaMemcpyDeviceToHost, nullptr));
  SFB_TRY(ctx, cudaMemcpyAsync(status, d_st, 1, cudaMemcpyDeviceToHost, nullptr));
  SFB_TRY(ctx, cudaStreamSynchronize(nullptr));
  if (wr > dst_cap) return SFB200_RC_CUDA_ERROR;  // (cannot happen)
  if (wr) SFB_TRY(ctx, cudaMemcpy(dst, ctx->d_dst, wr, cudaMemcpyDeviceToHost));  // only what was produced
  if (written) *written = wr;
  return SFB200_RC_OK;
}

int sfb200_decompressed_size(sfb200_ctx* ctx, const uint8_t* src, size_t src_len, uint8_t* status,
                             uint64_t* size)
{
  if (!ctx || !status || !size || (!src && src_len)) return SFB200_RC_BAD_ARGUMENT;
  SFB_ENTER(ctx);
  int rc = grow(ctx, &ctx->d_src, &ctx->d_src_cap, static_cast<uint64_t>(src_len) + 64);
  if (rc != SFB200_RC_OK) return rc;
  // metadata lives behind the stream in the same staging buffer: off, len, size (u64), status (u8)
  const uint64_t meta_at = (static_cast<uint64_t>(src_len) + 15) & ~15ull;
  uint64_t meta[3] = {0, static_cast<uint64_t>(src_len), 0};
  uint64_t* const d_m = reinterpret_cast<uint64_t*>(ctx->d_src + meta_at);
  uint8_t* const d_st = reinterpret_cast<uint8_t*>(d_m + 3);
  if (src_len) SFB_TRY(ctx, cudaMemcpyAsync(ctx->d_src, src, src_len, cudaMemcpyHostToDevice, nullptr));
  SFB_TRY(ctx, cudaMemcpyAsync(d_m, meta, sizeof(meta), cudaMemcpyHostToDevice, nullptr));
  rc = sfb200_decompressed_size_batch_device(ctx, ctx->d_src, d_m, d_m + 1, d_st, d_m + 2, 1, nullptr);
  if (rc != SFB200_RC_OK) return rc;
  SFB_TRY(ctx, cudaMemcpyAsync(size, d_m + 2, 8, cudaMemcpyDeviceToHost, nullptr));
  SFB_TRY(ctx, cudaMemcpyAsync(status, d_st, 1, cudaMemcpyDeviceToHost, nullptr));
  SFB_TRY(ctx, cudaStreamSynchronize(nullptr));
  return SFB200_RC_OK;
}

}  // extern "C"
