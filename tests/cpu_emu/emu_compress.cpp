// TEST INFRASTRUCTURE ONLY — the compression kernel (deflate_compress.cuh: one warp per stream) run
// on the host, one warp = 32 threads in lock-step through real barriers (see cuda_shim_warp.h).
#define SFB_CPU_EMU 1
#include "cuda_shim_warp.h"

#include <cstdlib>
#include <thread>
#include <vector>

static uint8_t* emu_cmp_smem = nullptr;  // one warp's shared memory, seen by all its lanes
#define SFB_EMU_SMEM emu_cmp_smem
#include "../../starflate_b200/csrc/deflate_compress.cuh"

// one stream -> (status, written); dst gets the raw-DEFLATE bytes
extern "C" int emu_compress(const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t dst_cap, uint64_t* written)
{
  EmuWarp warp;
  emu_warp = &warp;
  std::vector<uint8_t> smem(sfb::CMP_WARP_BYTES + 64, 0xDD);
  emu_cmp_smem = smem.data();
  // the kernel reads src in aligned words that hold at least one stream byte: give it a padded,
  // word-aligned copy at the caller's phase
  std::vector<uint8_t> sbuf(src_len + 32, 0xEE);
  uint8_t* sp = sbuf.data() + 8;
  while ((reinterpret_cast<uintptr_t>(sp) & 3u) != (reinterpret_cast<uintptr_t>(src) & 3u)) ++sp;
  if (src_len) std::memcpy(sp, src, src_len);
  unsigned long long counter = 0;
  const uint64_t zero = 0;
  uint8_t st = 0xEE;
  sfb::CompressArgs a;
  a.src_base = sp;
  a.src_off = &zero;
  a.src_len = &src_len;
  a.dst_base = dst;
  a.dst_off = &zero;
  a.dst_cap = &dst_cap;
  a.status = &st;
  a.written = written;
  a.n = 1;
  a.counter = &counter;
  a.fixed_only = std::getenv("SFB200_COMPRESS_FIXED") != nullptr;
  std::vector<std::thread> lanes;
  for (unsigned l = 0; l < 32; ++l)
    lanes.emplace_back([&a, l] {
      threadIdx.x = l;
      blockIdx.x = 0;
      blockDim.x = 32;
      gridDim.x = 1;
      sfb::deflate_compress_kernel(a);
    });
  for (auto& t : lanes) t.join();
  emu_warp = nullptr;
  return st;
}
