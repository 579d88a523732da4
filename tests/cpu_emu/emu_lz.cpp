// TEST INFRASTRUCTURE ONLY — the pass-2 kernel (lz_window.cuh) run on the host, one warp = 32
// threads in lock-step through real barriers (see cuda_shim_warp.h).
#define SFB_CPU_EMU 1
#include "cuda_shim_warp.h"

#include <thread>
#include <vector>

#include <cstdlib>
#include <cstdio>

#include "../../starflate_b200/csrc/lz_window.cuh"

extern "C" void emu_lz_resolve(uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* written,
                               const uint32_t* match_bits, uint64_t n)
{
  EmuWarp warp;
  emu_warp = &warp;
  unsigned long long counter = 0;
  sfb::ResolveArgs a;
  a.dst_base = dst_base;
  a.dst_delta = 0;
  a.dst_off = dst_off;
  a.written = written;
  a.match_bits = match_bits;
  a.n = n;
  a.idx_base = 0;
  a.todo_list = nullptr;
  a.stream_counter = &counter;
  std::vector<std::thread> lanes;
  for (unsigned l = 0; l < 32; ++l)
    lanes.emplace_back([&a, l] {
      threadIdx.x = l;
      blockIdx.x = 0;
      blockDim.x = 32;
      gridDim.x = 1;
      if (std::getenv("SFB_EMU_LZ_V1")) sfb::lz_resolve_kernel(a);  // (the first-generation kernel, kept for A/B)
      else sfb::lz_window_kernel<5>(a);
    });
  for (auto& t : lanes) t.join();
  emu_warp = nullptr;
}
