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
  sfb::ResolveArgs a{};
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

// The segment-wise form lz_window_queue_kernel uses (one stream; segments in order, state handed
// from one to the next through memory, reads bounded by (k + 1) SEG + LAG - 16 as on the device):
// the result must be what one lzw_run over the whole stream gives.
extern "C" void emu_lz_resolve_segments(uint8_t* dst_base, uint64_t dst_off, uint64_t written,
                                        const uint32_t* match_bits, uint32_t seg_shift, uint32_t lag)
{
  EmuWarp warp;
  emu_warp = &warp;
  uint32_t state[5] = {0, 0, 0, 0, 0};
  std::vector<std::thread> lanes;
  for (unsigned l = 0; l < 32; ++l)
    lanes.emplace_back([&, l] {
      threadIdx.x = l;
      blockIdx.x = 0;
      blockDim.x = 32;
      gridDim.x = 1;
      sfb::LzwView v;
      v.base = dst_base + (dst_off & ~1023ull);
      v.bits = match_bits + ((dst_off & ~1023ull) >> 5);
      v.q = static_cast<uint32_t>(dst_off & 1023u);
      v.lane = l;
      // pass 1 publishes segment k once written >= (k + 1) SEG + LAG (+ 8): those get their own item
      uint32_t k = 0;
      for (;; ++k) {
        sfb::LzwState s;
        if (k == 0) {
          s.W = 0; s.cur = v.q; s.c_o = 0; s.c_end = 0; s.c_d = 1;
        } else {
          s.W = state[0]; s.cur = state[1]; s.c_o = state[2]; s.c_end = state[3]; s.c_d = state[4];
        }
        const bool fin = written < (static_cast<uint64_t>(k + 1) << seg_shift) + lag + 8;
        if (fin) {
          v.end = v.q + static_cast<uint32_t>(written);
          if (written) sfb::lzw_run<true>(v, s, 0xffffffffu, v.end);
          break;
        }
        const uint32_t lim = v.q + ((k + 1u) << seg_shift);
        v.end = lim + lag - 16u;
        sfb::lzw_run<true>(v, s, lim & ~1023u, v.end);
        __syncwarp();
        if (l == 0) { state[0] = s.W; state[1] = s.cur; state[2] = s.c_o; state[3] = s.c_end; state[4] = s.c_d; }
        __syncwarp();
      }
    });
  for (auto& t : lanes) t.join();
  emu_warp = nullptr;
}
