// TEST INFRASTRUCTURE ONLY — CPU emulation of the CUDA kernels' logic (see cuda_shim.h).
// Pass 1 (huff_lanes.cuh) runs one emulated lane per stream; pass 2 is either the real kernel
// on 32 host threads (emu_lz.cpp, mode 1) or a scalar restatement of the in-place token
// format (mode 0, fast: it checks what pass 1 wrote, not how pass 2 is implemented).
#define SFB_CPU_EMU 1
#include "cuda_shim.h"

#include <vector>

// the kernel's dynamic shared memory: one emulated lane, laid out exactly as on the device
static thread_local uint16_t* emu_smem = nullptr;
#define SFB_EMU_SMEM emu_smem
static unsigned long long emu_stat_tokens = 0, emu_stat_slow_tokens = 0, emu_stat_long_tokens = 0;
#define SFB_STAT(name) (++emu_stat_##name)
#include "../../starflate_b200/csrc/huff_lanes.cuh"

using EmuCfg = sfb::Cfg<SFB_EMU_ROOT_LIT, SFB_EMU_ROOT_DIST, SFB_EMU_POOL, 1>;
using EmuSmall = sfb::Cfg<6, 5, 96, 1>;  // the product's small geometry (capi.cu: SmallCfg)
static unsigned long long emu_stat_deferred = 0;

extern "C" void emu_lz_resolve(uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* written,
                               const uint32_t* match_bits, uint64_t n);  // dst_base 128-byte aligned
extern "C" void emu_lz_resolve_segments(uint8_t* dst_base, uint64_t dst_off, uint64_t written,
                                        const uint32_t* match_bits, uint32_t seg_shift, uint32_t lag);

// the in-place token format of huff_lanes.cuh, resolved byte by byte
static void scalar_resolve(uint8_t* dst_base, uint64_t off, uint64_t written, const uint32_t* bits)
{
  uint8_t* d = dst_base + off;
  uint64_t p = 0;
  while (p < written) {
    const uint64_t k = off + p;
    if ((bits[k >> 5] >> (k & 31)) & 1u) {
      const unsigned len = d[p] + 3u;
      const unsigned dist = (d[p + 1] | (d[p + 2] << 8)) + 1u;
      for (unsigned j = 0; j < len; ++j) d[p + j] = d[p + j - dist];
      p += len;
    } else {
      ++p;
    }
  }
}

// Each stream is run in private, padded copies of its src/dst regions that keep the original
// address alignment (mod 16) and the original dst offset modulo 32 (the bitmap phase).  The
// kernels' aligned word accesses may legitimately touch the padding of src (reads) but must
// never modify anything outside [dst, dst+cap): the canaries around dst are verified.
// Returns 0, or 1000+i if stream i wrote out of bounds.
extern "C" int emu_decompress_batch(const uint8_t* src, const uint64_t* src_off,
                                    const uint64_t* src_len, uint8_t* dst,
                                    const uint64_t* dst_off, const uint64_t* dst_cap,
                                    uint8_t* status, uint64_t* written, uint64_t n, int warp_pass2,
                                    int small_first)
{
  std::vector<uint16_t> smem((EmuCfg::SMEM_BYTES > EmuSmall::SMEM_BYTES ? EmuCfg::SMEM_BYTES : EmuSmall::SMEM_BYTES) / 2 + 64, 0xDEAD);
  std::vector<uint32_t> lens(sfb::SCRATCH_WORDS * 32, 0xDEADBEEFu);
  emu_smem = smem.data();
  blockDim.x = 1;  // one emulated lane
  gridDim.x = 1;
  constexpr size_t PAD = 320;
  for (uint64_t i = 0; i < n; ++i) {
    const uint8_t* s = src + src_off[i];
    uint8_t* d = dst + dst_off[i];
    std::vector<uint8_t> sbuf(src_len[i] + 2 * PAD + 16, 0xEE);
    std::vector<uint8_t> dbuf(dst_cap[i] + 2 * PAD + 16, 0xC3);
    auto place = [](std::vector<uint8_t>& v, const void* like) {
      uint8_t* p = v.data() + PAD;
      while ((reinterpret_cast<uintptr_t>(p) & 15u) != (reinterpret_cast<uintptr_t>(like) & 15u)) ++p;
      return p;
    };
    uint8_t* sp = place(sbuf, s);
    uint8_t* dp = place(dbuf, d);
    if (src_len[i]) std::memcpy(sp, s, src_len[i]);
    if (dst_cap[i]) std::memcpy(dp, d, dst_cap[i]);
    // the kernels want a 128-byte aligned dst base: the aligned address below dp, with the
    // stream at the offset that falls out (its low bits follow the caller's layout mod 16)
    uint8_t* dbase = reinterpret_cast<uint8_t*>(reinterpret_cast<uintptr_t>(dp) & ~uintptr_t{127});
    const uint64_t doff = static_cast<uint64_t>(dp - dbase);
    std::vector<uint32_t> bits((doff + dst_cap[i]) / 32 + 8, 0u);
    unsigned long long counter = 0;
    const uint64_t zero = 0;
    uint64_t wr = 0;
    sfb::BatchArgs a{};
    a.src_base = sp;
    a.src_off = &zero;
    a.src_len = &src_len[i];
    a.dst_base = dbase;
    a.dst_delta = 0;
    a.dst_off = &doff;
    a.dst_cap = &dst_cap[i];
    a.status = &status[i];
    a.written = &wr;
    a.n = 1;
    a.idx_base = 0;
    a.group_counter = &counter;
    a.lens_scratch = lens.data();
    a.match_bits = bits.data();
    a.defer_list = nullptr;
    a.defer_count = nullptr;
    a.todo_list = nullptr;
    a.todo_count = nullptr;
    a.no_pair = 0;
    a.handled = nullptr;
    a.start_bit = nullptr;
    a.start_out = nullptr;
    a.blk_end = nullptr;
    threadIdx.x = 0;
    blockIdx.x = 0;
    uint32_t handed[1] = {0xffffffffu};
    unsigned long long n_handed = 0;
    bool run_large = true;
    if (small_first) {  // as capi.cu does: small geometry first, the large one for what it hands on
      a.defer_list = handed;
      a.defer_count = &n_handed;
      sfb::huff_lanes_kernel<EmuSmall>(a);
      run_large = n_handed != 0;
      emu_stat_deferred += n_handed;
      a.defer_list = nullptr;
      a.defer_count = nullptr;
      a.todo_list = handed;
      a.todo_count = &n_handed;
      counter = 0;
    }
    if (run_large) sfb::huff_lanes_kernel<EmuCfg>(a);
    if (wr > dst_cap[i]) return 2000 + static_cast<int>(i);
    if (warp_pass2 == 2) emu_lz_resolve_segments(dbase, doff, wr, bits.data(), 11, 4096);  // (2 KiB segments)
    else if (warp_pass2) emu_lz_resolve(dbase, &doff, &wr, bits.data(), 1);
    else scalar_resolve(dbase, doff, wr, bits.data());
    if (written) written[i] = wr;
    for (uint8_t* q = dbuf.data(); q < dp; ++q)
      if (*q != 0xC3) return 1000 + static_cast<int>(i);
    for (uint8_t* q = dp + dst_cap[i]; q < dbuf.data() + dbuf.size(); ++q)
      if (*q != 0xC3) return 1000 + static_cast<int>(i);
    if (dst_cap[i]) std::memcpy(d, dp, dst_cap[i]);
  }
  return 0;
}

// Size discovery (huff_lanes_kernel<C, true>): no dst at all — every dst-side pointer is null, so
// any store the counting mode did would crash the test.
extern "C" void emu_decompressed_size(const uint8_t* src, const uint64_t* src_off, const uint64_t* src_len,
                                      uint8_t* status, uint64_t* size, uint64_t n)
{
  std::vector<uint16_t> smem(EmuCfg::SMEM_BYTES / 2 + 64, 0xDEAD);
  std::vector<uint32_t> lens(sfb::SCRATCH_WORDS * 32, 0xDEADBEEFu);
  emu_smem = smem.data();
  blockDim.x = 1;
  gridDim.x = 1;
  for (uint64_t i = 0; i < n; ++i) {
    std::vector<uint8_t> sbuf(src_len[i] + 64, 0xEE);
    if (src_len[i]) std::memcpy(sbuf.data() + 16, src + src_off[i], src_len[i]);
    unsigned long long counter = 0;
    const uint64_t zero = 0;
    sfb::BatchArgs a{};
    a.src_base = sbuf.data() + 16;
    a.src_off = &zero;
    a.src_len = &src_len[i];
    a.status = &status[i];
    a.written = &size[i];
    a.n = 1;
    a.group_counter = &counter;
    a.lens_scratch = lens.data();
    threadIdx.x = 0;
    blockIdx.x = 0;
    sfb::huff_lanes_kernel<EmuCfg, true>(a);
  }
}

// Chunked input (BatchArgs::start_bit / start_out / blk_end): one stream, decoding from bit
// `start_bit` with `start_out` bytes of earlier output at the front of dst (128-byte aligned here).
// -> status; written (incl. start_out), the last block boundary passed (bit, output position).
extern "C" int emu_decompress_resume(const uint8_t* src, uint64_t src_len, uint64_t start_bit, uint8_t* dst,
                                     uint64_t start_out, uint64_t cap, uint64_t* written, uint64_t* blk_end)
{
  std::vector<uint16_t> smem(EmuCfg::SMEM_BYTES / 2 + 64, 0xDEAD);
  std::vector<uint32_t> lens(sfb::SCRATCH_WORDS * 32, 0xDEADBEEFu);
  emu_smem = smem.data();
  blockDim.x = 1;
  gridDim.x = 1;
  std::vector<uint8_t> sbuf(src_len + 64, 0xEE);
  if (src_len) std::memcpy(sbuf.data() + 16, src, src_len);
  std::vector<uint8_t> dbuf(cap + 512, 0xC3);
  uint8_t* dbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dbuf.data()) + 255) & ~uintptr_t{127});
  std::memcpy(dbase, dst, cap);
  std::vector<uint32_t> bits(cap / 32 + 8, 0u);
  unsigned long long counter = 0;
  const uint64_t zero = 0;
  uint8_t st = 0xEE;
  uint64_t wr = 0;
  sfb::BatchArgs a{};
  a.src_base = sbuf.data() + 16;
  a.src_off = &zero;
  a.src_len = &src_len;
  a.dst_base = dbase;
  a.dst_off = &zero;
  a.dst_cap = &cap;
  a.status = &st;
  a.written = &wr;
  a.n = 1;
  a.group_counter = &counter;
  a.lens_scratch = lens.data();
  a.match_bits = bits.data();
  a.start_bit = &start_bit;
  a.start_out = &start_out;
  a.blk_end = blk_end;
  threadIdx.x = 0;
  blockIdx.x = 0;
  sfb::huff_lanes_kernel<EmuCfg>(a);
  scalar_resolve(dbase, 0, wr, bits.data());
  std::memcpy(dst, dbase, cap);
  *written = wr;
  return st;
}

extern "C" void emu_stats(unsigned long long* out)
{
  out[0] = emu_stat_tokens;
  out[1] = emu_stat_slow_tokens;
  out[2] = emu_stat_deferred;
  out[3] = emu_stat_long_tokens;
}
