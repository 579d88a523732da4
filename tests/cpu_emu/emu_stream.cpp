// TEST INFRASTRUCTURE ONLY — the single-stream pass-1 kernel (huff_stream.cuh: 32 lanes decode 32
// spans of one block speculatively) run on the host, one warp = 32 threads in lock-step through
// real barriers (see cuda_shim_warp.h).
#define SFB_CPU_EMU 1
#include "cuda_shim_warp.h"

#include <thread>
#include <vector>

static uint16_t* emu_stream_smem = nullptr;  // one warp's shared memory, seen by all its lanes
#define SFB_EMU_SMEM emu_stream_smem
#include "../../starflate_b200/csrc/huff_stream.cuh"
static uint32_t* emu_crc_smem = nullptr;
#define SFB_EMU_CRC_SMEM emu_crc_smem
#include "../../starflate_b200/csrc/container.cuh"

using StreamCfg = sfb::Cfg<7, 5, 96, 1>;  // as capi.cu

// one stream; dst_base must be 128-byte aligned, the stream's region starts at dst_base + dst_off
static uint32_t emu_rec_cap = 1u << 16;
extern "C" void emu_set_rec_cap(uint32_t cap) { emu_rec_cap = cap; }

static void run_warp(const sfb::StreamArgs& a)
{
  std::vector<std::thread> lanes;
  for (unsigned l = 0; l < 32; ++l)
    lanes.emplace_back([&a, l] {
      threadIdx.x = l;
      blockIdx.x = 0;
      blockDim.x = 32;
      gridDim.x = 1;
      sfb::huff_stream_kernel<StreamCfg>(a);
    });
  for (auto& t : lanes) t.join();
}

// The GPU's block finder (find_candidates_kernel, verify_candidates_kernel) on one host thread:
// the bit positions it accepts as dynamic-block headers -> out[0 .. return value)
extern "C" uint32_t emu_find_blocks(const uint8_t* src, uint64_t src_len, uint64_t* out, uint32_t max_out)
{
  std::vector<uint8_t> sbuf(src_len + 64, 0xEE);  // as the decoder sees it: not 4-byte aligned
  if (src_len) std::memcpy(sbuf.data() + 17, src, src_len);
  const uint64_t zero = 0;
  const uint32_t cand_cap = static_cast<uint32_t>(8 * src_len / 16 + 64), job_cap = cand_cap;
  std::vector<uint64_t> cand(cand_cap);
  std::vector<sfb::BlockJob> jobs(job_cap);
  uint32_t tab_size = 64;
  while (tab_size < 2 * job_cap) tab_size *= 2;
  std::vector<uint32_t> tab(tab_size, 0u);
  uint32_t cand_count = 0, job_count = 0, tail = 0;
  sfb::FindArgs f{};
  f.src_base = sbuf.data() + 17;
  f.src_off = &zero;
  f.src_len = &src_len;
  f.idx = 0;
  f.cand = cand.data();
  f.cand_count = &cand_count;
  f.cand_cap = cand_cap;
  f.jobs = jobs.data();
  f.job_count = &job_count;
  f.job_cap = job_cap;
  f.job_tab = tab.data();
  f.tab_mask = tab_size - 1;
  f.tail_job = &tail;
  threadIdx.x = 0;
  blockIdx.x = 0;
  blockDim.x = 1;
  gridDim.x = 1;
  sfb::find_candidates_kernel(f);
  sfb::verify_candidates_kernel(f);
  uint32_t n = 0;
  for (uint32_t j = 1; j < job_count && j < job_cap && n < max_out; ++j) out[n++] = jobs[j].start_bit;
  return n;
}

// n_cand > 0: the blocks-side-by-side route of block_finder.cuh with the given candidate block
// starts (bit positions; the GPU's finder kernels are replaced by the caller's list, which may
// hold true starts, wrong ones and duplicates): count jobs, chain, writing jobs.
// stats[0] = jobs on the chain, stats[1] = tail job
extern "C" void emu_huff_stream_jobs(const uint8_t* src, uint64_t src_len, uint8_t* dst_base, uint64_t dst_off,
                                     uint64_t dst_cap, uint32_t* match_bits, uint8_t* status, uint64_t* written,
                                     const uint64_t* cand, uint32_t n_cand, uint32_t* stats)
{
  EmuWarp warp;
  emu_warp = &warp;
  std::vector<uint16_t> smem(StreamCfg::SMEM_BYTES / 2 + 64, 0xDEAD);
  std::vector<uint32_t> lens(sfb::SCRATCH_WORDS * 32, 0xDEADBEEFu);
  emu_stream_smem = smem.data();
  const uint64_t zero = 0;
  uint32_t tab_size = 64;
  while (tab_size < 2 * (n_cand + 1)) tab_size *= 2;
  std::vector<uint32_t> job_tab(tab_size, 0u);
  std::vector<sfb::BlockJob> jobs(n_cand + 1);
  for (uint32_t k = 0; k <= n_cand; ++k) {
    sfb::BlockJob jb{};
    jb.start_bit = k ? cand[k - 1] : 0;
    jb.end_bit = jb.start_bit;
    jb.base = sfb::JOB_NONE;
    jb.flags = sfb::JOB_ENDS;
    jobs[k] = jb;
    if (k) {
      uint32_t slot = sfb::job_slot(jb.start_bit, tab_size - 1);
      while (job_tab[slot]) slot = (slot + 1) & (tab_size - 1);
      job_tab[slot] = k;
    }
  }
  uint32_t job_count = n_cand + 1, tail_job = 0, rec_count = 0;
  std::vector<sfb::WinRec> recs(emu_rec_cap + 1);
  sfb::StreamArgs a{};
  a.src_base = src;
  a.src_off = &zero;
  a.src_len = &src_len;
  a.dst_base = dst_base;
  a.dst_delta = 0;
  a.dst_off = &dst_off;
  a.dst_cap = &dst_cap;
  a.status = status;
  a.written = written;
  a.list = nullptr;
  a.idx_base = 0;
  a.n = 1;
  a.lens_scratch = lens.data();
  a.match_bits = match_bits;
  a.jobs = jobs.data();
  a.job_count = &job_count;
  a.job_cap = n_cand + 1;
  a.job_tab = job_tab.data();
  a.tab_mask = tab_size - 1;
  a.tail_job = &tail_job;
  a.recs = recs.data();
  a.rec_count = &rec_count;
  a.rec_cap = emu_rec_cap;
  unsigned long long counter = 0;
  a.stream_counter = &counter;
  a.mode = 1;
  run_warp(a);
  sfb::FindArgs f{};
  f.idx = 0;
  f.jobs = jobs.data();
  f.job_count = &job_count;
  f.job_cap = n_cand + 1;
  f.tail_job = &tail_job;
  f.dst_cap = &dst_cap;
  threadIdx.x = 0;
  blockIdx.x = 0;
  sfb::chain_kernel(f);
  counter = 0;
  a.mode = 2;
  run_warp(a);
  if (stats) {
    uint32_t on = 0;
    for (auto& jb : jobs) on += jb.base != sfb::JOB_NONE;
    stats[0] = on;
    stats[1] = tail_job;
  }
  emu_warp = nullptr;
  emu_stream_smem = nullptr;
}

extern "C" void emu_huff_stream(const uint8_t* src, uint64_t src_len, uint8_t* dst_base, uint64_t dst_off,
                                uint64_t dst_cap, uint32_t* match_bits, uint8_t* status, uint64_t* written)
{
  EmuWarp warp;
  emu_warp = &warp;
  std::vector<uint16_t> smem(StreamCfg::SMEM_BYTES / 2 + 64, 0xDEAD);
  std::vector<uint32_t> lens(sfb::SCRATCH_WORDS * 32, 0xDEADBEEFu);
  emu_stream_smem = smem.data();
  unsigned long long counter = 0;
  const uint64_t zero = 0;
  sfb::StreamArgs a{};
  a.src_base = src;
  a.src_off = &zero;
  a.src_len = &src_len;
  a.dst_base = dst_base;
  a.dst_delta = 0;
  a.dst_off = &dst_off;
  a.dst_cap = &dst_cap;
  a.status = status;
  a.written = written;
  a.list = nullptr;
  a.idx_base = 0;
  a.n = 1;
  a.stream_counter = &counter;
  a.lens_scratch = lens.data();
  a.match_bits = match_bits;
  run_warp(a);
  emu_warp = nullptr;
  emu_stream_smem = nullptr;
}

extern "C" void emu_lz_resolve(uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* written,
                               const uint32_t* match_bits, uint64_t n);

// One stream through the single-stream pass 1 (above) and the real pass 2 (emu_lz.cpp), in a
// private padded copy of dst with canaries.  `dst_phase` (0..127) places the region relative to a
// 128-byte boundary.  Returns 0, or 1 if anything outside [dst, dst+cap) changed.
static int stream_decompress_impl(const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t dst_cap,
                                  uint32_t dst_phase, uint8_t* status, uint64_t* written, const uint64_t* cand,
                                  uint32_t n_cand, uint32_t* stats);

extern "C" int emu_stream_decompress(const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t dst_cap,
                                     uint32_t dst_phase, uint8_t* status, uint64_t* written)
{
  return stream_decompress_impl(src, src_len, dst, dst_cap, dst_phase, status, written, nullptr, 0, nullptr);
}

// the same through the blocks-side-by-side route, see emu_huff_stream_jobs
extern "C" int emu_stream_decompress_jobs(const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t dst_cap,
                                          uint32_t dst_phase, uint8_t* status, uint64_t* written,
                                          const uint64_t* cand, uint32_t n_cand, uint32_t* stats)
{
  return stream_decompress_impl(src, src_len, dst, dst_cap, dst_phase, status, written, cand, n_cand, stats);
}

static int stream_decompress_impl(const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t dst_cap,
                                  uint32_t dst_phase, uint8_t* status, uint64_t* written, const uint64_t* cand,
                                  uint32_t n_cand, uint32_t* stats)
{
  constexpr size_t PAD = 512;
  std::vector<uint8_t> sbuf(src_len + 64, 0xEE);
  if (src_len) std::memcpy(sbuf.data() + 16, src, src_len);
  std::vector<uint8_t> dbuf(dst_cap + 2 * PAD + 256, 0xC3);
  uint8_t* dbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dbuf.data()) + PAD) & ~uintptr_t{127});
  const uint64_t doff = dst_phase & 127u;
  uint8_t* dp = dbase + doff;
  if (dst_cap) std::memcpy(dp, dst, dst_cap);
  std::vector<uint32_t> bits((doff + dst_cap) / 32 + 8, 0u);
  uint64_t wr = 0;
  if (cand)
    emu_huff_stream_jobs(sbuf.data() + 16, src_len, dbase, doff, dst_cap, bits.data(), status, &wr, cand, n_cand,
                         stats);
  else
    emu_huff_stream(sbuf.data() + 16, src_len, dbase, doff, dst_cap, bits.data(), status, &wr);
  if (wr > dst_cap) return 2;
  emu_lz_resolve(dbase, &doff, &wr, bits.data(), 1);
  *written = wr;
  for (uint8_t* q = dbuf.data(); q < dp; ++q)
    if (*q != 0xC3) return 1;
  for (uint8_t* q = dp + dst_cap; q < dbuf.data() + dbuf.size(); ++q)
    if (*q != 0xC3) return 1;
  if (dst_cap) std::memcpy(dst, dp, dst_cap);
  return 0;
}

// One zlib / gzip container (container.cuh): the parse kernel on one host thread, the payload
// through the single-stream pass 1 + real pass 2 above, the check kernel on the 32 threads of
// one emulated warp.  dst must be dst_cap bytes; returns 0, or != 0 if something outside dst changed.
extern "C" int emu_container(const uint8_t* src, uint64_t src_len, uint32_t container, uint8_t* dst, uint64_t dst_cap,
                             uint8_t* status, uint64_t* written)
{
  std::vector<uint8_t> sbuf(src_len + 64, 0xEE);
  if (src_len) std::memcpy(sbuf.data() + 19, src, src_len);
  const uint64_t zero = 0;
  uint64_t pay_off = 0, pay_len = 0;
  uint32_t kind = 0, expect = 0, isize = 0;
  sfb::ContainerArgs c{};
  c.src_base = sbuf.data() + 19;
  c.src_off = &zero;
  c.src_len = &src_len;
  c.container = container;
  c.n = 1;
  c.pay_off = &pay_off;
  c.pay_len = &pay_len;
  c.kind = &kind;
  c.expect = &expect;
  c.isize = &isize;
  c.dst_base = dst;
  c.dst_off = &zero;
  c.status = status;
  c.written = written;
  threadIdx.x = 0;
  blockIdx.x = 0;
  blockDim.x = 1;
  gridDim.x = 1;
  sfb::container_parse_kernel(c);
  const int rc = stream_decompress_impl(sbuf.data() + 19 + pay_off, pay_len, dst, dst_cap, 37, status, written,
                                        nullptr, 0, nullptr);
  if (rc) return rc;
  EmuWarp warp;
  emu_warp = &warp;
  std::vector<uint32_t> tab(4 * 256 * 32);
  emu_crc_smem = tab.data();
  std::vector<std::thread> lanes;
  for (unsigned l = 0; l < 32; ++l)
    lanes.emplace_back([&c, l] {
      threadIdx.x = l;
      blockIdx.x = 0;
      blockDim.x = 32;
      gridDim.x = 1;
      sfb::container_check_kernel(c);
    });
  for (auto& t : lanes) t.join();
  emu_warp = nullptr;
  emu_crc_smem = nullptr;
  return 0;
}
