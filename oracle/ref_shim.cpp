// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Thin extern "C" veneer over the UNMODIFIED reference implementation so that
// Python tests / bench.py can call it through ctypes.  This file contains no
// algorithm: it includes the reference headers from -I/root/reference and
// forwards to
//   starflate::decompress            (/root/reference/src/decompress.cpp:402)
//   starflate::detail::read_header   (/root/reference/src/decompress.cpp:370)
//   starflate::detail::copy_from_before (/root/reference/src/decompress.cpp:388)
// It is compiled together with /root/reference/src/decompress.cpp by
// oracle/Makefile into oracle/_ref/ (git-ignored; travels to the GPU box).
#include "src/decompress.hpp"

#include <cstddef>
#include <cstdint>
#include <span>
#include <thread>
#include <vector>

extern "C" {

// returns the numeric value of starflate::DecompressStatus
int ref_decompress(const std::uint8_t* src, std::size_t src_len,
                   std::uint8_t* dst, std::size_t dst_cap)
{
  const std::span<const std::byte> s{reinterpret_cast<const std::byte*>(src), src_len};
  const std::span<std::byte> d{reinterpret_cast<std::byte*>(dst), dst_cap};
  return static_cast<int>(starflate::decompress(s, d));
}

// out[0]=has_value, out[1]=final, out[2]=type, out[3]=error status,
// out[4]=bits consumed
void ref_read_header(const std::uint8_t* src, std::size_t bit_size,
                     std::uint8_t bit_offset, int* out)
{
  starflate::huffman::bit_span bits{reinterpret_cast<const std::byte*>(src), bit_size, bit_offset};
  const auto before = bits.size();
  const auto h = starflate::detail::read_header(bits);
  out[0] = h.has_value() ? 1 : 0;
  out[1] = h.has_value() ? static_cast<int>(h->final) : 0;
  out[2] = h.has_value() ? static_cast<int>(h->type) : 0;
  out[3] = h.has_value() ? 0 : static_cast<int>(h.error());
  out[4] = static_cast<int>(before - bits.size());
}

void ref_copy_from_before(std::uint8_t* buf, std::size_t buf_len,
                          std::size_t dst_index, std::uint16_t distance,
                          std::uint16_t n)
{
  std::span<std::byte> s{reinterpret_cast<std::byte*>(buf), buf_len};
  starflate::detail::copy_from_before(
      distance, s.begin() + static_cast<std::ptrdiff_t>(dst_index), n);
}

// Batched driver used as the CPU baseline: one stream per thread-slot,
// static contiguous partition over `threads` std::threads.
void ref_decompress_batch(const std::uint8_t* src, const std::uint64_t* src_off,
                          const std::uint64_t* src_len, std::uint8_t* dst,
                          const std::uint64_t* dst_off,
                          const std::uint64_t* dst_cap, std::uint8_t* status,
                          std::uint64_t n, int threads)
{
  if (threads < 1) threads = 1;
  auto work = [&](std::uint64_t lo, std::uint64_t hi) {
    for (std::uint64_t i = lo; i < hi; ++i) {
      status[i] = static_cast<std::uint8_t>(
          ref_decompress(src + src_off[i], src_len[i], dst + dst_off[i], dst_cap[i]));
    }
  };
  std::vector<std::thread> pool;
  const std::uint64_t per = (n + static_cast<std::uint64_t>(threads) - 1) /
                            static_cast<std::uint64_t>(threads);
  for (int t = 0; t < threads; ++t) {
    const std::uint64_t lo = per * static_cast<std::uint64_t>(t);
    const std::uint64_t hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    pool.emplace_back(work, lo, hi);
  }
  for (auto& t : pool) t.join();
}

}  // extern "C"
